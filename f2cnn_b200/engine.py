"""Host-side engine: plans, batches and launch helpers over the C ABI.

PyTorch is used only as plumbing (device memory, pinned host buffers, streams); every
computation of the hot path happens in libf2cnn_b200.so.  Nothing here falls back to
numpy/scipy/torch math: without the library or a CUDA device the calls raise."""
import ctypes
import hashlib
import threading

import numpy as np
import torch

from . import _native
from ._native import F2_F32, F2_F64, F2_I16, RunArgs, check

_NP2F2 = {np.dtype(np.int16): F2_I16, np.dtype(np.float32): F2_F32, np.dtype(np.float64): F2_F64}
_T2F2 = {torch.int16: F2_I16, torch.float32: F2_F32, torch.float64: F2_F64}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("f2cnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _stream_ptr(stream):
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def lowpass_coefficients(cutoff_hz):
    """(b0, a1) of butter(1, cutoff/8000, 'low') -- reference EnvelopeExtraction.py:47."""
    b0 = ctypes.c_double()
    a1 = ctypes.c_double()
    check(_native.lib().f2_lowpass_coefficients(float(cutoff_hz), ctypes.byref(b0), ctypes.byref(a1)))
    return b0.value, a1.value


class Plan:
    """One gammatone filterbank on one device.  coefs = make_erb_filters output (C,10)."""

    def __init__(self, coefs, device=None):
        _require_cuda()
        coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        if coefs.ndim != 2 or coefs.shape[1] != 10:
            raise ValueError("coefs must be (n_channels, 10) as returned by make_erb_filters")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.n_channels = int(coefs.shape[0])
        self.coefs = coefs
        self._h = ctypes.c_void_p()
        check(_native.lib().f2_plan_create(coefs.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), self.n_channels,
                                           self.device.index, ctypes.byref(self._h)))
        self._ws = None

    def __del__(self):
        try:
            h = getattr(self, "_h", None)
            if h is not None and h.value:
                _native.lib().f2_plan_destroy(h)
                self._h = None
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass

    def set_warmup(self, w_imag=0, w_edge=0, w_casc=0):
        check(_native.lib().f2_plan_set_warmup(self._h, int(w_imag), int(w_edge), int(w_casc)))

    def get_warmup(self):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(_native.lib().f2_plan_get_warmup(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def batch(self, lengths, step=160, phase=0, target_items=0):
        return Batch(self, lengths, step, phase, target_items)

    def workspace(self, nbytes):
        """Grow-only scratch shared by this plan's launches (caller-owned per the C ABI)."""
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    # -- stand-alone envelope of matrix rows (ExtractEnvelopeFromMatrix on foreign data) ----
    def envelope_rows(self, matrix_dev, lpf, cutoff, out_dtype=torch.float64, stream=None, op=0):
        rows, n = matrix_dev.shape
        out = torch.empty((rows, n), dtype=out_dtype, device=self.device)
        if rows == 0 or n == 0:
            return out
        L = _native.lib()
        ws = self.workspace(L.f2_envelope_rows_workspace_bytes(rows, n))
        check(L.f2_rows_op(self._h, _ptr(matrix_dev), _T2F2[matrix_dev.dtype], rows, n, int(op), int(bool(lpf)),
                           float(cutoff), _ptr(out), _T2F2[out_dtype], _ptr(ws), ws.numel(), _stream_ptr(stream)))
        return out


class Batch:
    """A set of utterances (by length) prepared for repeated runs on one plan."""

    def __init__(self, plan, lengths, step=160, phase=0, target_items=0):
        self.plan = plan
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64).reshape(-1)
        self.n_utts = int(self.lengths.shape[0])
        self.step, self.phase = int(step), int(phase)
        self._h = ctypes.c_void_p()
        L = _native.lib()
        check(L.f2_batch_create(plan._h, self.lengths.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), self.n_utts,
                                self.step, self.phase, int(target_items), ctypes.byref(self._h)))
        self.total_samples = int(L.f2_batch_total_samples(self._h))
        self.total_frames = int(L.f2_batch_total_frames(self._h))
        self.num_items = int(L.f2_batch_num_items(self._h))
        self.frame_offsets = np.zeros(self.n_utts + 1, dtype=np.int64)
        check(L.f2_batch_frame_offsets(self._h, self.frame_offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
        self.sample_offsets = np.zeros(self.n_utts + 1, dtype=np.int64)
        np.cumsum(self.lengths, out=self.sample_offsets[1:])
        self._grid = {}

    def __del__(self):
        try:
            h = getattr(self, "_h", None)
            if h is not None and h.value:
                _native.lib().f2_batch_destroy(h)
                self._h = None
        except Exception:
            pass

    def workspace_bytes(self, want_full_gfb=False, want_full_env=False):
        return int(_native.lib().f2_batch_workspace_bytes(self._h, int(want_full_gfb), int(want_full_env)))

    def grid_windows(self, dots=11):
        """The label grid of LabelDataGenerator.py:38-50 on this batch's decimated frames: utterance
        u has nb = max(n // step - dots - 1, 0) windows, window k = frames k .. k+dots-1.  Returns the
        device int64 row offsets (n_utts+1) `run(windows=...)` wants, and the total row count."""
        key = int(dots)
        if key not in self._grid:
            nb = np.maximum(self.lengths // self.step - key - 1, 0)
            off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
            self._grid[key] = (torch.from_numpy(off).to(self.plan.device), int(off[-1]))
        return self._grid[key]

    def run(self, wave_dev, lpf=False, cutoff=100, gfb=None, env=None, env_t=False, dec=False, stream=None,
            out=None, fused_events=None, windows=None):
        """wave_dev: flat device tensor (int16/float32/float64) of total_samples.
        gfb/env: None or torch.float64/float32 -> (C*total_samples,) reference-layout blocks;
        env_t: time-major [total_samples, C] float32; dec: [total_frames, C] float32.
        `out` may carry preallocated tensors under the same keys.
        windows=(row_offsets_dev, dots[, out_tensor]): the fused kernel writes the (rows, dots, C)
        float32 windows itself (window k of utterance u = its decimated frames k..k+dots-1, rows
        row_offsets[u] + k); a stand-alone output mode, see grid_windows()."""
        plan = self.plan
        C = plan.n_channels
        if wave_dev.numel() != self.total_samples:
            raise ValueError("wave has %d samples, batch expects %d" % (wave_dev.numel(), self.total_samples))
        if wave_dev.device != plan.device:
            raise ValueError("wave is on %s, plan on %s" % (wave_dev.device, plan.device))
        res = {} if out is None else dict(out)
        if gfb is not None and "gfb" not in res:
            res["gfb"] = torch.empty(C * self.total_samples, dtype=gfb, device=plan.device)
        if env is not None and "env" not in res:
            res["env"] = torch.empty(C * self.total_samples, dtype=env, device=plan.device)
        if env_t and "env_t" not in res:
            res["env_t"] = torch.empty((self.total_samples, C), dtype=torch.float32, device=plan.device)
        if dec and "dec" not in res:
            res["dec"] = torch.empty((self.total_frames, C), dtype=torch.float32, device=plan.device)
        a = RunArgs()
        a.wave = wave_dev.data_ptr()
        a.wave_dtype = _T2F2[wave_dev.dtype]
        a.lpf = int(bool(lpf))
        a.cutoff_hz = float(cutoff) if lpf else 0.0
        g, e = res.get("gfb"), res.get("env")
        a.gfb = g.data_ptr() if g is not None else None
        a.gfb_dtype = _T2F2[g.dtype] if g is not None else F2_F64
        a.env = e.data_ptr() if e is not None else None
        a.env_dtype = _T2F2[e.dtype] if e is not None else F2_F64
        a.env_t = res["env_t"].data_ptr() if "env_t" in res else None
        a.dec = res["dec"].data_ptr() if "dec" in res else None
        if windows is not None:
            if g is not None or e is not None or "env_t" in res or "dec" in res:
                raise ValueError("windows is a stand-alone output mode")
            offs, dots = windows[0], int(windows[1])
            if offs.dtype != torch.int64 or offs.device != plan.device or offs.numel() != len(self.lengths) + 1:
                raise ValueError("windows offsets: int64 device tensor of n_utts+1 entries")
            if len(windows) > 2 and windows[2] is not None:
                res["windows"] = windows[2]
            elif "windows" not in res:
                res["windows"] = torch.empty((int(offs[-1].item()), dots, C), dtype=torch.float32, device=plan.device)
            w = res["windows"]
            if w.dtype != torch.float32 or not w.is_contiguous() or w.device != plan.device:
                raise ValueError("windows output: contiguous float32 tensor on the plan's device")
            a.windows, a.win_offsets, a.win_dots = w.data_ptr(), offs.data_ptr(), dots
        if fused_events is not None:  # (DeviceEvent, DeviceEvent) around the fused kernel
            a.ev_fused_start, a.ev_fused_stop = fused_events[0].handle, fused_events[1].handle
        need = self.workspace_bytes(g is not None, e is not None and "env_t" not in res)
        ws = plan.workspace(need)
        check(_native.lib().f2_batch_run(self._h, ctypes.byref(a), _ptr(ws), ws.numel(), _stream_ptr(stream)))
        return res


class DeviceEvent:
    """cudaEvent_t owned through the C ABI (recorded on the stream the kernels run on)."""

    def __init__(self):
        self.handle = ctypes.c_void_p()
        check(_native.lib().f2_event_create(ctypes.byref(self.handle)))

    def record(self, stream=None):
        check(_native.lib().f2_event_record(self.handle, _stream_ptr(stream)))

    def synchronize(self):
        check(_native.lib().f2_event_synchronize(self.handle))

    def elapsed_ms(self, stop):
        ms = ctypes.c_float()
        check(_native.lib().f2_event_elapsed_ms(self.handle, stop.handle, ctypes.byref(ms)))
        return float(ms.value)

    def __del__(self):
        try:
            if self.handle is not None and self.handle.value:
                _native.lib().f2_event_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ---- windowing helpers ---------------------------------------------------------------------
def gather_windows(frames_dev, base_rows_dev, dots, stride_rows=1, out=None, stream=None):
    """out[w, j, :] = frames[base_rows[w] + j*stride_rows, :]  (InputGenerator.py:73-80)."""
    n, C = int(base_rows_dev.numel()), int(frames_dev.shape[1])
    if out is None:
        out = torch.empty((n, dots, C), dtype=torch.float32, device=frames_dev.device)
    check(_native.lib().f2_gather_windows(_ptr(frames_dev), C, _ptr(base_rows_dev), n, int(dots), int(stride_rows),
                                          _ptr(out), _stream_ptr(stream)))
    return out


def gather_index(src_dev, idx_dev, out=None, stream=None):
    n, C = int(idx_dev.numel()), int(src_dev.shape[1])
    if out is None:
        out = torch.empty((n, C), dtype=torch.float32, device=src_dev.device)
    check(_native.lib().f2_gather_index(_ptr(src_dev), C, _ptr(idx_dev), n, _ptr(out), _stream_ptr(stream)))
    return out


def gather_windows_cn(env_dev, idx_dev, out=None, stream=None):
    """out[i, :] = float32(env[:, idx[i]]) for a (C, n) float64/float32 device matrix."""
    C, n = int(env_dev.shape[0]), int(env_dev.shape[1])
    m = int(idx_dev.numel())
    if out is None:
        out = torch.empty((m, C), dtype=torch.float32, device=env_dev.device)
    check(_native.lib().f2_gather_windows_cn(_ptr(env_dev), _T2F2[env_dev.dtype], C, n, _ptr(idx_dev), m, _ptr(out),
                                             _stream_ptr(stream)))
    return out


def dense_frames(env_t_dev, dots, step, i0, i1, normalize=False, out_dtype=torch.float64, stream=None):
    """Frames i0..i1 of Evaluating.py:70-78 (+ Training.normalizeInput when normalize)."""
    C = int(env_t_dev.shape[1])
    out = torch.empty((max(i1 - i0, 0), dots, C), dtype=out_dtype, device=env_t_dev.device)
    flag = torch.zeros(1, dtype=torch.int32, device=env_t_dev.device)
    check(_native.lib().f2_dense_frames(_ptr(env_t_dev), C, int(dots), int(step), int(i0), int(i1),
                                        int(bool(normalize)), _ptr(out), _T2F2[out_dtype], _ptr(flag),
                                        _stream_ptr(stream)))
    return out, flag


def label_fit(formant_dev, first_dev, center_dev, dots, step, stream=None):
    """(slope, intercept, r, p) per timepoint: LabelDataGenerator.py:60-68 for every item at once.
    formant_dev float64, first_dev int64, center_dev int32 device tensors; returns (N, 4) float64."""
    if formant_dev.dtype != torch.float64 or first_dev.dtype != torch.int64 or center_dev.dtype != torch.int32:
        raise TypeError("label_fit wants float64 formants, int64 first indices, int32 centers")
    n = int(first_dev.shape[0])
    if int(center_dev.shape[0]) != n:
        raise ValueError("first and center differ in length")
    out = torch.empty((n, 4), dtype=torch.float64, device=formant_dev.device)
    check(_native.lib().f2_label_fit(_ptr(formant_dev), _ptr(first_dev), _ptr(center_dev), n, int(dots), int(step),
                                     _ptr(out), _stream_ptr(stream)))
    return out


class WindowPipeline:
    """Host waves in, host (N, 2R+1, C) float32 windows out, for a whole corpus.

    The utterances are cut into `n_sub` contiguous sub-batches; while sub-batch i is being
    filtered on the compute stream, the windows of sub-batch i-1 travel to the host on a copy
    stream (double-buffered device output); the waves of all sub-batches are uploaded on a third
    stream right at the start (2 bytes per sample).  PCIe is the bottleneck of this path (the window tensor is 11x the decimated
    envelope), so hiding the 50-odd ms of compute behind the D2H copy is what matters.

    bases[u]: int64 array, for every window of utterance u the index (within the utterance)
    of its FIRST decimated frame; windows are `dots` consecutive frames (on-grid labels)."""

    def __init__(self, plan, lengths, bases, dots=11, step=160, phase=0, lpf=True, cutoff=50, n_sub=8):
        self.plan, self.dots, self.lpf, self.cutoff = plan, int(dots), bool(lpf), cutoff
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        U = int(lengths.shape[0])
        n_sub = max(1, min(int(n_sub), U))
        cum = np.concatenate([[0], np.cumsum(lengths)])
        # sub-batches of total/n_sub samples, except that the first few grow from an eighth of that: the
        # D2H copy -- the bottleneck -- cannot start before the first sub-batch is uploaded and computed
        per_sub = cum[-1] / n_sub
        marks, size, at = [0.0], per_sub / 8, 0.0
        while at + size < cum[-1]:
            at += size
            marks.append(at)
            size = min(per_sub, size * 2)
        cuts = [int(np.searchsorted(cum, m)) for m in marks] + [U]
        cuts[0] = 0
        cuts = sorted(set(cuts))
        dev = plan.device
        self.subs = []
        row = 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            # whole utterances only (target_items=1): same arithmetic as one big batch, bit for bit
            batch = plan.batch(lengths[a:b], step=step, phase=phase, target_items=1)
            base = np.concatenate([batch.frame_offsets[u - a] + np.asarray(bases[u], dtype=np.int64)
                                   for u in range(a, b)] + [np.zeros(0, dtype=np.int64)])
            # windows that are exactly the label grid (window k = frames k..k+dots-1, k < nb) are written by
            # the fused kernel itself; anything else goes through the decimated frames and the gather
            nb = np.maximum(lengths[a:b] // int(step) - self.dots - 1, 0)
            grid = all(len(bases[u]) == nb[u - a] and
                       (nb[u - a] == 0 or np.array_equal(np.asarray(bases[u]), np.arange(nb[u - a])))
                       for u in range(a, b))
            self.subs.append(dict(batch=batch, base=torch.from_numpy(base).to(dev), s0=int(cum[a]), s1=int(cum[b]),
                                  r0=row, r1=row + int(base.shape[0]),
                                  grid=batch.grid_windows(self.dots)[0] if grid else None))
            row += int(base.shape[0])
        self.n_windows = row
        C = plan.n_channels
        max_w = max(s["r1"] - s["r0"] for s in self.subs)
        max_f = max(s["batch"].total_frames for s in self.subs)
        self._wave_all = None           # device copy of the whole flat wave buffer (2 B per sample)
        self._total_s = int(cum[-1])
        self._win = [torch.empty((max(max_w, 1), self.dots, C), dtype=torch.float32, device=dev) for _ in range(2)]
        self._dec = torch.empty((max(max_f, 1), C), dtype=torch.float32, device=dev)
        self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self._ev_in = [torch.cuda.Event() for _ in self.subs]
        self._ev_done = [torch.cuda.Event() for _ in range(2)]
        self._ev_free = [torch.cuda.Event() for _ in range(2)]

    def run(self, wave_host, out_host):
        """wave_host: flat (pinned) host tensor of all samples; out_host: (N, dots, C) float32
        host tensor (pinned for full speed).  Returns after everything has landed."""
        comp = torch.cuda.current_stream()
        dev = self.plan.device
        if self._wave_all is None or self._wave_all.dtype != wave_host.dtype:
            self._wave_all = torch.empty(max(self._total_s, 1), dtype=wave_host.dtype, device=dev)
        self._s_in.wait_stream(comp)
        self._s_out.wait_stream(comp)
        # all uploads are queued at once (444 MB for the corpus): the H2D traffic is over after the first
        # few milliseconds and the D2H copies -- the bottleneck -- have the link to themselves afterwards
        with torch.cuda.stream(self._s_in):
            for i, sub in enumerate(self.subs):
                self._wave_all[sub["s0"]:sub["s1"]].copy_(wave_host[sub["s0"]:sub["s1"]], non_blocking=True)
                self._ev_in[i].record(self._s_in)
        for i, sub in enumerate(self.subs):
            k = i & 1
            comp.wait_event(self._ev_in[i])
            if i >= 2:
                comp.wait_event(self._ev_free[k])  # the D2H of sub-batch i-2 has drained this buffer
            n_w = sub["r1"] - sub["r0"]
            if sub["grid"] is not None and n_w:
                sub["batch"].run(self._wave_all[sub["s0"]:sub["s1"]], lpf=self.lpf, cutoff=self.cutoff,
                                 windows=(sub["grid"], self.dots, self._win[k][:n_w]))
            else:
                sub["batch"].run(self._wave_all[sub["s0"]:sub["s1"]], lpf=self.lpf, cutoff=self.cutoff,
                                 out={"dec": self._dec[:max(sub["batch"].total_frames, 1)]})
                if n_w:
                    gather_windows(self._dec, sub["base"], self.dots, 1, out=self._win[k][:n_w])
            self._ev_done[k].record(comp)
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(self._ev_done[k])
                if n_w:
                    out_host[sub["r0"]:sub["r1"]].copy_(self._win[k][:n_w], non_blocking=True)
                self._ev_free[k].record(self._s_out)
        comp.wait_stream(self._s_out)
        return out_host


class MultiGpuWindowPipeline:
    """WindowPipeline over several GPUs of one box from ONE process: the utterances are cut into
    contiguous ranges of about equal sample count, one per device; every device runs its own
    WindowPipeline into its slice of the shared host output.  Utterances are independent, so
    there is no collective -- the "gather" is each device's D2H copy landing at its row offset
    (SURVEY.md section 8e).  All launches are asynchronous; one host thread drives all devices."""

    def __init__(self, coefs, lengths, bases, devices=None, **kw):
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        U = int(lengths.shape[0])
        devices = list(devices)[:max(1, min(len(devices), U))]
        cum = np.concatenate([[0], np.cumsum(lengths)])
        cuts = [int(np.searchsorted(cum, cum[-1] * k / len(devices))) for k in range(len(devices) + 1)]
        cuts[0], cuts[-1] = 0, U
        self.parts = []
        row = 0
        for d, a, b in zip(devices, cuts[:-1], cuts[1:]):
            if b <= a:
                continue
            with torch.cuda.device(d):
                plan = plan_for(coefs, d)
                pipe = WindowPipeline(plan, lengths[a:b], bases[a:b], **kw)
            self.parts.append(dict(device=d, pipe=pipe, s0=int(cum[a]), s1=int(cum[b]), r0=row, r1=row + pipe.n_windows))
            row += pipe.n_windows
        self.n_windows = row

    def run(self, wave_host, out_host):
        for part in self.parts:
            with torch.cuda.device(part["device"]):
                part["pipe"].run(wave_host[part["s0"]:part["s1"]], out_host[part["r0"]:part["r1"]])
        for part in self.parts:
            torch.cuda.synchronize(part["device"])
        return out_host


# ---- plan cache for the numpy-in / numpy-out drop-in functions -----------------------------
_plans = {}
_plans_lock = threading.Lock()


def plan_for(coefs, device=None):
    """Plans are immutable; cache them by the bytes of the coefficient matrix."""
    _require_cuda()
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    dev = torch.cuda.current_device() if device is None else int(device)
    key = (hashlib.sha1(coefs.tobytes()).hexdigest(), coefs.shape, dev)
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = Plan(coefs, dev)
            _plans[key] = p
        return p


_generic_plan = {}


def any_plan(device=None):
    """A plan is only a device handle for the filterbank-independent entry points."""
    dev = torch.cuda.current_device() if device is None else int(device)
    with _plans_lock:
        for (k, shape, d), p in _plans.items():
            if d == dev:
                return p
    from .gammatone import filters
    return plan_for(filters.make_erb_filters(16000, filters.centre_freqs(16000, 4, 100)), dev)
