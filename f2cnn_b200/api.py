"""numpy-in / numpy-out entry points behind the reference-named modules.

Each function stages its inputs through pinned host memory, launches the C-ABI sequence on
the current CUDA stream and returns freshly allocated numpy arrays of the reference's dtype
and layout.  All arithmetic happens in libf2cnn_b200.so."""
import hashlib
import os
import threading
from collections import OrderedDict

import numpy as np
import torch

from . import engine

_WAVE_DTYPES = (np.int16, np.float32, np.float64)
_PIPELINE_BYTES = 256 << 20  # window tensors above this size go through engine.WindowPipeline


def _as_wave(wave):
    """The reference feeds scipy.signal.lfilter an int16 WAV array or a float64 noise-mixed
    array (Evaluating.py:200); anything else is promoted to float64 like lfilter would."""
    w = np.asarray(wave)
    if w.ndim != 1:
        raise ValueError("wave must be one-dimensional, got shape %s" % (w.shape,))
    if w.dtype not in [np.dtype(d) for d in _WAVE_DTYPES]:
        if w.dtype.kind in "iub" and w.dtype.itemsize <= 2:
            w = w.astype(np.int16) if w.dtype != np.uint16 else w.astype(np.float64)
        else:
            w = w.astype(np.float64)
    return np.ascontiguousarray(w)


def _to_device(arr, device):
    t = torch.from_numpy(arr)
    if arr.nbytes >= (1 << 16):
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


def _to_host(t):
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=t.numel() * t.element_size() >= (1 << 16))
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def erb_filterbank(wave, coefs, dtype=np.float64):
    """gammatone/filters.py:195-239 -> (C, n) float64 (the reference's dtype; float32 on request)."""
    w = _as_wave(wave)
    coefs = np.asarray(coefs, dtype=np.float64)
    plan = engine.plan_for(coefs)
    n = int(w.shape[0])
    if n == 0:
        return np.zeros((plan.n_channels, 0), dtype=dtype)
    batch = plan.batch([n])
    res = batch.run(_to_device(w, plan.device), gfb=torch.float64 if np.dtype(dtype) == np.float64 else torch.float32)
    return _to_host(res["gfb"]).reshape(plan.n_channels, n)


def filterbank_envelope(wave, coefs, LPF=False, CUTOFF=100, with_gfb=False, dtype=np.float64):
    """Fused erb_filterbank + ExtractEnvelopeFromMatrix on one waveform: the (C,n) envelope
    (and optionally the filterbank output) without the intermediate host round trip."""
    w = _as_wave(wave)
    coefs = np.asarray(coefs, dtype=np.float64)
    plan = engine.plan_for(coefs)
    n = int(w.shape[0])
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    if n == 0:
        z = np.zeros((plan.n_channels, 0), dtype=dtype)
        return (z, z.copy()) if with_gfb else z
    batch = plan.batch([n])
    res = batch.run(_to_device(w, plan.device), lpf=LPF, cutoff=CUTOFF, env=tdt, gfb=tdt if with_gfb else None)
    env = _to_host(res["env"]).reshape(plan.n_channels, n)
    if with_gfb:
        return _to_host(res["gfb"]).reshape(plan.n_channels, n), env
    return env


def gammatonegram(wave, coefs, hop, LPF=False, CUTOFF=100, dtype=np.float64):
    """Every `hop`-th sample of the envelope of `filterbank_envelope` -- (C, ceil(n / hop)), the samples at
    t = 0, hop, 2*hop, ... -- stored by the fused kernel itself: what a plot of a whole file needs
    (scripts/plotting/PlottingProcessing.py:101-109 takes the full-rate matrix for it)."""
    w = _as_wave(wave)
    coefs = np.asarray(coefs, dtype=np.float64)
    plan = engine.plan_for(coefs)
    n, hop = int(w.shape[0]), int(hop)
    if hop < 1:
        raise ValueError("hop must be >= 1")
    if n == 0:
        return np.zeros((plan.n_channels, 0), dtype=dtype)
    batch = plan.batch([n], step=hop, phase=0)
    dec = batch.run(_to_device(w, plan.device), lpf=LPF, cutoff=CUTOFF, dec=True)["dec"]
    return np.ascontiguousarray(_to_host(dec).T.astype(dtype, copy=False))


def extract_envelope_from_matrix(matrix, LPF=False, CUTOFF=100, dtype=np.float64):
    """scripts/processing/EnvelopeExtraction.py:51-67 on an arbitrary (rows, n) matrix:
    abs(paddedHilbert(row)) then lowPassFilter(row, CUTOFF) iff LPF -> float64 (float32 on request),
    same shape."""
    m = np.asarray(matrix)
    if m.ndim != 2:
        raise ValueError("matrix must be two-dimensional (channels x samples)")
    if m.dtype not in (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.int16)):
        m = m.astype(np.float64)
    m = np.ascontiguousarray(m)
    rows, n = m.shape
    if rows == 0:
        return np.zeros(m.shape, dtype=dtype)
    if n == 0:
        # paddedHilbert(empty) -> scipy.signal.hilbert raises "N must be positive."
        raise ValueError("N must be positive.")
    plan = engine.any_plan()
    out = plan.envelope_rows(_to_device(m, plan.device), LPF, CUTOFF,
                             out_dtype=torch.float64 if np.dtype(dtype) == np.float64 else torch.float32)
    return _to_host(out)


class MatrixStream:
    """Full-rate (C, n) matrices for MANY utterances -- what the file drivers `prepare filter` and `prepare
    envelope` write (GammatoneFiltering.py:69-83, EnvelopeExtraction.py:101-117) -- without paying the
    per-call set-up of the array functions: utterances are processed in batches that fill one of two
    pinned host slots; the matrices handed back are VIEWS into the slot, so writer threads save them
    without another copy while the next batch is computed into the other slot.  A slot is recycled when
    the futures registered with retire() for it have finished.

    mode "filterbank": waves -> erb_filterbank (+ envelope with with_env): process(list of waves).
    mode "envelope":   loaded .GFB matrices -> ExtractEnvelopeFromMatrix: process(list of (C, n) arrays)."""

    def __init__(self, coefs, mode="filterbank", LPF=False, CUTOFF=100, with_env=False, dtype=np.float64,
                 slot_bytes=768 << 20):
        # the envelope of a loaded matrix does not depend on a filterbank: any plan of the device will do
        self.plan = engine.any_plan() if coefs is None else engine.plan_for(np.asarray(coefs, dtype=np.float64))
        self.mode, self.lpf, self.cutoff, self.with_env = mode, bool(LPF), CUTOFF, bool(with_env)
        self.dtype = np.dtype(dtype)
        self.tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        self.slot_bytes = int(slot_bytes)
        self._slots = [torch.empty(self.slot_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        self._guards = [[], []]
        self._dev = torch.empty(self.slot_bytes, dtype=torch.uint8, device=self.plan.device)
        self._stage = None
        self._k = 0
        self.per_value = self.dtype.itemsize * (2 if (mode == "filterbank" and with_env) else 1)

    def batches(self, sizes):
        """Split a list of matrix sizes (elements: C * n per utterance) into runs that fit a slot."""
        out, cur, used = [], [], 0
        for i, sz in enumerate(sizes):
            need = int(sz) * self.per_value
            if need > self.slot_bytes:
                raise ValueError("one matrix of %d bytes does not fit a slot of %d" % (need, self.slot_bytes))
            if used + need > self.slot_bytes and cur:
                out.append(cur)
                cur, used = [], 0
            cur.append(i)
            used += need
        if cur:
            out.append(cur)
        return out

    def retire(self, futures):
        """Futures (writer jobs) that still read the views returned by the last process() call."""
        self._guards[self._k ^ 1] = list(futures)

    def _pinned_stage(self, nbytes):
        if self._stage is None or self._stage.numel() < nbytes:
            self._stage = torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)
        return self._stage

    def process(self, items):
        k = self._k
        for fut in self._guards[k]:
            fut.result()
        self._guards[k] = []
        C = self.plan.n_channels
        slot, dev = self._slots[k], self._dev
        stream = torch.cuda.current_stream(self.plan.device)
        views = []
        if self.mode == "filterbank":
            waves = [_as_wave(w) for w in items]
            if len({w.dtype for w in waves}) > 1:
                waves = [w.astype(np.float64) for w in waves]
            lengths = np.asarray([w.shape[0] for w in waves], dtype=np.int64)
            total = int(lengths.sum())
            if total == 0:
                z = np.zeros((C, 0), dtype=self.dtype)
                return [(z, z.copy()) if self.with_env else z for _ in waves]
            wdt = torch.from_numpy(waves[0][:0]).dtype
            stage = self._pinned_stage(total * waves[0].dtype.itemsize).view(wdt)[:total]
            np.concatenate(waves, out=stage.numpy())
            wave_dev = stage.to(self.plan.device, non_blocking=True)
            nval = C * total
            out = {"gfb": dev[:nval * self.dtype.itemsize].view(self.tdt)}
            if self.with_env:
                out["env"] = dev[nval * self.dtype.itemsize:2 * nval * self.dtype.itemsize].view(self.tdt)
            self.plan.batch(lengths).run(wave_dev, lpf=self.lpf, cutoff=self.cutoff, out=out, gfb=self.tdt,
                                         env=self.tdt if self.with_env else None)
            nbytes = nval * self.per_value
            slot[:nbytes].copy_(dev[:nbytes], non_blocking=True)
            stream.synchronize()
            host = slot.numpy()
            off = 0
            for n in lengths:
                sz = C * int(n) * self.dtype.itemsize
                g = host[off:off + sz].view(self.dtype).reshape(C, int(n))
                if self.with_env:
                    e0 = nval * self.dtype.itemsize + off
                    views.append((g, host[e0:e0 + sz].view(self.dtype).reshape(C, int(n))))
                else:
                    views.append(g)
                off += sz
        else:
            off = 0
            placed = []
            for m in items:
                m = np.asarray(m)
                if m.ndim != 2:
                    raise ValueError("matrix must be two-dimensional (channels x samples)")
                if m.dtype not in (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.int16)):
                    m = m.astype(np.float64)
                rows, n = m.shape
                if n == 0:
                    raise ValueError("N must be positive.")
                src = self._pinned_stage(m.nbytes)[:m.nbytes]
                np.copyto(src.numpy().view(m.dtype).reshape(m.shape), m)
                m_dev = src.to(self.plan.device, non_blocking=True).view(torch.from_numpy(m[:0, :0]).dtype).view(rows, n)
                res = self.plan.envelope_rows(m_dev, self.lpf, self.cutoff, out_dtype=self.tdt)
                sz = rows * n * self.dtype.itemsize
                slot[off:off + sz].copy_(res.view(-1).view(torch.uint8), non_blocking=True)
                stream.synchronize()   # the staging buffer is reused by the next matrix
                placed.append((off, sz, rows, n))
                off += sz
            host = slot.numpy()
            views = [host[o:o + sz].view(self.dtype).reshape(r, n) for o, sz, r, n in placed]
        self._k ^= 1
        return views


def _rows_op(matrix, op, lpf, cutoff):
    m = np.asarray(matrix)
    if m.dtype not in (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.int16)):
        m = m.astype(np.float64)
    m = np.ascontiguousarray(m)
    if m.shape[1] == 0:
        raise ValueError("N must be positive.")
    plan = engine.any_plan()
    out = plan.envelope_rows(_to_device(m, plan.device), lpf, cutoff, out_dtype=torch.float64, op=op)
    return _to_host(out)


def hilbert_imag_rows(matrix):
    """Imaginary part of paddedHilbert(row) for every row (EnvelopeExtraction.py:20-36)."""
    return _rows_op(matrix, 1, False, 100)


def lowpass_rows(matrix, freq):
    """lowPassFilter(row, freq) for every row (EnvelopeExtraction.py:39-48)."""
    return _rows_op(matrix, 2, True, freq)


def gather_windows_from_matrix(envelopes, timepoints, radius=5, step=160):
    """InputGenerator.py:73-80 on a loaded (C, n) envelope matrix -> (m, 2R+1, C) float32."""
    env = np.ascontiguousarray(envelopes)
    if env.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        env = env.astype(np.float64)
    C, n = env.shape
    idx = window_indices(n, timepoints, radius, step)
    plan = engine.any_plan()
    out = engine.gather_windows_cn(_to_device(env, plan.device),
                                   _to_device(np.ascontiguousarray(idx.reshape(-1)), plan.device))
    return _to_host(out).reshape(idx.shape[0], 2 * radius + 1, C)


def window_indices(n, centers, radius, step):
    """Sample indices read by InputGenerator.py:76 for one file, with Python list
    semantics: a negative index wraps once, anything else out of range is an IndexError."""
    centers = np.asarray(centers, dtype=np.int64).reshape(-1)
    offs = step * (np.arange(2 * radius + 1, dtype=np.int64) - radius)
    idx = centers[:, None] + offs[None, :]
    idx = np.where(idx < 0, idx + n, idx)
    if idx.size and (idx.min() < 0 or idx.max() >= n):
        raise IndexError("index out of bounds")
    return idx


_pipelines = OrderedDict()   # corpus-sized requests: cached engine.WindowPipeline objects
_pipelines_lock = threading.Lock()
_MAX_PIPELINES = 2


def _pipeline_for(plan, lengths, dots, step, phase, LPF, CUTOFF, src_offsets=None, share=1):
    """The sub-batch layout, device buffers, pinned staging and worker pool of a corpus-sized request
    depend only on the utterance lengths and the grid: built once, reused by every later call with the
    same corpus (a second `prepare input --cutoff K` run, the next epoch's noise draw, a benchmark)."""
    key = (id(plan), hashlib.sha1(lengths.tobytes()).hexdigest(), int(dots), int(step), int(phase), int(share),
           None if src_offsets is None else hashlib.sha1(np.ascontiguousarray(src_offsets).tobytes()).hexdigest())
    with _pipelines_lock:
        pipe = _pipelines.get(key)
        if pipe is None:
            # `share` processes of this box run a pipeline each: split the host cores between their pools
            # (one core stays free for the CUDA host-callback thread that hands frames to the pool)
            threads = max(1, engine.host_cores() // max(int(share), 1) - (1 if int(share) == 1 else 0))
            if os.environ.get("F2CNN_B200_PLACER_THREADS"):   # development knob
                threads = max(1, int(os.environ["F2CNN_B200_PLACER_THREADS"]))
            pipe = engine.WindowPipeline(plan, lengths, dots=dots, step=step, phase=phase, lpf=LPF, cutoff=CUTOFF,
                                         src_offsets=src_offsets, placer=engine.Placer(threads))
            _pipelines[key] = pipe
            while len(_pipelines) > _MAX_PIPELINES:
                _pipelines.popitem(last=False)
        else:
            _pipelines.move_to_end(key)
        pipe.lpf, pipe.cutoff = bool(LPF), CUTOFF
        return pipe


def release_cached_pipelines():
    """Drop the cached corpus pipelines (device buffers, pinned staging, worker threads) and the pooled
    output block."""
    with _pipelines_lock:
        _pipelines.clear()
    engine.release_host_pool()


def features_to_windows(waves, coefs, timepoints, LPF=False, CUTOFF=100, radius=5, step=160, device_out=False,
                        out=None, counts=None, shard=None):
    """Fused path from waveforms to the (N, 2R+1, C) float32 input tensor.

    waves: list of 1-D arrays (same dtype), or (flat, lengths) as returned by ingest.read_corpus
    (one buffer of all samples, pinned for full speed); timepoints: list of int arrays (window
    centres per utterance, in output order) -- or, with `counts`, ONE flat int array of all centres
    and counts[u] of them per utterance.  Equivalent to
    erb_filterbank -> ExtractEnvelopeFromMatrix(LPF, CUTOFF) -> the gather of
    InputGenerator.py:73-80 per utterance, rows concatenated in list order.  `out`: an existing
    C-contiguous (N, 2R+1, C) float32 numpy array to fill (a corpus-sized result is 7.5 GB: reusing
    it saves its page faults); by default a fresh array is returned.
    shard=(rank, world): one of `world` processes (one per GPU, e.g. under torchrun) that were all
    handed the SAME corpus and the SAME `out` -- a shared mapping, hostmem.SharedArray.  This call
    computes the utterances engine.shard_utterances deals to `rank` and writes their rows, and only
    theirs, at their final offsets; there is nothing to exchange afterwards."""
    coefs = np.asarray(coefs, dtype=np.float64)
    plan = engine.plan_for(coefs)
    C = plan.n_channels
    dots = 2 * radius + 1
    flat_in = None
    if isinstance(waves, tuple) and len(waves) == 2:
        flat_in, lengths = waves
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        if not torch.is_tensor(flat_in):
            flat_in = torch.from_numpy(_as_wave(flat_in))
        if flat_in.dim() != 1 or flat_in.numel() != int(lengths.sum()):
            raise ValueError("flat buffer holds %d samples, lengths add up to %d" % (flat_in.numel(), int(lengths.sum())))
        n_utts = int(lengths.shape[0])
    else:
        waves = [_as_wave(w) for w in waves]
        dts = {w.dtype for w in waves}
        if len(dts) > 1:
            waves = [w.astype(np.float64) for w in waves]
        lengths = np.asarray([w.shape[0] for w in waves], dtype=np.int64)
        n_utts = len(waves)
    if counts is not None:
        centers = np.ascontiguousarray(timepoints, dtype=np.int64).reshape(-1)
        counts = np.ascontiguousarray(counts, dtype=np.int64).reshape(-1)
        if counts.shape[0] != n_utts or int(counts.sum()) != centers.shape[0]:
            raise ValueError("counts: one entry per utterance, adding up to the number of timepoints")
    else:
        if len(timepoints) != n_utts:
            raise ValueError("%d utterances, %d timepoint lists" % (n_utts, len(timepoints)))
        counts = np.fromiter((len(t) for t in timepoints), dtype=np.int64, count=n_utts)
        centers = (np.concatenate([np.asarray(t, dtype=np.int64).reshape(-1) for t in timepoints])
                   if n_utts else np.zeros(0, dtype=np.int64))
    total = int(counts.sum())
    if out is not None:
        if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and
                out.shape == (total, dots, C)):
            raise ValueError("out must be a C-contiguous float32 array of shape %s" % ((total, dots, C),))
        if device_out:
            raise ValueError("out= is a host array; it cannot be combined with device_out")
    if shard is not None:
        rank, world = int(shard[0]), int(shard[1])
        if out is None or device_out or not 0 <= rank < world:
            raise ValueError("shard=(rank, world) needs 0 <= rank < world and the shared `out` array")
    if total == 0:
        return np.zeros((0, dots, C), dtype=np.float32) if out is None else out

    if shard is not None or (not device_out and total * dots * C * 4 >= _PIPELINE_BYTES and n_utts >= 8):
        # corpus-sized request: decimated frames over PCIe, rows placed on the host (engine.WindowPipeline)
        first = int(centers[0]) - radius * step
        phase = first % step if first >= 0 else 0
        sel, src_offsets, row_offsets, share = slice(None), None, None, 1
        cen, lens, cnts = centers, lengths, counts
        if shard is not None:
            sel = engine.shard_utterances(lengths, world)[rank]
            mine = np.zeros(n_utts, dtype=bool)
            mine[sel] = True
            cen, lens, cnts = centers[np.repeat(mine, counts)], lengths[sel], counts[sel]
            src_offsets = (np.cumsum(lengths) - lengths)[sel]
            row_offsets = (np.cumsum(counts) - counts)[sel]
            share = world
        # cheap look at the first utterances before anything is built: a request that is not on one grid
        # goes straight to the general path
        head = int(min(8, cnts.shape[0]))
        n_dec_h = np.where(lens[:head] > phase, (lens[:head] - phase + step - 1) // step, 0)
        probe, phase_h, _ = engine.window_runs(cen[:int(cnts[:head].sum())], cnts[:head], lens[:head],
                                               np.concatenate([[0], np.cumsum(n_dec_h)]), radius, step, phase)
        on_one_grid = probe is not None and phase_h == phase
    else:
        on_one_grid = False
    if on_one_grid:
        pipe = _pipeline_for(plan, lens, dots, step, phase, LPF, CUTOFF, src_offsets, share)
        result = engine.host_empty((total, dots, C), np.float32) if out is None else out
        state = {}

        def detect():
            # runs on the host while the first sub-batch is already on the device
            runs, phase2, _ = engine.window_runs(cen, cnts, lens, pipe.frame_offsets, radius, step, phase,
                                                 row_offsets=row_offsets)
            state["ok"] = runs is not None and phase2 == phase
            return runs if state["ok"] else np.zeros((0, 3), dtype=np.int64)

        if flat_in is not None:
            pipe.run(flat_in, detect, result)
        else:
            pipe.run(waves if shard is None else [waves[u] for u in sel], detect, result)
        if state["ok"]:
            return result
        # legal timepoints that are not windows of consecutive frames of one grid (a wrapping index, mixed
        # phases) further down the corpus: nothing was placed; the general path below answers
        del result
    if shard is not None:
        raise ValueError("shard= needs timepoints that are windows of consecutive frames of one decimated grid")

    # general path: one batch, windows gathered on the device
    cpos = np.concatenate([[0], np.cumsum(counts)])
    idx = [window_indices(int(n), centers[cpos[u]:cpos[u + 1]], radius, step) for u, n in enumerate(lengths)]
    # one decimated grid serves every window when all indices share a residue mod step
    allidx = np.concatenate([i.reshape(-1) for i in idx if i.size])
    phase = int(allidx[0] % step)
    on_grid = bool(np.all(allidx % step == phase))
    if flat_in is None:
        flat_in = torch.from_numpy(np.concatenate(waves) if len(waves) > 1 else waves[0])
    wave_dev = (flat_in if flat_in.is_pinned() or flat_in.numel() * flat_in.element_size() < (1 << 16)
                else flat_in.pin_memory()).to(plan.device, non_blocking=True)
    if on_grid:
        batch = plan.batch(lengths, step=step, phase=phase)
        res = batch.run(wave_dev, lpf=LPF, cutoff=CUTOFF, dec=True)
        bases, strided = [], True
        for u, i in enumerate(idx):
            if not i.size:
                continue
            rows = (i - phase) // step + batch.frame_offsets[u]
            strided = strided and bool(np.all(np.diff(rows, axis=1) == 1))
            bases.append(rows)
        rows = np.concatenate(bases)
        if strided:
            base_dev = _to_device(np.ascontiguousarray(rows[:, 0]), plan.device)
            res_w = engine.gather_windows(res["dec"], base_dev, dots, 1)
        else:
            res_w = engine.gather_index(res["dec"], _to_device(np.ascontiguousarray(rows.reshape(-1)), plan.device))
            res_w = res_w.view(total, dots, C)
    else:
        batch = plan.batch(lengths, step=step, phase=0)
        res = batch.run(wave_dev, lpf=LPF, cutoff=CUTOFF, env_t=True)
        rows = np.concatenate([i + batch.sample_offsets[u] for u, i in enumerate(idx) if i.size])
        res_w = engine.gather_index(res["env_t"], _to_device(np.ascontiguousarray(rows.reshape(-1)), plan.device))
        res_w = res_w.view(total, dots, C)
    if device_out:
        return res_w
    host = _to_host(res_w)
    if out is not None:
        out[...] = host
        return out
    return host


class FrameWindows:
    """The input tensor of a corpus WITHOUT its (2R+1)-fold redundancy: the decimated envelope frames of every
    utterance in one page-locked host array, and the windows as overlapping VIEWS of it.

    Row k of utterance u of `input_data.npy` is frames k .. k+dots-1 of that utterance
    (InputGenerator.py:73-80 on the label grid of LabelDataGenerator.py:38-50), i.e. dots*C consecutive floats
    starting at frame frame_offsets[u] + k: `windows(u)` is that (counts[u], dots, C) array as a strided view,
    `self[r]` is global row r, `materialize()` builds the reference's tensor (bit-identical to
    api.features_to_windows).  0.71 GB instead of 7.5 GB for the 4620-utterance corpus: a consumer that reads
    windows through the views (a training loop that gathers mini-batches) never pays for the expansion, which is
    what bounds the end-to-end time of the full tensor on any number of GPUs.

    `frames` belongs to the cached pipeline of this corpus layout: it is valid until the next
    features_to_frames / features_to_windows call with the same lengths (copy it to keep it)."""

    def __init__(self, frames, frame_offsets, counts, dots, utterances, placer):
        self.frames, self.frame_offsets, self.counts, self.dots = frames, frame_offsets, counts, int(dots)
        self.utterances = utterances          # indices into the caller's utterance list (a shard holds a subset)
        self.row_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self._placer = placer

    def __len__(self):
        return int(self.row_offsets[-1])

    def windows(self, u):
        """(counts[u], dots, C) float32 view of the u-th utterance held here (no copy)."""
        C = self.frames.shape[1]
        k, f0 = int(self.counts[u]), int(self.frame_offsets[u])
        if k == 0:
            return np.zeros((0, self.dots, C), dtype=np.float32)
        base = self.frames[f0:f0 + k + self.dots - 1]
        return np.lib.stride_tricks.as_strided(base, shape=(k, self.dots, C), strides=(C * 4, C * 4, 4), writeable=False)

    def __getitem__(self, r):
        r = int(r)
        if r < 0:
            r += len(self)
        if not 0 <= r < len(self):
            raise IndexError("row %d of %d" % (r, len(self)))
        u = int(np.searchsorted(self.row_offsets, r, side="right") - 1)
        f = int(self.frame_offsets[u]) + r - int(self.row_offsets[u])
        return self.frames[f:f + self.dots]

    def materialize(self, out=None):
        """The (N, dots, C) float32 tensor of the reference, rows placed by the library's host pool."""
        C = self.frames.shape[1]
        out = engine.host_empty((len(self), self.dots, C), np.float32) if out is None else out
        runs = np.stack([self.frame_offsets[:-1], self.row_offsets[:-1], self.counts], axis=1).astype(np.int64)
        runs = np.ascontiguousarray(runs[runs[:, 2] > 0])
        if runs.shape[0]:
            self._placer.submit(torch.from_numpy(self.frames), runs, out, dots=self.dots, stream=None, after_stream=False)
            self._placer.wait()
        return out


def features_to_frames(waves, coefs, LPF=False, CUTOFF=100, radius=5, step=160, counts=None, shard=None):
    """Waveforms -> FrameWindows: the same filterbank -> envelope -> decimation as features_to_windows on the
    label grid (window k of an utterance is centred on step*radius + step*k, LabelDataGenerator.py:38-50), but the
    call ends when the decimated frames are on the host -- nothing is expanded.  waves: (flat, lengths) or a list
    of 1-D arrays; counts[u]: windows of utterance u (default: the label grid's int(n/step - dots - 1));
    shard=(rank, world): this process computes the utterances engine.shard_utterances deals to `rank`
    (FrameWindows.utterances)."""
    coefs = np.asarray(coefs, dtype=np.float64)
    plan = engine.plan_for(coefs)
    dots = 2 * radius + 1
    flat_in = None
    if isinstance(waves, tuple) and len(waves) == 2:
        flat_in, lengths = waves
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        if not torch.is_tensor(flat_in):
            flat_in = torch.from_numpy(_as_wave(flat_in))
        if flat_in.dim() != 1 or flat_in.numel() != int(lengths.sum()):
            raise ValueError("flat buffer holds %d samples, lengths add up to %d" % (flat_in.numel(), int(lengths.sum())))
    else:
        waves = [_as_wave(w) for w in waves]
        if len({w.dtype for w in waves}) > 1:
            waves = [w.astype(np.float64) for w in waves]
        lengths = np.asarray([w.shape[0] for w in waves], dtype=np.int64)
    n_frames = (lengths + step - 1) // step
    grid = np.maximum((lengths / step - dots - 1).astype(np.int64), 0)
    if counts is None:
        counts = grid
    else:
        counts = np.ascontiguousarray(counts, dtype=np.int64).reshape(-1)
        if counts.shape[0] != lengths.shape[0] or np.any(counts < 0) or np.any(counts + dots - 1 > np.maximum(n_frames, dots - 1)):
            raise IndexError("counts: one entry per utterance, windows must lie inside the utterance's frames")
    sel, src_offsets, share = np.arange(lengths.shape[0]), None, 1
    if shard is not None:
        rank, world = int(shard[0]), int(shard[1])
        if not 0 <= rank < world:
            raise ValueError("shard=(rank, world) needs 0 <= rank < world")
        sel = engine.shard_utterances(lengths, world)[rank]
        src_offsets = (np.cumsum(lengths) - lengths)[sel]
        share = world
    pipe = _pipeline_for(plan, lengths[sel], dots, step, 0, LPF, CUTOFF, src_offsets, share)
    no_runs = np.zeros((0, 3), dtype=np.int64)
    if flat_in is not None:
        pipe.run(flat_in, no_runs, None)
    else:
        pipe.run([waves[u] for u in sel], no_runs, None)
    return FrameWindows(pipe._dec_host.numpy(), pipe.frame_offsets, counts[sel], dots, sel, pipe.placer)


def dense_frames(wave, coefs, LPF=False, CUTOFF=100, radius=5, step=160, normalize=True, dtype=np.float64,
                 frames=None):
    """In-memory front end of Evaluating.EvaluateOneWavArray (Evaluating.py:52-80):
    filterbank -> envelope -> dense stride-1 framing (-> normalizeInput per frame).
    Returns (nb, 2R+1, C); `frames=(i0, i1)` restricts the frame range."""
    w = _as_wave(wave)
    coefs = np.asarray(coefs, dtype=np.float64)
    plan = engine.plan_for(coefs)
    n = int(w.shape[0])
    dots = 2 * radius + 1
    nb = int(n - dots * step)
    i0, i1 = (0, nb) if frames is None else (int(frames[0]), int(frames[1]))
    if frames is not None and not (0 <= i0 <= i1 <= max(nb, 0)):
        raise IndexError("frames=(%d, %d) outside the %d frames of this utterance" % (i0, i1, max(nb, 0)))
    if nb <= 0 or i1 <= i0:
        return np.zeros((0, dots, plan.n_channels), dtype=dtype)
    batch = plan.batch([n], step=step)
    res = batch.run(_to_device(w, plan.device), lpf=LPF, cutoff=CUTOFF, env_t=True)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    out, flag = engine.dense_frames(res["env_t"], dots, step, i0, i1, normalize=normalize, out_dtype=tdt)
    host = _to_host(out)
    if normalize and int(flag.item()) != 0:
        raise ValueError("values must all be positive")  # Training.py:18-20
    return host


def evaluate_front_end(wave, coefs, LPF=False, CUTOFF=100, radius=5, step=160, frames=True, dtype=np.float64):
    """ONE filterbank pass for `cnn eval*` (Evaluating.py:52-80): the (C, n) envelopes the figure shows
    and the (nb, 2R+1, C) normalised frames model.predict takes, both from the same time-major envelope.
    frames=False returns (envelopes, device env_t) for a consumer on the device (api.cnn_evaluate)."""
    w = _as_wave(wave)
    coefs = np.asarray(coefs, dtype=np.float64)
    plan = engine.plan_for(coefs)
    n = int(w.shape[0])
    dots = 2 * radius + 1
    if n == 0:
        raise ValueError("N must be positive.")
    batch = plan.batch([n], step=step)
    res = batch.run(_to_device(w, plan.device), lpf=LPF, cutoff=CUTOFF, env=torch.float64, env_t=True)
    envelopes = _to_host(res["env"]).reshape(plan.n_channels, n)
    if not frames:
        return envelopes, res["env_t"]
    nb = int(n - dots * step)
    if nb <= 0:
        return envelopes, np.zeros((0, dots, plan.n_channels), dtype=dtype)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    out, flag = engine.dense_frames(res["env_t"], dots, step, 0, nb, normalize=True, out_dtype=tdt)
    host = _to_host(out)
    if int(flag.item()) != 0:
        raise ValueError("values must all be positive")  # Training.py:18-20
    return envelopes, host


_networks = OrderedDict()


def _network_for(weights):
    """Tensor-core network for a list of Keras-layout arrays (or an F2CNN module), cached by content."""
    from . import cnn
    arrays = cnn.keras_arrays(weights) if isinstance(weights, torch.nn.Module) else [np.asarray(a) for a in weights]
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float32).tobytes())
    key = (h.hexdigest(), torch.cuda.current_device())
    net = _networks.get(key)
    if net is None:
        net = cnn.TensorCoreCNN(arrays)
        _networks[key] = net
        while len(_networks) > 4:
            _networks.popitem(last=False)
    return net


def cnn_evaluate(wave, coefs, weights, LPF=False, CUTOFF=100, radius=5, step=160, with_envelopes=True):
    """`cnn eval*` from waveform to scores on the device (Evaluating.py:52-87): filterbank -> envelope ->
    every stride-1 frame normalised and run through the reference network on the tensor cores
    (csrc/f2_cnn.cu), without materialising the (nb, 11, 128) frame tensor.  weights: the 12 arrays of
    Keras' model.get_weights() or an f2cnn_b200.cnn.F2CNN module.  Returns (scores (nb, 2) float32,
    envelopes (C, n) float64 or None).  Only the configured geometry (RADIUS = 5, 128 channels)."""
    coefs = np.asarray(coefs, dtype=np.float64)
    if radius != 5 or coefs.shape[0] != 128:
        raise ValueError("the tensor-core network is built for RADIUS = 5 and 128 channels")
    net = _network_for(weights)
    if with_envelopes:
        envelopes, env_t = evaluate_front_end(wave, coefs, LPF, CUTOFF, radius, step, frames=False)
    else:
        w = _as_wave(wave)
        plan = engine.plan_for(coefs)
        envelopes = None
        env_t = plan.batch([int(w.shape[0])], step=step).run(_to_device(w, plan.device), lpf=LPF, cutoff=CUTOFF,
                                                             env_t=True)["env_t"]
    scores = net.predict_envelope(env_t, step)
    return _to_host(scores), envelopes


def label_fit(tracks, firsts, centers, radius=5, step=160):
    """Slope labels for a whole corpus in one launch (LabelDataGenerator.py:60-68).

    tracks[u]: float64 formant track of utterance u (Hz per 10 ms frame); firsts[u]: int array,
    index in tracks[u] of the first of the 2*radius+1 frames of each timepoint; centers[u]: the
    timepoints in samples.  Returns a float64 (N, 4) host array of (slope, intercept, r, p), rows in
    utterance order then timepoint order."""
    plan = engine.any_plan()
    dots = 2 * int(radius) + 1
    offs = np.concatenate([[0], np.cumsum([len(t) for t in tracks])]).astype(np.int64)
    first_parts, center_parts = [], []
    for u, (f, c) in enumerate(zip(firsts, centers)):
        f = np.asarray(f, dtype=np.int64)
        c = np.asarray(c, dtype=np.int64)
        if f.shape != c.shape:
            raise ValueError("firsts[%d] and centers[%d] differ in length" % (u, u))
        if f.size and (f.min() < 0 or f.max() + dots > len(tracks[u])):
            raise IndexError("utterance %d: a window leaves the formant track" % u)
        first_parts.append(f + offs[u])
        center_parts.append(c)
    first = np.concatenate(first_parts + [np.zeros(0, np.int64)])
    center = np.concatenate(center_parts + [np.zeros(0, np.int64)])
    if first.size == 0:
        return np.zeros((0, 4))
    if center.max() > np.iinfo(np.int32).max:
        raise OverflowError("timepoint beyond int32")
    flat = np.concatenate([np.asarray(t, dtype=np.float64) for t in tracks])
    out = engine.label_fit(_to_device(flat, plan.device), _to_device(first, plan.device),
                           _to_device(center.astype(np.int32), plan.device), dots, step)
    return _to_host(out)
