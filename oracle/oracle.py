"""ctypes wrapper around oracle/libf2oracle.so -- TEST INFRASTRUCTURE ONLY.

The float64 CPU restatement of the reference hot path (see f2_oracle.c).  Importable
only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.  The product package f2cnn_b200 never imports this module.

Function names follow the reference symbols they restate (file:line relative to
/root/reference):
  centre_freqs / make_erb_filters / erb_filterbank   gammatone/filters.py:74,89,195
  padded_hilbert / low_pass_filter / extract_envelope scripts/processing/EnvelopeExtraction.py:20,39,51
  gather_windows                                      scripts/processing/InputGenerator.py:73-80
  dense_frames / normalize_input                      scripts/CNN/Evaluating.py:70-78, scripts/CNN/Training.py:13-28
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libf2oracle.so")
_lib = None

_d = ctypes.POINTER(ctypes.c_double)
_f = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C oracle in place (gcc only; a few hundred ms)."""
    src = os.path.join(_HERE, "f2_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libf2oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.f2o_erb_space.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, _d]
        L.f2o_centre_freqs.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_double, _d]
        L.f2o_make_erb_filters.argtypes = [ctypes.c_double, _d, ctypes.c_int, ctypes.c_double, _d]
        L.f2o_lfilter3.argtypes = [_d, _d, _d, _d, ctypes.c_int64]
        L.f2o_erb_filterbank.argtypes = [_d, ctypes.c_int64, _d, ctypes.c_int, _d]
        L.f2o_next_pow2.argtypes = [ctypes.c_int64]
        L.f2o_next_pow2.restype = ctypes.c_int64
        L.f2o_padded_hilbert.argtypes = [_d, ctypes.c_int64, _d, _d]
        L.f2o_butter1_lowpass.argtypes = [ctypes.c_double, _d, _d]
        L.f2o_low_pass_filter.argtypes = [_d, ctypes.c_int64, ctypes.c_double, _d]
        L.f2o_extract_envelope.argtypes = [_d, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_double, _d]
        L.f2o_gather_windows.argtypes = [_d, ctypes.c_int, ctypes.c_int64, _i64, ctypes.c_int64, ctypes.c_int,
                                         ctypes.c_int64, _f]
        L.f2o_gather_windows.restype = ctypes.c_int64
        L.f2o_dense_frames.argtypes = [_d, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_int64, _d]
        L.f2o_normalize_input.argtypes = [_d, ctypes.c_int64]
        L.f2o_normalize_input.restype = ctypes.c_int
        L.f2o_utterance.argtypes = [_d, ctypes.c_int64, _d, ctypes.c_int, ctypes.c_int, ctypes.c_double, _i64,
                                    ctypes.c_int64, ctypes.c_int, ctypes.c_int64, _d, _d, _f]
        L.f2o_utterance.restype = ctypes.c_int64
        L.f2o_num_threads.restype = ctypes.c_int
        L.f2o_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_d)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads():
    return lib().f2o_num_threads()


def set_num_threads(n):
    lib().f2o_set_num_threads(int(n))


def erb_space(low_freq, high_freq, num):
    out = np.empty(int(num), dtype=np.float64)
    lib().f2o_erb_space(float(low_freq), float(high_freq), int(num), _dp(out))
    return out


def centre_freqs(fs, num_freqs, cutoff):
    out = np.empty(int(num_freqs), dtype=np.float64)
    lib().f2o_centre_freqs(float(fs), int(num_freqs), float(cutoff), _dp(out))
    return out


def make_erb_filters(fs, cfs, width=1.0):
    cfs = _c64(cfs)
    out = np.empty((cfs.shape[0], 10), dtype=np.float64)
    lib().f2o_make_erb_filters(float(fs), _dp(cfs), cfs.shape[0], float(width), _dp(out))
    return out


def erb_filterbank(wave, coefs):
    wave = _c64(wave)
    coefs = _c64(coefs)
    out = np.empty((coefs.shape[0], wave.shape[0]), dtype=np.float64)
    lib().f2o_erb_filterbank(_dp(wave), wave.shape[0], _dp(coefs), coefs.shape[0], _dp(out))
    return out


def padded_hilbert(signal):
    signal = _c64(signal)
    re = np.empty_like(signal)
    im = np.empty_like(signal)
    lib().f2o_padded_hilbert(_dp(signal), signal.shape[0], _dp(re), _dp(im))
    return re + 1j * im


def butter1_lowpass(Wn):
    b = np.empty(2)
    a = np.empty(2)
    lib().f2o_butter1_lowpass(float(Wn), _dp(b), _dp(a))
    return b, a


def low_pass_filter(signal, freq):
    signal = _c64(signal)
    out = np.empty_like(signal)
    lib().f2o_low_pass_filter(_dp(signal), signal.shape[0], float(freq), _dp(out))
    return out


def extract_envelope(matrix, LPF=False, CUTOFF=100):
    matrix = _c64(matrix)
    out = np.empty_like(matrix)
    C, n = matrix.shape
    lib().f2o_extract_envelope(_dp(matrix), C, n, int(bool(LPF)), float(CUTOFF), _dp(out))
    return out


def gather_windows(env, centers, radius=5, step=160):
    env = _c64(env)
    centers = np.ascontiguousarray(centers, dtype=np.int64)
    C, n = env.shape
    out = np.zeros((centers.shape[0], 2 * radius + 1, C), dtype=np.float32)
    bad = lib().f2o_gather_windows(_dp(env), C, n, centers.ctypes.data_as(_i64), centers.shape[0], int(radius),
                                   int(step), out.ctypes.data_as(_f))
    if bad:
        raise IndexError("index out of bounds in window %d" % (bad - 1))
    return out


def dense_frames(env, radius=5, step=160, i0=0, i1=None):
    env = _c64(env)
    C, n = env.shape
    nb = int(n - (2 * radius + 1) * step)
    if i1 is None:
        i1 = nb
    out = np.empty((max(i1 - i0, 0), 2 * radius + 1, C), dtype=np.float64)
    if i1 > i0:
        lib().f2o_dense_frames(_dp(env), C, n, int(radius), int(step), int(i0), int(i1), _dp(out))
    return out


def normalize_input(frame):
    out = _c64(frame).copy()
    rc = lib().f2o_normalize_input(_dp(out), out.size)
    if rc == 1:
        raise ValueError("values must all be positive")
    if rc == 2:
        raise ValueError("minvalue must be less than or equal to maxvalue")
    return out


def utterance(wave, coefs, LPF, CUTOFF, centers=None, radius=5, step=160):
    """filterbank -> envelope -> windows for one utterance; returns (gfb, env, windows)."""
    wave = _c64(wave)
    coefs = _c64(coefs)
    C, n = coefs.shape[0], wave.shape[0]
    gfb = np.empty((C, n))
    env = np.empty((C, n))
    if centers is None:
        centers = np.zeros(0, dtype=np.int64)
    centers = np.ascontiguousarray(centers, dtype=np.int64)
    win = np.zeros((centers.shape[0], 2 * radius + 1, C), dtype=np.float32)
    bad = lib().f2o_utterance(_dp(wave), n, _dp(coefs), C, int(bool(LPF)), float(CUTOFF),
                              centers.ctypes.data_as(_i64), centers.shape[0], int(radius), int(step), _dp(gfb),
                              _dp(env), win.ctypes.data_as(_f))
    if bad:
        raise IndexError("index out of bounds in window %d" % (bad - 1))
    return gfb, env, win


# ---- label generation: numpy/scipy restatement --------------------
SILENTS = ('pau', 'epi', 'h#')  # PHNFileReader.py:17


def label_fit(values, x):
    """LabelDataGenerator.py:62-68 for one timepoint, with the reference's own library calls:
    (a, b) = lstsq([x, 1], values); (r, p) = pearsonr(values, a*x + b)."""
    import warnings
    from scipy.stats import pearsonr
    values = np.asarray(values, dtype=np.float64)
    x = np.asarray(x)
    A = np.vstack([x, np.ones(len(x))]).T
    (a, b), _, _, _ = np.linalg.lstsq(A, values, rcond=None)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r, p = pearsonr(values, a * x + b)
    return float(a), float(b), float(r), float(p)


def phoneme_at(segments, t):
    """PHNFileReader.py:33-37: first (phoneme, first, last) whose closed interval holds t."""
    for name, first, last in segments:
        if first <= t <= last:
            return name
    return 'h#'


def extract_label(track_hz, segments, n_samples, ident, framerate=16000, radius=5, risk=0.05, period=10000):
    """LabelDataGenerator.py:22-77 from in-memory inputs: track_hz = the chosen formant column in
    Hz, segments = [(phoneme, first, last)], ident = [TEST|TRAIN, region, speaker, sentence]."""
    dots = 2 * radius + 1
    to_formant = framerate * period * (1.0 / 1000000)
    nb = int(n_samples / to_formant - dots - 1)
    step = int(to_formant)
    start = int(step * radius)
    rows = []
    for k in range(nb):
        t = start + k * step
        name = phoneme_at(segments, t)
        if name in SILENTS:
            continue
        if not name:
            break
        lo, hi = int(t / to_formant - radius), int(t / to_formant + radius) + 1  # FBFileReader.py:77-78
        if lo < 0 or hi >= len(track_hz):
            raise IndexError("window outside the formant track")  # the reference prints and exits
        x = np.array([t + (j - radius) * step for j in range(dots)])
        a, _, _, p = label_fit(track_hz[lo:hi], x)
        if p < risk:
            rows.append(list(ident) + [name, t, round(a, 5), round(p, 5), 1 if a > 0 else 0])
    return rows or None
