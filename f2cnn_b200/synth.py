"""Seeded synthetic 16 kHz inputs shared by the golden-vector
generator, the tests and bench.py.  There is no TIMIT on the box, so every workload
is generated; white noise is the worst case for the envelope's edge terms."""
import numpy as np

FS = 16000


def white_noise_i16(n, seed=0, sigma=3000.0):
    rng = np.random.default_rng(seed)
    return np.clip(np.round(sigma * rng.standard_normal(n)), -32768, 32767).astype(np.int16)


def delta_i16(n, amp=10000, at=0):
    x = np.zeros(n, dtype=np.int16)
    if n > at:
        x[at] = amp
    return x


def tone_i16(n, freq=1000.0, amp=8000.0, fs=FS):
    t = np.arange(n) / fs
    return np.round(amp * np.sin(2 * np.pi * freq * t)).astype(np.int16)


def chirp_i16(n, f0=100.0, f1=7000.0, amp=8000.0, fs=FS):
    t = np.arange(n) / fs
    dur = max(n, 1) / fs
    phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
    return np.round(amp * np.sin(phase)).astype(np.int16)


def speech_like_i16(n, seed=7, fs=FS):
    """Noise shaped by Hann bursts with silent ends -- speech-like energy contour."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n)
    env = np.zeros(n)
    pos = int(0.05 * n)
    while pos < int(0.9 * n):
        ln = int(rng.integers(int(0.03 * fs), int(0.25 * fs)))
        ln = min(ln, int(0.95 * n) - pos)
        if ln <= 8:
            break
        env[pos:pos + ln] += np.hanning(ln) * rng.uniform(0.2, 1.0)
        pos += ln + int(rng.integers(0, int(0.05 * fs) + 1))
    return np.clip(np.round(6000.0 * x * env), -32768, 32767).astype(np.int16)


def label_grid(n, radius=5, step=160):
    """Timepoint grid of LabelDataGenerator.ExtractLabel (reference
    scripts/processing/LabelDataGenerator.py:44-50) with every step kept:
    START + k*STEP for k < int(n/STEP - (2R+1) - 1)."""
    nb = int(n / step - (2 * radius + 1) - 1)
    return np.asarray([step * radius + k * step for k in range(max(nb, 0))], dtype=np.int64)


def corpus_lengths(n_utts=4620, lo=32000, hi=64000, seed=1):
    """TIMIT-TRAIN-sized corpus (config 2): utterance lengths ~ U(lo, hi)."""
    rng = np.random.default_rng(seed)
    return rng.integers(lo, hi + 1, size=n_utts).astype(np.int64)


def corpus_waves_i16(lengths, seed=1, sigma=3000.0):
    """One flat int16 buffer + offsets for the whole corpus."""
    lengths = np.asarray(lengths, dtype=np.int64)
    offsets = np.zeros(lengths.shape[0] + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    rng = np.random.default_rng(seed + 1000)
    total = int(offsets[-1])
    flat = np.empty(total, dtype=np.int16)
    blk = 1 << 22
    for s in range(0, total, blk):
        e = min(total, s + blk)
        flat[s:e] = np.clip(np.round(sigma * rng.standard_normal(e - s, dtype=np.float32)), -32768, 32767).astype(
            np.int16)
    return flat, offsets
