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


# ---- label-generation fixtures: VTR-style formant tracks and TIMIT-style phoneme segments -----
_PHONES = ["iy", "ae", "s", "n", "pau", "aa", "l", "epi", "uw", "t", "er", "m", "ow", "sh", "ih", "k"]


def formant_tracks_khz(n_frames, seed=0):
    """(n_frames, 8) float32 in kHz (F1..F4, B1..B4), one frame per 10 ms: smooth random walks with
    glides (clear slopes), plateaus (flat: no significant slope) and jitter, like a VTR .FB file."""
    rng = np.random.default_rng(seed)
    base = np.array([0.5, 1.5, 2.5, 3.5, 0.08, 0.1, 0.12, 0.15])
    out = np.empty((n_frames, 8))
    for col in range(8):
        track = np.empty(n_frames)
        v = base[col]
        i = 0
        while i < n_frames:
            seg = int(rng.integers(6, 30))
            kind = rng.integers(0, 3)
            slope = 0.0 if kind == 0 else rng.normal(0, 0.02 * base[col])
            for j in range(i, min(n_frames, i + seg)):
                v = min(max(v + slope, 0.3 * base[col]), 2.0 * base[col])
                track[j] = v
            i += seg
        out[:, col] = track + rng.normal(0, 0.004 * base[col], n_frames)
    return out.astype(np.float32)


def phoneme_segments(n_samples, seed=0):
    """[(first sample, last sample, phoneme)] covering [0, n_samples): silence at both ends, TIMIT
    style (segment k ends where segment k+1 starts)."""
    rng = np.random.default_rng(seed)
    cuts = [0, int(rng.integers(1500, 2600))]
    while cuts[-1] < n_samples - 4000:
        cuts.append(cuts[-1] + int(rng.integers(600, 3200)))
    cuts.append(n_samples)
    segs = []
    for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        name = "h#" if k in (0, len(cuts) - 2) else _PHONES[int(rng.integers(0, len(_PHONES)))]
        segs.append((a, b, name))
    return segs


def write_fb(path, tracks_khz, samp_period=10000):
    """VTR .FB layout: big-endian int32 nFrame, int32 sampPeriod, int16 sampSize, int16 fileType,
    then 8 big-endian float32 per frame."""
    import struct
    tracks = np.asarray(tracks_khz, dtype=">f4")
    with open(path, "wb") as f:
        f.write(struct.pack(">iihh", tracks.shape[0], samp_period, 32, 9))
        f.write(tracks.tobytes())


def write_phn(path, segments):
    with open(path, "w") as f:
        for a, b, name in segments:
            f.write("%d %d %s\n" % (a, b, name))
