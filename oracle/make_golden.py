#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs only in the build container, where /root/reference exists: it imports the
reference's own modules (gammatone.filters, scripts.processing.*, scripts.CNN.Training)
with stub `sphfile` / `matplotlib` modules (the only missing imports on this path; the
stubs are never called), feeds them the seeded synthetic inputs of SURVEY.md section 8d
and stores inputs + reference outputs.  Full (C,n) float64 matrices are 49 MB each, so
for the 3 s cases only a fixed set of time indices (both edges + interior) is kept for
every channel, plus per-channel RMS and sums; small cases are stored whole.

    python oracle/make_golden.py            # rewrites tests/golden/
    python oracle/make_golden.py --only-labels   # only tests/golden/labels.npz (section L)

The reference has no tests of its own, so these fixtures are what
pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
import csv
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("F2CNN_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree %s not present: golden vectors can only be regenerated in the "
                         "build container" % REF)
    sph = types.ModuleType("sphfile")
    sph.SPHFile = object
    sys.modules.setdefault("sphfile", sph)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    sys.path.insert(0, REF)
    from gammatone import filters
    from scripts.processing import EnvelopeExtraction, GammatoneFiltering, InputGenerator
    from scripts.CNN import Training
    return filters, GammatoneFiltering, EnvelopeExtraction, InputGenerator, Training


def sample_indices(n, count, seed):
    """Edges (first/last 48) + seeded interior points, sorted & unique."""
    edge = min(48, n // 2)
    idx = set(range(edge)) | set(range(n - edge, n))
    rng = np.random.default_rng(seed)
    if n > 2 * edge:
        idx |= set(int(v) for v in rng.integers(edge, n - edge, size=max(count - 2 * edge, 0)))
    return np.asarray(sorted(idx), dtype=np.int64)


def rms(a):
    return np.sqrt(np.mean(np.square(a), axis=-1))


LABEL_UTTS = [("TRAIN", "DR1", "FAAA0", "SX1", 30000, 11), ("TRAIN", "DR3", "MBBB0", "SI22", 47000, 12),
              ("TEST", "DR2", "FCCC0", "SA1", 52000, 13), ("TEST", "DR8", "MDDD0", "SX300", 39000, 14)]
LABEL_CONF = ("[FILTERBANK]\nFRAMERATE=16000\nNCHANNELS=8\nLOW_FREQ=100\n"
              "[CNN]\nFORMANT=2\nCENTERED=True\nRADIUS=5\nBATCH_SIZE=32\nEPOCHS=20\nRISK=0.05\n"
              "SAMPLING_PERIOD=10000\n")


def write_label_tree(synth, root):
    """The synthetic resources/f2cnn tree of the label fixtures (also used by the tests)."""
    from scipy.io import wavfile
    made = {}
    for (tt, dr, spk, sent, n, seed) in LABEL_UTTS:
        d = os.path.join(root, "resources", "f2cnn", tt)
        os.makedirs(d, exist_ok=True)
        stem = os.path.join(d, "%s.%s.%s" % (dr, spk, sent))
        wavfile.write(stem + ".WAV", 16000, synth.white_noise_i16(n, seed=seed))
        tracks = synth.formant_tracks_khz(n // 160 + 3, seed=seed)
        segs = synth.phoneme_segments(n, seed=seed)
        synth.write_fb(stem + ".FB", tracks)
        synth.write_phn(stem + ".PHN", segs)
        made["%s_%s_%s_%s" % (tt, dr, spk, sent)] = (tracks, segs, n)
    return made


def labels_section(synth):
    """L. LabelDataGenerator end to end (reference GenerateLabelData / ExtractLabel, files on disk)."""
    from configparser import ConfigParser
    from scripts.processing import LabelDataGenerator as LG
    from scripts.processing import FBFileReader as FB
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            with open("configF2CNN.conf", "w") as f:
                f.write(LABEL_CONF)
            made = write_label_tree(synth, tmp)
            LG.GenerateLabelData()
            with open(os.path.join("trainingData", "label_data.csv")) as f:
                csv_text = f.read()
            cfg = ConfigParser()
            cfg.read("configF2CNN.conf")
            one = LG.ExtractLabel(os.path.join("resources", "f2cnn", "TEST", "DR2.FCCC0.SA1.WAV"), cfg)
            fb_hz, _ = FB.ExtractFBFile(os.path.join("resources", "f2cnn", "TEST", "DR2.FCCC0.SA1.FB"))
            # raw (slope, p) of EVERY grid step of one file, before the phoneme / RISK filters
            from scipy.stats import pearsonr
            track = fb_hz[:, 1]
            raw = []
            for k in range(int(52000 / 160.0 - 11 - 1)):
                step = 800 + 160 * k
                vals = np.array(FB.GetFromantFrequenciesAround(track, step, 5, 160.0))
                x = np.array([step + (j - 5) * 160 for j in range(11)])
                A = np.vstack([x, np.ones(len(x))]).T
                [a, b], _, _, _ = np.linalg.lstsq(A, vals, rcond=None)
                r, p = pearsonr(vals, a * x + b)
                raw.append([step, a, b, r, p])
        finally:
            os.chdir(cwd)
    d = dict(csv=np.asarray(csv_text), one_file_rows=np.asarray(repr(one)), fb_hz=fb_hz, raw=np.asarray(raw))
    for k, (tracks, segs, n) in made.items():
        d["tracks_" + k] = tracks
        d["segs_" + k] = np.asarray(["%d %d %s" % s for s in segs])
        d["n_" + k] = np.int64(n)
    np.savez_compressed(os.path.join(OUT, "labels.npz"), **d)
    print("labels done:", csv_text.count("\n"), "rows")


def main():
    sys.path.insert(0, ROOT)
    from f2cnn_b200 import synth
    filters, GF, EE, IG, TR = import_reference()
    os.makedirs(OUT, exist_ok=True)
    if "--only-labels" in sys.argv:
        labels_section(synth)
        return

    # ---- A. coefficient design ------------------------------------------------------
    coef = {}
    for tag, (fs, C, low) in {"fs16000_c128_l100": (16000, 128, 100), "fs16000_c256_l100": (16000, 256, 100),
                              "fs16000_c8_l100": (16000, 8, 100), "fs8000_c32_l50": (8000, 32, 50),
                              "fs44100_c64_l20": (44100, 64, 20)}.items():
        cf = filters.centre_freqs(fs, C, low)
        coef["cf_" + tag] = cf
        coef["coefs_" + tag] = filters.make_erb_filters(fs, cf)
    coef["erb_space_default"] = filters.erb_space()
    coef["coefs_width2"] = filters.make_erb_filters(16000, filters.centre_freqs(16000, 16, 100), width=2.0)
    from scipy.signal import butter
    for f in (20, 50, 100, 400):
        b, a = butter(1, f / (16000 / 2), 'low')
        coef["butter_%d" % f] = np.concatenate([b, a])
    np.savez_compressed(os.path.join(OUT, "coefs.npz"), **coef)

    cf128 = filters.centre_freqs(16000, 128, 100)
    co128 = filters.make_erb_filters(16000, cf128)

    # ---- B/C. 3 s utterances, 128 channels: sampled columns -------------------------
    n = 48000
    cases = {
        "white": synth.white_noise_i16(n, seed=0),
        "delta": synth.delta_i16(n),
        "tone1k": synth.tone_i16(n),
        "chirp": synth.chirp_i16(n),
        "speech": synth.speech_like_i16(n),
    }
    for name, wave in cases.items():
        idx = sample_indices(n, 256 if name == "white" else 160, seed=11)
        gfb = GF.GetFilteredOutputFromArray(wave, co128)
        env50 = EE.ExtractEnvelopeFromMatrix(gfb, True, 50)
        envno = EE.ExtractEnvelopeFromMatrix(gfb, False)
        d = dict(wave=wave, idx=idx, gfb=gfb[:, idx], env_lpf50=env50[:, idx], env_nolpf=envno[:, idx],
                 gfb_rms=rms(gfb), env_lpf50_rms=rms(env50), env_nolpf_rms=rms(envno),
                 gfb_sum=gfb.sum(axis=1), env_lpf50_sum=env50.sum(axis=1), env_nolpf_sum=envno.sum(axis=1))
        if name == "white":
            env20 = EE.ExtractEnvelopeFromMatrix(gfb, True, 20)
            env100 = EE.ExtractEnvelopeFromMatrix(gfb, True)  # default CUTOFF=100
            d.update(env_lpf20=env20[:, idx], env_lpf100=env100[:, idx], env_lpf20_rms=rms(env20),
                     env_lpf100_rms=rms(env100))
            # decimated grid actually read by GenerateInputData for the full label grid
            dec = np.arange(0, n, 160)
            d.update(dec_idx=dec, env_lpf50_dec=env50[:, dec], env_nolpf_dec=envno[:, dec])
        np.savez_compressed(os.path.join(OUT, "utt3s_%s.npz" % name), **d)
        print("utt3s", name, "done")

    # ---- D. small / ragged lengths, 8 channels, stored whole ------------------------
    cf8 = filters.centre_freqs(16000, 8, 100)
    co8 = filters.make_erb_filters(16000, cf8)
    small = {}
    for nn in (1, 2, 3, 4, 5, 16, 17, 255, 256, 257, 1000, 4096, 4097):
        w = synth.white_noise_i16(nn, seed=100 + nn)
        gfb = GF.GetFilteredOutputFromArray(w, co8)
        small["wave_%d" % nn] = w
        small["gfb_%d" % nn] = gfb
        small["env_lpf50_%d" % nn] = EE.ExtractEnvelopeFromMatrix(gfb, True, 50)
        small["env_nolpf_%d" % nn] = EE.ExtractEnvelopeFromMatrix(gfb)
    small["coefs"] = co8
    np.savez_compressed(os.path.join(OUT, "small.npz"), **small)

    # ---- E. lengths at / just under a power of two (no padding), 16 channels --------
    cf16 = filters.centre_freqs(16000, 16, 100)
    co16 = filters.make_erb_filters(16000, cf16)
    pw = dict(coefs=co16)
    for nn in (65530, 65535, 65536, 65537):
        w = synth.white_noise_i16(nn, seed=200 + (nn % 97))
        idx = sample_indices(nn, 160, seed=12)
        gfb = GF.GetFilteredOutputFromArray(w, co16)
        e50 = EE.ExtractEnvelopeFromMatrix(gfb, True, 50)
        eno = EE.ExtractEnvelopeFromMatrix(gfb, False)
        pw.update({"wave_%d" % nn: w, "idx_%d" % nn: idx, "gfb_%d" % nn: gfb[:, idx],
                   "env_lpf50_%d" % nn: e50[:, idx], "env_nolpf_%d" % nn: eno[:, idx],
                   "gfb_rms_%d" % nn: rms(gfb), "env_lpf50_rms_%d" % nn: rms(e50), "env_nolpf_rms_%d" % nn: rms(eno)})
    np.savez_compressed(os.path.join(OUT, "pow2.npz"), **pw)
    print("pow2 done")

    # ---- F. 256 channels, float64 (noise-mixed) input: evalnoise-style --------------
    cf256 = filters.centre_freqs(16000, 256, 100)
    co256 = filters.make_erb_filters(16000, cf256)
    nn = 20000
    base = synth.speech_like_i16(nn, seed=3)
    rng = np.random.default_rng(3)
    true_rms = np.sqrt(np.mean(base.astype(np.float64) ** 2))
    wave64 = base + rng.normal(scale=true_rms / 10 ** (10 / 20.0), size=nn)  # true 10 dB SNR
    gfb = GF.GetFilteredOutputFromArray(wave64, co256)
    e50 = EE.ExtractEnvelopeFromMatrix(gfb, True, 50)
    idx = sample_indices(nn, 160, seed=13)
    # dense framing of Evaluating.py:70-81 restated on a handful of frames + the reference normalizeInput
    frames_i = np.asarray([0, 1, 159, 160, 5000, nn - 1761], dtype=np.int64)
    frames = np.stack([np.stack([e50[:, 800 + i + (k - 5) * 160] for k in range(11)]) for i in frames_i])
    frames_norm = np.stack([TR.normalizeInput(fr.copy()) for fr in frames])
    np.savez_compressed(os.path.join(OUT, "c256_f64.npz"), wave=wave64, idx=idx, gfb=gfb[:, idx],
                        env_lpf50=e50[:, idx], gfb_rms=rms(gfb), env_lpf50_rms=rms(e50), frames_i=frames_i,
                        frames=frames, frames_norm=frames_norm,
                        ref_rms_int16=np.float64(np.sqrt(np.mean(np.square(base)))))
    print("c256 done")

    # ---- I. InputGenerator end to end (reference GenerateInputData, files on disk) --
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            with open("configF2CNN.conf", "w") as f:
                f.write("[FILTERBANK]\nFRAMERATE=16000\nNCHANNELS=8\nLOW_FREQ=100\n"
                        "[CNN]\nFORMANT=2\nCENTERED=True\nRADIUS=5\nBATCH_SIZE=32\nEPOCHS=20\nRISK=0.05\n"
                        "SAMPLING_PERIOD=10000\n")
            os.makedirs("trainingData")
            # deliberately NOT file-sorted, to pin the sorted-key / CSV-order row contract
            utts = [("TRAIN", "DR2", "SPK1", "SX10", 9000), ("TEST", "DR1", "SPK0", "SA1", 7000),
                    ("TRAIN", "DR1", "SPK2", "SI99", 8000)]
            rows = []
            waves = {}
            for (tt, dr, spk, sent, nn) in utts:
                os.makedirs(os.path.join("resources", "f2cnn", tt), exist_ok=True)
                w = synth.white_noise_i16(nn, seed=nn)
                waves["%s_%s_%s_%s" % (tt, dr, spk, sent)] = w
                env = EE.ExtractEnvelopeFromMatrix(GF.GetFilteredOutputFromArray(w, co8), True, 50)
                np.save(os.path.join("resources", "f2cnn", tt, "%s.%s.%s.ENV1" % (dr, spk, sent)), env)
                grid = synth.label_grid(nn)
                keep = grid[::3][::-1] if sent == "SX10" else grid[1::4]  # one file in descending order
                for tp in keep:
                    rows.append([tt, dr, spk, sent, "aa", int(tp), 0.5, 0.01, 1])
            # interleave rows of different files
            rows = rows[::2] + rows[1::2]
            with open(os.path.join("trainingData", "label_data.csv"), "w") as f:
                wr = csv.writer(f, lineterminator="\n")
                for r in rows:
                    wr.writerow(r)
            IG.GenerateInputData(LPF=True, CUTOFF=50)
            inp = np.load(os.path.join("trainingData", "input_data_LPF50.npy"))
            with open(os.path.join("trainingData", "label_data.csv")) as f:
                csv_text = f.read()
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "inputgen.npz"), input_data=inp, csv=np.asarray(csv_text), coefs=co8,
                        **{"wave_" + k: v for k, v in waves.items()})
    print("inputgen done", inp.shape, inp.dtype)
    labels_section(synth)


if __name__ == "__main__":
    main()
