"""Label generation: the reference's LabelDataGenerator run in the
build container produced tests/golden/labels.npz (oracle/make_golden.py section L).  CPU tests pin
the numpy/scipy restatement and the host-side readers to it; the GPU tests hold the CUDA kernel
(f2_label_fit through the C ABI) and the GenerateLabelData / ExtractLabel drop-ins to it."""
import ast
import csv
import io
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, "oracle"))

CONF = ("[FILTERBANK]\nFRAMERATE=16000\nNCHANNELS=8\nLOW_FREQ=100\n"
        "[CNN]\nFORMANT=2\nCENTERED=True\nRADIUS=5\nBATCH_SIZE=32\nEPOCHS=20\nRISK=0.05\nSAMPLING_PERIOD=10000\n")
UTTS = [("TRAIN", "DR1", "FAAA0", "SX1"), ("TRAIN", "DR3", "MBBB0", "SI22"), ("TEST", "DR2", "FCCC0", "SA1"),
        ("TEST", "DR8", "MDDD0", "SX300")]


def golden_rows(g):
    return [r for r in csv.reader(io.StringIO(str(g["csv"])))]


def segments_of(g, key):
    out = []
    for line in g["segs_" + key]:
        a, b, name = str(line).split(" ")
        out.append((name, int(a), int(b)))
    return out


def assert_same_rows(got, want):
    """Row for row the reference's CSV.  One licence: the slope is a multiple of 1/1 760 000 (0.01 Hz
    formant grid, 11 abscissae 160 apart), so about 1 % of the slopes are EXACT ties at the fifth
    decimal; the reference rounds those by the round-off noise of its SVD least squares, the kernel by
    its closed form, and the two may land on either side (never the class column: a slope that
    small has p ~ 1 and is dropped)."""
    assert len(got) == len(want)
    ties = 0
    for a, b in zip(got, want):
        assert a[:6] == b[:6] and a[7:] == b[7:], (a, b)
        if a[6] != b[6]:
            sa, sb = float(a[6]), float(b[6])
            assert abs(abs(sa - sb) - 1e-5) < 1e-9, (a, b)
            m = (sa + sb) / 2 * 1760000          # the exact slope, in units of 1/1 760 000
            assert abs(m - round(m)) < 1e-3 and round(m) % 88 == 44, (a, b, m)
            ties += 1
    assert ties <= max(3, len(want) // 25)


@pytest.fixture()
def label_tree(tmp_path, monkeypatch):
    """The same synthetic resources/f2cnn tree the golden CSV was generated from."""
    import make_golden
    from f2cnn_b200 import synth
    monkeypatch.chdir(tmp_path)
    (tmp_path / "configF2CNN.conf").write_text(CONF)
    make_golden.write_label_tree(synth, str(tmp_path))
    return tmp_path


def test_fixture_tree_is_the_one_the_golden_was_made_from(label_tree):
    g = load_golden("labels.npz")
    from f2cnn_b200 import synth
    import make_golden
    for (tt, dr, spk, sent, n, seed) in make_golden.LABEL_UTTS:
        key = "%s_%s_%s_%s" % (tt, dr, spk, sent)
        assert int(g["n_" + key]) == n
        assert np.array_equal(g["tracks_" + key], synth.formant_tracks_khz(n // 160 + 3, seed=seed))
        assert [str(s) for s in g["segs_" + key]] == ["%d %d %s" % s for s in synth.phoneme_segments(n, seed=seed)]


def test_oracle_restatement_matches_reference_rows(oracle):
    g = load_golden("labels.npz")
    want = golden_rows(g)
    got = []
    # GenerateLabelData walks sorted(glob): TEST before TRAIN, then by file name
    for (tt, dr, spk, sent) in sorted(UTTS, key=lambda u: os.path.join(u[0], ".".join(u[1:]))):
        key = "%s_%s_%s_%s" % (tt, dr, spk, sent)
        track = np.round(g["tracks_" + key][:, 1].astype(np.float64) * 1000, 2)
        rows = oracle.extract_label(track, segments_of(g, key), int(g["n_" + key]), [tt, dr, spk, sent])
        got.extend(rows or [])
    assert len(got) == len(want) == 489
    assert [[str(v) for v in r] for r in got] == want
    # single timepoints against the raw (slope, intercept, r, p) table of one file
    raw = g["raw"]
    track = g["fb_hz"][:, 1]
    for row in raw[::37]:
        t = int(row[0])
        a, b, r, p = oracle.label_fit(track[t // 160 - 5:t // 160 + 6], np.array([t + (j - 5) * 160 for j in range(11)]))
        assert (a, b, r, p) == tuple(row[1:])


def test_fb_and_phn_readers_and_wav_header(label_tree):
    from f2cnn_b200.scripts.processing import FBFileReader as FB, PHNFileReader as PHN
    from f2cnn_b200.scripts.processing.LabelDataGenerator import wav_shape
    g = load_golden("labels.npz")
    stem = os.path.join("resources", "f2cnn", "TEST", "DR2.FCCC0.SA1")
    hz, period = FB.ExtractFBFile(stem + ".FB")
    assert period == 10000 and hz.dtype == np.float64 and np.array_equal(hz, g["fb_hz"])   # reference's own output
    f2, _ = FB.GetFormantFrequencies(stem + ".FB", 2)
    assert np.array_equal(f2, g["fb_hz"][:, 1])
    assert np.array_equal(FB.GetFromantFrequenciesAround(f2, 8000, 5, 160.0), f2[45:56])
    assert FB.ExtractFBFile("missing.FB") == (None, 0) and FB.GetFormantFrequencies("missing.FB", 2) == (None, None)
    with pytest.raises(SystemExit):
        FB.GetFromantFrequenciesAround(f2, 100, 5, 160.0)       # start < 0
    with pytest.raises(SystemExit):
        FB.GetFromantFrequenciesAround(f2, 160 * (len(f2) - 5), 5, 160.0)   # end >= len
    segs = PHN.ExtractPhonemes(stem + ".PHN")
    assert segs == segments_of(g, "TEST_DR2_FCCC0_SA1") and PHN.ExtractPhonemes("missing.PHN") is None
    ts = np.arange(0, 52000, 160)
    assert PHN.phoneme_at(segs, ts) == [PHN.GetPhonemeFromArrayAt(segs, int(t)) for t in ts]
    assert PHN.GetPhonemeFromArrayAt(segs, segs[1][1]) == segs[0][0]    # shared boundary: the earlier segment
    assert PHN.GetPhonemeFromArrayAt(segs, 10 ** 9) == 'h#' and PHN.GetPhonemeAt(stem + ".PHN", 3000) == segs[1][0]
    assert wav_shape(stem + ".WAV") == (16000, 52000)
    # NIST SPHERE header and a float64 RIFF file (the noise-mixed WAV of cnn evalnoise)
    sph = label_tree / "x.WAV"
    head = ("NIST_1A\n   1024\nsample_count -i 777\nsample_rate -i 8000\nchannel_count -i 1\nsample_n_bytes -i 2\n"
            "sample_byte_format -s2 01\nsample_coding -s3 pcm\nend_head\n").encode()
    sph.write_bytes(head.ljust(1024, b" ") + b"\0" * 1554)
    assert wav_shape(str(sph)) == (8000, 777)
    from scipy.io import wavfile
    wavfile.write(str(label_tree / "f64.WAV"), 16000, np.zeros(1234, dtype=np.float64))
    assert wav_shape(str(label_tree / "f64.WAV")) == (16000, 1234)


@pytest.mark.gpu
def test_label_fit_kernel_matches_reference_library_calls():
    import torch
    assert torch.cuda.is_available()
    from f2cnn_b200 import api
    g = load_golden("labels.npz")
    raw = g["raw"]
    track = g["fb_hz"][:, 1]
    centers = raw[:, 0].astype(np.int64)
    fit = api.label_fit([track], [centers // 160 - 5], [centers], radius=5, step=160)
    assert fit.shape == (len(raw), 4) and fit.dtype == np.float64
    want = raw[:, 1:]
    assert np.max(np.abs(fit[:, 0] - want[:, 0]) / np.maximum(np.abs(want[:, 0]), 1e-12)) <= 1e-9   # slope
    assert np.max(np.abs(fit[:, 1] - want[:, 1]) / np.abs(want[:, 1])) <= 1e-9                        # intercept
    # r, p: the reference correlates the values with the ROUNDED fitted line a*x+b, which costs it ~1e-11
    # when the line is almost flat (observed 1.7e-11 at r = 0.007); the CSV keeps 5 decimals
    assert np.max(np.abs(fit[:, 2] - want[:, 2])) <= 1e-9                                             # r
    assert np.max(np.abs(fit[:, 3] - want[:, 3])) <= 1e-9                                             # p
    # degenerate windows: flat track -> NaN like scipy's constant-input path; exact line -> r = 1, p = 0
    flat = np.full(40, 1500.0)
    line = 1000.0 + 0.25 * 160 * np.arange(40)
    c = np.array([800, 1600, 3200], dtype=np.int64)
    out = api.label_fit([flat, line], [c // 160 - 5, c // 160 - 5], [c, c])
    assert np.all(np.isnan(out[:3, 2:])) and np.all(out[:3, 0] == 0)
    assert np.allclose(out[3:, 0], 0.25, rtol=1e-12) and np.all(out[3:, 2] > 1 - 1e-12) and np.all(out[3:, 3] < 1e-12)
    with pytest.raises(IndexError):
        api.label_fit([flat], [np.array([35])], [np.array([6400])])


@pytest.mark.gpu
def test_generate_label_data_and_extract_label_drop_ins(label_tree):
    import torch
    assert torch.cuda.is_available()
    from configparser import ConfigParser
    from f2cnn_b200 import dropin
    g = load_golden("labels.npz")
    dropin.install()
    try:
        from scripts.processing import LabelDataGenerator as LG
        assert LG.__name__.startswith("f2cnn_b200.")
        LG.GenerateLabelData()
        with open(os.path.join("trainingData", "label_data.csv")) as f:
            text = f.read()
        assert_same_rows([r for r in csv.reader(io.StringIO(text))], golden_rows(g))
        cfg = ConfigParser()
        cfg.read("configF2CNN.conf")
        rows = LG.ExtractLabel(os.path.join("resources", "f2cnn", "TEST", "DR2.FCCC0.SA1.WAV"), cfg)
        want = ast.literal_eval(re.sub(r"np\.float64\(([^)]*)\)", r"\1", str(g["one_file_rows"])))
        assert_same_rows([[str(v) for v in r] for r in rows], [[str(v) for v in r] for r in want])
        os.remove(os.path.join("resources", "f2cnn", "TEST", "DR2.FCCC0.SA1.FB"))
        assert LG.ExtractLabel(os.path.join("resources", "f2cnn", "TEST", "DR2.FCCC0.SA1.WAV"), cfg) is None
    finally:
        dropin.uninstall()
