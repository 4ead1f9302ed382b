"""Pins the CPU oracle (oracle/f2_oracle.c) to outputs of the UNMODIFIED reference
recorded in tests/golden/ by oracle/make_golden.py.  float64 against float64: the only
differences are operation order inside the FFT / gain product, so the bar is
1e-9 x per-channel RMS (observed ~1e-13)."""
import csv
import io

import numpy as np
import pytest

from conftest import load_golden

REL = 1e-9


def _close(got, want, scale, rel=REL):
    scale = np.maximum(np.asarray(scale, dtype=np.float64), 1e-300)
    err = np.max(np.abs(got - want), axis=-1) / scale
    assert np.all(err <= rel), "max rel err %.3e" % err.max()


def test_coefficients(oracle):
    g = load_golden("coefs.npz")
    for tag, (fs, C, low) in {"fs16000_c128_l100": (16000, 128, 100), "fs16000_c256_l100": (16000, 256, 100),
                              "fs16000_c8_l100": (16000, 8, 100), "fs8000_c32_l50": (8000, 32, 50),
                              "fs44100_c64_l20": (44100, 64, 20)}.items():
        cf = oracle.centre_freqs(fs, C, low)
        np.testing.assert_allclose(cf, g["cf_" + tag], rtol=1e-14, atol=0)
        co = oracle.make_erb_filters(fs, g["cf_" + tag])
        np.testing.assert_allclose(co, g["coefs_" + tag], rtol=2e-12, atol=0)
        assert cf[0] > cf[-1]  # descending: high -> low channel order
    np.testing.assert_allclose(oracle.erb_space(100, 44100 / 4, 100), g["erb_space_default"], rtol=1e-14)
    np.testing.assert_allclose(oracle.make_erb_filters(16000, oracle.centre_freqs(16000, 16, 100), 2.0),
                               g["coefs_width2"], rtol=2e-12)
    for f in (20, 50, 100, 400):
        b, a = oracle.butter1_lowpass(f / 8000.0)
        np.testing.assert_allclose(np.concatenate([b, a]), g["butter_%d" % f], rtol=1e-13)


@pytest.mark.parametrize("name", ["white", "delta", "tone1k", "chirp", "speech"])
def test_utt3s(oracle, name):
    g = load_golden("utt3s_%s.npz" % name)
    co = load_golden("coefs.npz")["coefs_fs16000_c128_l100"]
    idx = g["idx"]
    gfb = oracle.erb_filterbank(g["wave"], co)
    _close(gfb[:, idx], g["gfb"], g["gfb_rms"])
    np.testing.assert_allclose(np.sqrt(np.mean(gfb ** 2, axis=1)), g["gfb_rms"], rtol=1e-10)
    e50 = oracle.extract_envelope(gfb, True, 50)
    _close(e50[:, idx], g["env_lpf50"], g["env_lpf50_rms"])
    eno = oracle.extract_envelope(gfb, False)
    _close(eno[:, idx], g["env_nolpf"], g["env_nolpf_rms"])
    np.testing.assert_allclose(e50.sum(axis=1), g["env_lpf50_sum"], rtol=1e-9)
    if name == "white":
        _close(oracle.extract_envelope(gfb, True, 20)[:, idx], g["env_lpf20"], g["env_lpf20_rms"])
        _close(oracle.extract_envelope(gfb, True, 100)[:, idx], g["env_lpf100"], g["env_lpf100_rms"])
        _close(e50[:, g["dec_idx"]], g["env_lpf50_dec"], g["env_lpf50_rms"])


def test_small_lengths(oracle):
    g = load_golden("small.npz")
    co = g["coefs"]
    for nn in (1, 2, 3, 4, 5, 16, 17, 255, 256, 257, 1000, 4096, 4097):
        w = g["wave_%d" % nn]
        gfb = oracle.erb_filterbank(w, co)
        assert gfb.shape == (8, nn) and gfb.dtype == np.float64
        scale = np.sqrt(np.mean(g["gfb_%d" % nn] ** 2, axis=1))
        _close(gfb, g["gfb_%d" % nn], scale)
        for key, (lpf, cut) in {"env_lpf50": (True, 50), "env_nolpf": (False, 100)}.items():
            want = g["%s_%d" % (key, nn)]
            _close(oracle.extract_envelope(gfb, lpf, cut), want, np.sqrt(np.mean(want ** 2, axis=1)))


def test_pow2_lengths(oracle):
    g = load_golden("pow2.npz")
    co = g["coefs"]
    for nn in (65530, 65535, 65536, 65537):
        idx = g["idx_%d" % nn]
        gfb = oracle.erb_filterbank(g["wave_%d" % nn], co)
        _close(gfb[:, idx], g["gfb_%d" % nn], g["gfb_rms_%d" % nn])
        _close(oracle.extract_envelope(gfb, True, 50)[:, idx], g["env_lpf50_%d" % nn], g["env_lpf50_rms_%d" % nn])
        _close(oracle.extract_envelope(gfb, False)[:, idx], g["env_nolpf_%d" % nn], g["env_nolpf_rms_%d" % nn])


def test_c256_float64_input_and_dense_frames(oracle):
    g = load_golden("c256_f64.npz")
    co = load_golden("coefs.npz")["coefs_fs16000_c256_l100"]
    idx = g["idx"]
    gfb = oracle.erb_filterbank(g["wave"], co)
    _close(gfb[:, idx], g["gfb"], g["gfb_rms"])
    env = oracle.extract_envelope(gfb, True, 50)
    _close(env[:, idx], g["env_lpf50"], g["env_lpf50_rms"])
    for j, i in enumerate(g["frames_i"]):
        fr = oracle.dense_frames(env, 5, 160, int(i), int(i) + 1)[0]
        np.testing.assert_allclose(fr, g["frames"][j], rtol=1e-9, atol=1e-9 * g["env_lpf50_rms"].max())
        np.testing.assert_allclose(oracle.normalize_input(g["frames"][j]), g["frames_norm"][j], rtol=1e-12,
                                   atol=1e-13)


def test_normalize_input_errors(oracle):
    with pytest.raises(ValueError):
        oracle.normalize_input(np.array([[1.0, 0.0], [2.0, 3.0]]))
    assert np.all(oracle.normalize_input(np.full((11, 4), 2.5)) == 0.0)


def test_input_generator_row_order(oracle):
    """InputGenerator.py:50,67-82: rows ordered by sorted file key, CSV order within a file;
    float32 cast only at the end (:83)."""
    g = load_golden("inputgen.npz")
    co = g["coefs"]
    files = {}
    for row in csv.reader(io.StringIO(str(g["csv"]))):
        key = (row[0], row[1], row[2], row[3])
        files.setdefault("%s/%s.%s.%s" % key, (key, []))[1].append(int(row[5]))
    out = []
    for fkey in sorted(files):
        key, tps = files[fkey]
        w = g["wave_%s_%s_%s_%s" % key]
        env = oracle.extract_envelope(oracle.erb_filterbank(w, co), True, 50)
        out.append(oracle.gather_windows(env, tps, 5, 160))
    got = np.concatenate(out)
    want = g["input_data"]
    assert got.shape == want.shape and got.dtype == want.dtype == np.float32
    # float32 rounding of float64 values that differ in the last bits can flip one ulp
    np.testing.assert_allclose(got, want, rtol=2e-7, atol=0)


def test_gather_windows_python_indexing(oracle):
    env = np.arange(2 * 2000, dtype=np.float64).reshape(2, 2000)
    w = oracle.gather_windows(env, [800, 100], 5, 160)  # 100-800 < 0 wraps like a Python list index
    assert w[0, 0, 0] == env[0, 0] and w[0, 10, 1] == env[1, 1600]
    assert w[1, 0, 0] == env[0, 2000 - 700]
    with pytest.raises(IndexError):
        oracle.gather_windows(env, [1300], 5, 160)
