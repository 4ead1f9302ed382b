"""The reference-named modules end to end on the GPU: file drivers (`prepare filter`,
`prepare envelope`, `prepare input`) in a temporary tree, and the small array functions."""
import csv
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-4
CONF = ("[FILTERBANK]\nFRAMERATE=16000\nNCHANNELS=32\nLOW_FREQ=100\n"
        "[CNN]\nFORMANT=2\nCENTERED=True\nRADIUS=5\nBATCH_SIZE=32\nEPOCHS=20\nRISK=0.05\nSAMPLING_PERIOD=10000\n")


def rel(got, want):
    r = np.sqrt(np.mean(want ** 2, axis=-1))
    return (np.max(np.abs(got - want), axis=-1) / np.maximum(r, 1e-300)).max()


@pytest.fixture()
def tree(tmp_path, monkeypatch):
    from scipy.io import wavfile
    from f2cnn_b200 import synth
    monkeypatch.chdir(tmp_path)
    (tmp_path / "configF2CNN.conf").write_text(CONF)
    waves = {}
    for tt, name, n in (("TRAIN", "DR1.SPK0.SA1", 9000), ("TRAIN", "DR2.SPK1.SX7", 12000), ("TEST", "DR1.SPK2.SI3", 7000)):
        d = tmp_path / "resources" / "f2cnn" / tt
        d.mkdir(parents=True, exist_ok=True)
        w = synth.speech_like_i16(n, seed=n)
        wavfile.write(str(d / (name + ".WAV")), 16000, w)
        synth.write_fb(str(d / (name + ".FB")), synth.formant_tracks_khz(n // 160 + 3, seed=n))
        synth.write_phn(str(d / (name + ".PHN")), synth.phoneme_segments(n, seed=n))
        waves[(tt, name)] = w
    return tmp_path, waves


def test_prepare_filter_envelope_input_drivers(tree, oracle):
    tmp_path, waves = tree
    import torch
    assert torch.cuda.is_available()
    from f2cnn_b200 import dropin, synth
    dropin.install()
    try:
        from gammatone import filters
        from scripts.processing import EnvelopeExtraction, GammatoneFiltering, InputGenerator
        co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 32, 100))
        GammatoneFiltering.FilterAllOrganisedFiles()
        EnvelopeExtraction.ExtractAllEnvelopes(True, 50)
        rows, want_rows = [], {}
        for (tt, name), w in sorted(waves.items()):
            base = tmp_path / "resources" / "f2cnn" / tt / name
            gfb = np.load(str(base) + ".GFB.npy")
            env = np.load(str(base) + ".ENV1.npy")
            assert gfb.dtype == env.dtype == np.float64 and gfb.shape == env.shape == (32, len(w))
            go, eo, _ = oracle.utterance(w, co, True, 50)
            assert rel(gfb, go) <= TOL and rel(env, eo) <= TOL
            assert np.array_equal(GammatoneFiltering.loadGFBMatrix(str(base) + ".GFB"), gfb)
            dr, spk, sent = name.split(".")
            tps = synth.label_grid(len(w))[::2]
            for tp in tps:
                rows.append([tt, dr, spk, sent, "aa", int(tp), 0.1, 0.01, 1])
            want_rows[os.path.join(tt, name + ".ENV1.npy")] = oracle.gather_windows(env, tps)
        os.makedirs("trainingData")
        with open(os.path.join("trainingData", "label_data.csv"), "w") as f:
            wr = csv.writer(f, lineterminator="\n")
            for r in rows[::-1]:  # reversed CSV: rows must come out by sorted file key, CSV order within a file
                wr.writerow(r)
        InputGenerator.GenerateInputData(LPF=True, CUTOFF=50)
        out = np.load(os.path.join("trainingData", "input_data_LPF50.npy"))
        assert out.dtype == np.float32 and np.array_equal(out, np.load(os.path.join("trainingData", "last_input_data.npy")))
        want = np.concatenate([want_rows[k][::-1] for k in sorted(want_rows)])
        assert out.shape == want.shape
        assert np.array_equal(out, want)  # pure gather + float64->float32 cast of the saved envelopes: exact
        # ENV1 file missing but WAV present: rows come from the fused kernel instead
        os.remove(str(tmp_path / "resources" / "f2cnn" / "TEST" / "DR1.SPK2.SI3.ENV1.npy"))
        InputGenerator.GenerateInputData(inputFile=os.path.join("trainingData", "fused.npy"), LPF=True, CUTOFF=50)
        out2 = np.load(os.path.join("trainingData", "fused.npy"))
        scale = np.sqrt(np.mean(want.astype(np.float64) ** 2, axis=(0, 1)))
        assert out2.shape == want.shape and np.max(np.abs(out2 - want) / scale[None, None, :]) <= TOL
    finally:
        dropin.uninstall()


def test_array_functions(oracle):
    import torch
    assert torch.cuda.is_available()
    from f2cnn_b200 import synth
    from f2cnn_b200.scripts.processing import EnvelopeExtraction as EE, GammatoneFiltering as GF
    from f2cnn_b200.gammatone import filters
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 16, 100))
    w = synth.white_noise_i16(5000, seed=8)
    gfb = GF.GetFilteredOutputFromArray(w, co)
    assert rel(gfb, oracle.erb_filterbank(w, co)) <= TOL
    row = gfb[7]
    an = EE.paddedHilbert(row)
    want = oracle.padded_hilbert(row)
    assert an.dtype == np.complex128 and np.array_equal(an.real, row)
    assert np.max(np.abs(an.imag - want.imag)) <= TOL * np.sqrt(np.mean(row ** 2))
    lp = EE.lowPassFilter(np.abs(want), 50)
    assert np.max(np.abs(lp - oracle.low_pass_filter(np.abs(want), 50))) <= TOL * np.sqrt(np.mean(np.abs(want) ** 2))
    env = EE.ExtractEnvelopeFromMatrix(gfb)  # defaults: LPF=False
    assert rel(env, oracle.extract_envelope(gfb, False)) <= TOL
    # lfilter(axis=0) semantics on a 2-D array: every column is one signal
    cols = np.abs(gfb[:3]).T.copy()
    lp2 = EE.lowPassFilter(cols, 120)
    assert lp2.shape == cols.shape
    for j in range(3):
        wantj = oracle.low_pass_filter(cols[:, j].copy(), 120)
        assert np.max(np.abs(lp2[:, j] - wantj)) <= TOL * np.sqrt(np.mean(wantj ** 2))


def test_evaluating_front_end_and_driver(tree, oracle, monkeypatch):
    """scripts.CNN.Evaluating: frames handed to model.predict match the reference's framing +
    normalizeInput of the float64 envelopes; the driver calls predict/plot with them (Keras and
    the reference-tree readers are stand-ins: they are outside the hot path)."""
    import sys
    import types
    from configparser import ConfigParser
    tmp_path, waves = tree
    from f2cnn_b200 import dropin
    from f2cnn_b200.gammatone import filters
    dropin.install()
    try:
        from scripts.CNN import Evaluating
        cfg = ConfigParser()
        cfg.read("configF2CNN.conf")
        w = waves[("TEST", "DR1.SPK2.SI3")]
        frames, centre, co, step = Evaluating.PrepareInputFromArray(w, 16000, cfg, True, 50)
        assert step == 160 and frames.dtype == np.float64 and frames.shape == (len(w) - 11 * 160, 11, 32)
        assert np.array_equal(centre, filters.centre_freqs(16000, 32, 100))
        _, eo, _ = oracle.utterance(w, co, True, 50)
        pick = [0, 1, 500, 777, 2500, frames.shape[0] - 1]
        raw = [oracle.dense_frames(eo, 5, 160, i, i + 1)[0] for i in pick]
        want = np.stack([oracle.normalize_input(r.copy()) for r in raw])
        # log-min-max domain: an absolute envelope error e (bar: 1e-4 x channel RMS) moves log(x) by e/x, and
        # the frame's minimum sets the scale, so the bound is per frame: 2 * (1e-4 * RMS / min) / (log max - log min).
        # Frames 0 and 1 contain the filter start-up (envelope ~1e-3 of RMS) and get the larger bound.
        rms = np.sqrt(np.mean(eo ** 2, axis=1))
        for k, r in enumerate(raw):
            bound = 2 * np.max(TOL * rms[None, :] / r) / (np.log(r.max()) - np.log(r.min()))
            assert np.max(np.abs(frames[pick[k]] - want[k])) <= max(bound, 1e-6), (pick[k], bound)
        assert np.max(np.abs(frames[pick[2:]] - want[2:])) <= 2e-3   # away from the start-up: as the evalnoise case

        seen = {}

        class _Net:
            def predict(self, x, verbose=0):
                seen["x"] = x
                return np.stack([np.linspace(0, 1, x.shape[0]), np.linspace(1, 0, x.shape[0])], axis=1)

        keras = types.ModuleType("keras")
        keras.models = types.SimpleNamespace(load_model=lambda name: _Net())
        keras.backend = types.SimpleNamespace(clear_session=lambda: None)
        monkeypatch.setitem(sys.modules, "keras", keras)
        for name, attrs in {"scripts.plotting": {}, "scripts.plotting.PlottingCNN": dict(
                PlotEnvelopesAndCNNResultsWithPhonemes=lambda *a: seen.update(plot=a))}.items():
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            if not attrs:
                m.__path__ = []
            monkeypatch.setitem(sys.modules, name, m)
        wav = os.path.join("resources", "f2cnn", "TEST", "DR1.SPK2.SI3.WAV")
        Evaluating.EvaluateOneWavArray(w, 16000, wav, LPF=True, CUTOFF=50)
        assert seen["x"].shape == (frames.shape[0], 11, 32, 1) and np.array_equal(seen["x"][..., 0], frames)
        envs, scores, acc, cf, phonemes, formants = seen["plot"][:6]
        assert envs.shape == (32, len(w)) and rel(envs, eo) <= TOL and scores.shape == (frames.shape[0], 2)
        # labels, phonemes and formants come from the .FB / .PHN files next to the WAV
        from f2cnn_b200 import synth
        segs = [(name, a, b) for a, b, name in synth.phoneme_segments(len(w), seed=len(w))]
        track = np.round(synth.formant_tracks_khz(len(w) // 160 + 3, seed=len(w)).astype(np.float64) * 1000, 2)
        assert phonemes == segs and np.array_equal(formants, track)
        rows = oracle.extract_label(track[:, 1], segs, len(w), ["TEST", "DR1", "SPK2", "SI3"])
        assert rows, "fixture should give at least one label"
        labels = [(r[-4], r[-1]) for r in rows]
        dec = [int(s[1] > s[0]) for s in scores]
        good = n = 0
        for t, d in enumerate(dec):
            for (b, cb), (a, ca) in zip(labels[:-1], labels[1:]):
                if b < t < a and min(t - b, a - t) < 160:
                    n += 1
                    good += int(d == (cb if t - b <= a - t else ca))
        assert n > 0 and acc == good / n
    finally:
        dropin.uninstall()


def test_features_to_windows_from_one_ingested_buffer(tree, oracle):
    """ingest.read_corpus -> api.features_to_windows((flat, lengths), ...) equals the list-of-arrays
    call bit for bit and the oracle within the bar."""
    import torch
    assert torch.cuda.is_available()
    tmp_path, waves = tree
    from f2cnn_b200 import api, ingest, synth
    from f2cnn_b200.gammatone import filters
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 32, 100))
    keys = sorted(waves)
    paths = [str(tmp_path / "resources" / "f2cnn" / tt / (name + ".WAV")) for tt, name in keys]
    flat, lengths, rates = ingest.read_corpus(paths)
    assert rates == [16000] * 3 and lengths.tolist() == [len(waves[k]) for k in keys]
    assert flat.is_pinned() and np.array_equal(flat.numpy(), np.concatenate([waves[k] for k in keys]))
    tps = [synth.label_grid(int(n))[::3] for n in lengths]
    a = api.features_to_windows((flat, lengths), co, tps, True, 50)
    b = api.features_to_windows([waves[k] for k in keys], co, tps, True, 50)
    assert a.dtype == np.float32 and np.array_equal(a, b)
    row = 0
    for k, tp in zip(keys, tps):
        _, eo, wo = oracle.utterance(waves[k], co, True, 50, tp)
        scale = np.sqrt(np.mean(eo ** 2, axis=1))
        assert np.max(np.abs(a[row:row + len(tp)] - wo) / scale[None, None, :]) <= TOL
        row += len(tp)
    with pytest.raises(ValueError):
        api.features_to_windows((flat[:-1], lengths), co, tps, True, 50)


def test_float32_npy_option_of_the_file_drivers(tree, oracle, monkeypatch):
    """F2CNN_B200_NPY_FLOAT32=1: `prepare filter` / `prepare envelope` store float32 matrices (SURVEY.md
    8f rank 4); `prepare input` reads them and produces the same rows within float32 rounding."""
    import torch
    assert torch.cuda.is_available()
    tmp_path, waves = tree
    from f2cnn_b200 import dropin, synth
    monkeypatch.setenv("F2CNN_B200_NPY_FLOAT32", "1")
    dropin.install()
    try:
        from gammatone import filters
        from scripts.processing import EnvelopeExtraction, GammatoneFiltering, InputGenerator
        co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 32, 100))
        GammatoneFiltering.FilterAllOrganisedFiles()
        EnvelopeExtraction.ExtractAllEnvelopes(True, 50)
        rows, want = [], []
        for (tt, name), w in sorted(waves.items()):
            base = tmp_path / "resources" / "f2cnn" / tt / name
            gfb, env = np.load(str(base) + ".GFB.npy"), np.load(str(base) + ".ENV1.npy")
            assert gfb.dtype == env.dtype == np.float32 and gfb.shape == env.shape == (32, len(w))
            go, eo, _ = oracle.utterance(w, co, True, 50)
            assert rel(gfb, go) <= TOL and rel(env, eo) <= 2 * TOL   # envelope of the float32-rounded GFB
            dr, spk, sent = name.split(".")
            tps = synth.label_grid(len(w))[::2]
            rows += [[tt, dr, spk, sent, "aa", int(tp), 0.1, 0.01, 1] for tp in tps]
            want.append(oracle.gather_windows(env.astype(np.float64), tps))
        os.makedirs("trainingData")
        with open(os.path.join("trainingData", "label_data.csv"), "w") as f:
            csv.writer(f, lineterminator="\n").writerows(rows)
        InputGenerator.GenerateInputData(LPF=True, CUTOFF=50)
        out = np.load(os.path.join("trainingData", "input_data_LPF50.npy"))
        assert out.dtype == np.float32 and np.array_equal(out, np.concatenate(want))
    finally:
        dropin.uninstall()


def test_plot_gtg_runs_on_the_decimated_envelope(tree, oracle, monkeypatch):
    """`plot gtg` (scripts/plotting/PlottingProcessing.py:81-133): the image handed to imshow is the reference's
    (every channel repeated by its ERB ratio) at every hop-th column, and the columns are the full-rate
    envelope's own samples.  matplotlib is absent here: a recording stand-in takes the calls."""
    import sys
    import types
    tmp_path, waves = tree
    from f2cnn_b200 import api
    from f2cnn_b200.gammatone import filters
    from f2cnn_b200.scripts.plotting import PlottingProcessing as pp
    w = waves[("TRAIN", "DR2.SPK1.SX7")]
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
    full = api.filterbank_envelope(w, co, False, 100)
    for hop in (1, 7, 160):
        g = api.gammatonegram(w, co, hop)
        assert g.dtype == np.float64 and g.shape == (128, -(-len(w) // hop))
        assert np.array_equal(g, full[:, ::hop])
    assert rel(api.gammatonegram(w, co, 3), oracle.utterance(w, co, False, 100)[1][:, ::3]) <= TOL
    calls = {}
    plt = types.ModuleType("matplotlib.pyplot")
    plt.imshow = lambda image, **kw: calls.update(image=image, kw=kw)
    for name in ("plot", "legend", "text", "title", "show"):
        setattr(plt, name, lambda *a, _n=name, **k: calls.setdefault(_n, []).append(a))
    mpl = types.ModuleType("matplotlib")
    colors = types.ModuleType("matplotlib.colors")
    colors.LogNorm = lambda: "lognorm"
    mpl.pyplot, mpl.colors = plt, colors
    for name, mod in (("matplotlib", mpl), ("matplotlib.pyplot", plt), ("matplotlib.colors", colors)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.setattr(pp, "MAX_COLUMNS", 1000)
    path = str(tmp_path / "resources" / "f2cnn" / "TRAIN" / "DR2.SPK1.SX7.WAV")
    pp.PlotEnvelopesAndFormantsFromFile(path, formantToPlot=2)
    hop = -(-len(w) // 1000)
    cfs = filters.centre_freqs(16000, 128, 100)
    assert np.array_equal(calls["image"], pp.ReshapeEnvelopesForSpectrogram(full[:, ::hop], cfs))
    assert calls["kw"]["extent"] == [0, calls["image"].shape[1] / (16000 / hop), 100, 8000]
    assert len(calls["plot"]) == 1 and len(calls["show"]) == 1
