"""The CNN forward of `cnn eval*` on the 5th-generation tensor cores (csrc/f2_cnn.cu) against a plain
float32 PyTorch forward of the same network (scripts/CNN/Training.py:93-114; no Keras here, no trained
model ships with the reference: seeded weights).  Tolerances, stated once: the kernels keep activations
and weights in bf16 (8 significant bits) with float32 accumulation, so intermediate tensors are held to
2e-2 of their largest value and the softmax scores to 5e-3 absolute; the decision (argmax) must agree
wherever the oracle's margin exceeds 2e-2."""
import os
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CONF128 = ("[FILTERBANK]\nFRAMERATE=16000\nNCHANNELS=128\nLOW_FREQ=100\n"
           "[CNN]\nFORMANT=2\nCENTERED=True\nRADIUS=5\nBATCH_SIZE=32\nEPOCHS=20\nRISK=0.05\nSAMPLING_PERIOD=10000\n")


def _coefs():
    from f2cnn_b200.gammatone import filters
    return filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))


def _oracle_forward(model, env_t, i0, i1, torch):
    """normalizeInput (Training.py:13-28, float64) + the float32 network, layer by layer."""
    import torch.nn.functional as F
    idx = torch.arange(i0, i1, device=env_t.device)[:, None] + 160 * torch.arange(11, device=env_t.device)[None, :]
    fr = env_t[idx].double()
    lo = fr.amin(dim=(1, 2), keepdim=True).log()
    hi = fr.amax(dim=(1, 2), keepdim=True).log()
    x = ((fr.log() - lo) / (hi - lo)).float().unsqueeze(1)
    with torch.no_grad():
        p2 = F.max_pool2d(F.relu(model.c2(F.relu(model.c1(x)))), 2)
        p4 = F.max_pool2d(F.relu(model.c4(F.relu(model.c3(p2)))), 2)
        feat = p4.permute(0, 2, 3, 1).reshape(p4.shape[0], -1)
        scores = F.softmax(model.d2(F.relu(model.d1(feat))), dim=1)
    return p2, feat, scores


@pytest.fixture(scope="module")
def setup():
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from f2cnn_b200 import cnn, engine, synth
    n = 20000
    w = synth.speech_like_i16(n, seed=11).astype(np.float64) + np.random.default_rng(2).normal(0, 30, n)
    plan = engine.plan_for(_coefs())
    env_t = plan.batch([n]).run(torch.from_numpy(w).cuda(), lpf=True, cutoff=50, env_t=True)["env_t"]
    model = cnn.seeded_model(seed=3)
    return torch, cnn, w, env_t, model


def test_tcgen05_operand_conventions_against_torch():
    """128 x N x K products through f2_umma_selftest: the K-major plane layout and the shifted
    descriptor the convolutions rest on (tools/umma_selftest.py runs the wider sweep)."""
    import torch
    from f2cnn_b200 import _native
    L = _native.lib()
    torch.manual_seed(1)
    for N, K, shift, rows in ((32, 288, 0, 128), (64, 64, 131, 264), (176, 64, 0, 128), (32, 32, 1, 136)):
        A = torch.randn(rows, K, device="cuda").to(torch.bfloat16).contiguous()
        B = torch.randn(N, K, device="cuda").to(torch.bfloat16).contiguous()
        D = torch.zeros((128, N), device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        _native.check(L.f2_umma_selftest(A.data_ptr(), rows, B.data_ptr(), N, K, shift, 0, D.data_ptr(), status.data_ptr(), None))
        want = A[shift:shift + 128].float() @ B.float().t()
        assert int(status.item()) == 0
        assert float((D - want).abs().max()) <= 1e-4 * float(want.abs().max())


def test_tensor_core_network_matches_the_float32_oracle_layer_by_layer(setup):
    torch, cnn, w, env_t, model = setup
    tc = cnn.TensorCoreCNN(model)
    nb = env_t.shape[0] - 1760
    for i0, i1 in ((0, 300), (5000, 5700), (nb - 257, nb)):
        got = tc.predict_envelope(env_t, 160, frames=(i0, i1))
        torch.cuda.synchronize()
        gp2, gfeat = tc.intermediates(i1 - i0)
        p2, feat, scores = _oracle_forward(model, env_t, i0, i1, torch)
        assert float((gp2 - p2).abs().max()) <= 2e-2 * float(p2.abs().max())
        assert float((gfeat - feat).abs().max()) <= 2e-2 * float(feat.abs().max())
        assert float((got - scores).abs().max()) <= 5e-3
        assert torch.allclose(got.sum(dim=1), torch.ones_like(got[:, 0]), atol=1e-5)
    # all frames of the utterance, several chunks of the persistent grid
    got = tc.predict_envelope(env_t, 160)
    ref = torch.cat([_oracle_forward(model, env_t, i, min(i + 4096, nb), torch)[2] for i in range(0, nb, 4096)])
    assert got.shape == (nb, 2) and float((got - ref).abs().max()) <= 5e-3
    clear = (ref[:, 0] - ref[:, 1]).abs() > 2e-2
    assert bool((got.argmax(1) == ref.argmax(1))[clear].all())
    # deterministic
    assert torch.equal(got, tc.predict_envelope(env_t, 160))


def test_tensor_core_network_from_keras_layout_arrays_and_edge_cases(setup, tmp_path):
    torch, cnn, w, env_t, model = setup
    arrays = cnn.keras_arrays(model)
    assert [a.shape for a in arrays][::2] == [(3, 3, 1, 32), (3, 3, 32, 32), (3, 3, 32, 64), (3, 3, 64, 64), (1920, 516), (516, 2)]
    path = str(tmp_path / "weights.npz")
    cnn.save_weights(path, model)
    loaded = cnn.load_weights(path)
    assert all(np.array_equal(a, b) for a, b in zip(arrays, loaded))
    a = cnn.TensorCoreCNN(model).predict_envelope(env_t, 160, frames=(100, 400))
    b = cnn.TensorCoreCNN(loaded).predict_envelope(env_t, 160, frames=(100, 400))
    assert torch.equal(a, b)
    # normalizeInput raises on values <= 0 (Training.py:18-20)
    bad = env_t.clone()
    bad[2000, 5] = 0.0
    with pytest.raises(ValueError, match="positive"):
        cnn.TensorCoreCNN(model).predict_envelope(bad, 160, frames=(1900, 2100))
    assert cnn.TensorCoreCNN(model).predict_envelope(bad, 160, frames=(2001, 2100)).shape == (99, 2)
    # a flat frame normalises to zeros (:21-23): scores are those of the all-zero input
    flat = torch.full_like(env_t[:4000], 3.0)
    z = cnn.TensorCoreCNN(model).predict_envelope(flat, 160, frames=(0, 10))
    with torch.no_grad():
        zero_scores = model(torch.zeros((1, 11, 128), device="cuda"))
    assert float((z - zero_scores).abs().max()) <= 5e-3
    # empty range, out-of-range frames, wrong geometry
    assert cnn.TensorCoreCNN(model).predict_envelope(env_t, 160, frames=(7, 7)).shape == (0, 2)
    with pytest.raises(IndexError):
        cnn.TensorCoreCNN(model).predict_envelope(env_t, 160, frames=(0, env_t.shape[0]))
    from f2cnn_b200 import _native
    with pytest.raises(_native.F2Error) as exc:
        cnn.TensorCoreCNN(model, dots=11, channels=64)
    assert exc.value.code == _native.F2_ERR_UNSUPPORTED
    with pytest.raises(ValueError):
        cnn.TensorCoreCNN(arrays[:-1])


def test_cnn_eval_driver_runs_on_the_tensor_core_network(setup, tmp_path, monkeypatch, oracle):
    """scripts.CNN.Evaluating.EvaluateOneWavFile with an .npz model (no Keras anywhere): one filterbank pass,
    scores from the tensor-core kernels, envelopes for the figure at float64 parity with the oracle."""
    torch, cnn, _, _, model = setup
    from scipy.io import wavfile
    from f2cnn_b200 import dropin, synth
    monkeypatch.chdir(tmp_path)
    (tmp_path / "configF2CNN.conf").write_text(CONF128)
    d = tmp_path / "resources" / "f2cnn" / "TEST"
    d.mkdir(parents=True)
    n = 9000
    w = synth.speech_like_i16(n, seed=n)
    w = np.where(w == 0, 1, w).astype(np.int16)      # silence would make normalizeInput raise, as in the reference
    w = (w.astype(np.int32) + np.random.default_rng(5).integers(-40, 41, n)).clip(-32768, 32767).astype(np.int16)
    wav = str(d / "DR1.SPK2.SI3.WAV")
    wavfile.write(wav, 16000, w)
    synth.write_fb(str(d / "DR1.SPK2.SI3.FB"), synth.formant_tracks_khz(n // 160 + 3, seed=n))
    synth.write_phn(str(d / "DR1.SPK2.SI3.PHN"), synth.phoneme_segments(n, seed=n))
    cnn.save_weights(str(tmp_path / "last_trained_model.npz"), model)
    seen = {}
    for name, attrs in {"scripts.plotting": {}, "scripts.plotting.PlottingCNN": dict(
            PlotEnvelopesAndCNNResultsWithPhonemes=lambda *a: seen.update(plot=a))}.items():
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        if not attrs:
            m.__path__ = []
        monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.setitem(sys.modules, "keras", None)   # importing keras must not be attempted
    dropin.install()
    try:
        from scripts.CNN import Evaluating
        Evaluating.EvaluateOneWavFile(os.path.join("resources", "f2cnn", "TEST", "DR1.SPK2.SI3.WAV"), LPF=True, CUTOFF=50)
    finally:
        dropin.uninstall()
    envs, scores, acc, cf = seen["plot"][:4]
    co = _coefs()
    _, eo, _ = oracle.utterance(w, co, True, 50)
    assert envs.shape == (128, n) and envs.dtype == np.float64
    assert (np.max(np.abs(envs - eo), axis=1) / np.sqrt(np.mean(eo ** 2, axis=1))).max() <= 1e-4
    assert scores.shape == (n - 1760, 2) and scores.dtype == np.float32
    # the oracle network on the oracle's float64 envelopes
    env_t = torch.from_numpy(eo.T.copy()).cuda().float()
    ref = torch.cat([_oracle_forward(model, env_t, i, min(i + 2048, n - 1760), torch)[2] for i in range(0, n - 1760, 2048)])
    assert float((torch.from_numpy(scores).cuda() - ref).abs().max()) <= 5e-3
    assert acc is None or 0.0 <= acc <= 1.0
