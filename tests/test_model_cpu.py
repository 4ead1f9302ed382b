"""CPU-only evidence for the algorithm the kernels implement (no GPU needed): the float32
model of the streaming ring formulation (oracle/f2_model.c -- delta-form biquads without the
common numerator gain, per-stage coefficient dithering, edge residuals from the last w_edge
samples, periodic imaginary path from a w_imag warm-up, injection kernel G) agrees with the
float64 oracle of the reference algorithm to the tolerance the GPU path is held to."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

_fp = ctypes.POINTER(ctypes.c_float)
_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def model():
    d = os.path.join(ROOT, "oracle")
    subprocess.check_call(["make", "-s", "-C", d, "libf2model.so"])
    M = ctypes.CDLL(os.path.join(d, "libf2model.so"))
    M.f2m_run.argtypes = [_fp, _fp, _fp, ctypes.c_int64, ctypes.c_int64, _dp, ctypes.c_int, ctypes.c_int,
                          ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _fp, _fp]
    M.f2m_set_direct_min_cy.argtypes = [ctypes.c_double]
    return M


def ring_inputs(wave):
    """x zero-padded to N2, Im hilbert(x_padded) and the injection kernel G (float32)."""
    n = len(wave)
    N2 = 1
    while N2 < n:
        N2 *= 2
    X = np.zeros(N2)
    X[:n] = wave
    F = np.fft.fft(X)
    h = np.zeros(N2)
    h[0] = h[N2 // 2] = 1
    h[1:N2 // 2] = 2
    xi = np.imag(np.fft.ifft(F * h))
    tau = np.arange(N2)
    l0 = (tau - n) % N2
    l = np.where(l0 % 2 == 1, l0, (tau - n - 1) % N2)
    G = (2.0 / N2) / np.tan(np.pi * l / N2)
    return n, N2, X.astype(np.float32), xi.astype(np.float32), G.astype(np.float32)


@pytest.mark.parametrize("kind,n", [("white", 16000), ("speech", 24000), ("white", 16384)])
def test_float32_streaming_model_matches_float64_oracle(model, oracle, kind, n):
    from f2cnn_b200 import synth
    co = oracle.make_erb_filters(16000, oracle.centre_freqs(16000, 128, 100))
    w = synth.white_noise_i16(n, seed=3) if kind == "white" else synth.speech_like_i16(n, seed=3)
    n, N2, xf, xi, G = ring_inputs(w)
    gfb = np.empty((128, n), np.float32)
    env = np.empty((128, n), np.float32)
    b, a = oracle.butter1_lowpass(50 / 8000.0)
    go = oracle.erb_filterbank(w, co)
    eo = oracle.extract_envelope(go, True, 50)

    def rel(got, want):
        return (np.max(np.abs(got - want), axis=1) / np.sqrt(np.mean(want ** 2, axis=1))).max()

    # the kernel's two settings of the section form: envelope-only runs (direct form down to
    # 1+B1+B2 = 0.035) and runs that store the filterbank output (direct form down to 0.25)
    for min_cy, check_gfb in ((0.035, False), (0.25, True)):
        model.f2m_set_direct_min_cy(min_cy)
        model.f2m_run(xf.ctypes.data_as(_fp), xi.ctypes.data_as(_fp), G.ctypes.data_as(_fp), n, N2,
                      np.ascontiguousarray(co).ctypes.data_as(_dp), 128, 1, b[0], a[1], 1536, 2048, 2,
                      gfb.ctypes.data_as(_fp), env.ctypes.data_as(_fp))
        assert rel(env, eo) <= 3e-5  # observed ~1e-5: an order of magnitude inside the 1e-4 bar
        if check_gfb:
            assert rel(gfb, go) <= 3e-5


def test_direct_form_round_off_grows_towards_z_equal_one(model, oracle):
    """Why the kernel keeps the delta form for the low channels: the all-direct model loses an
    order of magnitude on the filterbank output of the lowest group, the all-delta model does not."""
    from f2cnn_b200 import synth
    co = oracle.make_erb_filters(16000, oracle.centre_freqs(16000, 128, 100))
    w = synth.speech_like_i16(24000, seed=5)
    n, N2, xf, xi, G = ring_inputs(w)
    go = oracle.erb_filterbank(w, co)
    b, a = oracle.butter1_lowpass(50 / 8000.0)
    err = {}
    for form in (0, 1):
        gfb = np.empty((128, n), np.float32)
        model.f2m_run(xf.ctypes.data_as(_fp), xi.ctypes.data_as(_fp), G.ctypes.data_as(_fp), n, N2,
                      np.ascontiguousarray(co).ctypes.data_as(_dp), 128, 1, b[0], a[1], 1536, 2048, form,
                      gfb.ctypes.data_as(_fp), None)
        err[form] = np.max(np.abs(gfb - go), axis=1) / np.sqrt(np.mean(go ** 2, axis=1))
    assert err[0].max() <= 2e-5
    assert err[1][:64].max() <= 2e-5          # 1+B1+B2 >= 0.28: the direct form is as good
    assert err[1][96:].max() >= 4 * err[0][96:].max()
