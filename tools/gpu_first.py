"""First-contact diagnostics on the GPU box: error tables of the CUDA path vs the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import api, synth, engine
from f2cnn_b200.gammatone import filters
from oracle import oracle as orc

def rel(got, want):
    r = np.sqrt(np.mean(want ** 2, axis=1))
    e = np.max(np.abs(got - want), axis=1) / np.maximum(r, 1e-300)
    return e

coefs = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
for name, w in (("white48000", synth.white_noise_i16(48000, 0)), ("white1000", synth.white_noise_i16(1000, 1)),
                ("white65536", synth.white_noise_i16(65536, 2)), ("white100", synth.white_noise_i16(100, 3)),
                ("speech40000", synth.speech_like_i16(40000))):
    gfb_o = orc.erb_filterbank(w, coefs)
    for lpf, cut in ((True, 50), (False, 100)):
        env_o = orc.extract_envelope(gfb_o, lpf, cut)
        t = time.time()
        gfb, env = api.filterbank_envelope(w, coefs, lpf, cut, with_gfb=True)
        dt = time.time() - t
        eg, ee = rel(gfb, gfb_o), rel(env, env_o)
        print("%-12s lpf=%d gfb max %.2e (ch %d) env max %.2e (ch %d)  nan=%d  %.1f ms" % (
            name, lpf, eg.max(), eg.argmax(), ee.max(), ee.argmax(), int(np.isnan(env).sum()), dt * 1e3), flush=True)
    # stand-alone rows path
    env_o = orc.extract_envelope(gfb_o, True, 50)
    env_r = api.extract_envelope_from_matrix(gfb_o, True, 50)
    er = rel(env_r, env_o)
    print("%-12s rows-envelope max %.2e (ch %d)" % (name, er.max(), er.argmax()), flush=True)
    # gfb only
    g2 = api.erb_filterbank(w, coefs)
    print("%-12s erb_filterbank max %.2e" % (name, rel(g2, gfb_o).max()), flush=True)
# chunked single utterance
w = synth.white_noise_i16(48000, 0)
plan = engine.plan_for(coefs)
gfb_o = orc.erb_filterbank(w, coefs); env_o = orc.extract_envelope(gfb_o, True, 50)
for target in (1, 8, 64, 4096):
    b = plan.batch([48000], target_items=target)
    res = b.run(torch.from_numpy(w).cuda(), lpf=True, cutoff=50, env=torch.float64, dec=True)
    env = res["env"].cpu().numpy().reshape(128, 48000)
    dec = res["dec"].cpu().numpy()
    ee = rel(env, env_o)
    ed = np.max(np.abs(dec.T - env_o[:, ::160]), axis=1) / np.sqrt(np.mean(env_o ** 2, axis=1))
    print("target_items %5d items %4d env max %.2e dec max %.2e" % (target, b.num_items, ee.max(), ed.max()), flush=True)
# windows
centers = synth.label_grid(48000)
win = api.features_to_windows([w], coefs, [centers], True, 50)
win_o = orc.gather_windows(env_o, centers)
print("windows", win.shape, "max err / rms %.2e" % (np.abs(win - win_o).max() / np.sqrt(np.mean(env_o ** 2))))
