"""Print the measured parity of the CUDA path against the float64 oracle on the 3 s known-answer
inputs: max |error| / per-channel RMS for the filterbank output, the envelope (LPF on / off) and the
decimated frames, overall and per group of 32 channels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import engine, synth
from f2cnn_b200.gammatone import filters
from oracle import oracle

co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
plan = engine.plan_for(co)
cases = {"white": synth.white_noise_i16(48000, seed=0), "speech": synth.speech_like_i16(48000),
         "chirp": synth.chirp_i16(48000), "tone1k": synth.tone_i16(48000), "delta": synth.delta_i16(48000)}
def rel(got, want, floor=0.0):
    r = np.sqrt(np.mean(want ** 2, axis=1))
    r = np.maximum(r, floor * r.max())
    return np.max(np.abs(got - want), axis=1) / r
grp = lambda e: " ".join("%.1e" % e[g * 32:(g + 1) * 32].max() for g in range(4))
for name, w in cases.items():
    wd = torch.from_numpy(w).cuda()
    go = oracle.erb_filterbank(w, co)
    floor = 0.01 if name in ("tone1k", "chirp", "delta") else 0.0   # stop-band channels: relative to the loudest
    b = plan.batch([len(w)], target_items=1)
    r = b.run(wd, lpf=True, cutoff=50, gfb=torch.float64, env=torch.float64)
    gfb = r["gfb"].cpu().numpy().reshape(128, -1); env = r["env"].cpu().numpy().reshape(128, -1)
    eo = oracle.extract_envelope(go, True, 50)
    print("%-7s gfb+env run : gfb [%s]  env50 [%s]" % (name, grp(rel(gfb, go, floor)), grp(rel(env, eo, floor))))
    for lpf, cut in ((True, 50), (False, 100)):
        eo = oracle.extract_envelope(go, lpf, cut)
        env = b.run(wd, lpf=lpf, cutoff=cut, env=torch.float64)["env"].cpu().numpy().reshape(128, -1)
        dec = b.run(wd, lpf=lpf, cutoff=cut, dec=True)["dec"].cpu().numpy().T
        print("%-7s env-only lpf=%d: env [%s]  dec [%s]" % (name, lpf, grp(rel(env, eo, floor)), grp(rel(dec, eo[:, ::160], floor))))
