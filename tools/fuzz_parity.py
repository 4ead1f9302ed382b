"""One-off wide fuzz of the CUDA path against the float64 oracle (the test suite runs a 16-case
version): random banks (fs, channels, LOW_FREQ, width), lengths, input kinds and dtypes, cut-offs,
decimation grids, chunk targets.  Prints the worst error / (1e-4 x channel RMS) seen per output."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import engine, synth
from f2cnn_b200.gammatone import filters
from oracle import oracle

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
worst = {"gfb": (0, None), "env": (0, None), "dec": (0, None)}
t0 = time.time()
for case in range(cases):
    fs = int(rng.choice([8000, 16000, 16000, 22050, 44100]))
    C = int(rng.choice([1, 7, 32, 40, 64, 96, 128, 160, 256]))
    low = int(rng.choice([20, 50, 100, 100, 300]))
    width = float(rng.choice([1.0, 1.0, 0.5, 2.0]))
    co = filters.make_erb_filters(fs, filters.centre_freqs(fs, C, low), width)
    n = int(rng.choice([300, 1000, 4096, 8191, 20000, 48000, 65536, 65537, 100000])) + int(rng.integers(0, 50))
    kind = str(rng.choice(["white", "speech", "tone", "chirp"]))
    if kind == "white":
        w = synth.white_noise_i16(n, seed=case)
    elif kind == "speech":
        w = synth.speech_like_i16(n, seed=case)
    elif kind == "tone":
        w = synth.tone_i16(n, freq=float(rng.choice([low, 2 * low, 440.0, 1000.0, fs / 4.0])), fs=fs)
    else:
        w = synth.chirp_i16(n, f0=float(low), f1=fs * 0.45, fs=fs)
    dtype = [np.int16, np.float32, np.float64][int(rng.integers(0, 3))]
    w = w.astype(dtype)
    lpf = bool(rng.random() < 0.7)
    cutoff = float(rng.choice([20, 50, 100, 400]))
    step = int(rng.choice([160, 80, 441]))
    target = int(rng.choice([0, 0, 1, 64]))
    try:
        plan = engine.plan_for(co)
    except Exception as e:  # banks the float32 guard of f2_plan_create refuses
        print("case %d refused: %s" % (case, str(e)[:120]))
        continue
    b = plan.batch([n], step=step, target_items=target)
    wd = torch.from_numpy(w).cuda()
    r = b.run(wd, lpf=lpf, cutoff=cutoff, gfb=torch.float64, env=torch.float64)
    dec = b.run(wd, lpf=lpf, cutoff=cutoff, dec=True)["dec"].cpu().numpy().T
    go = oracle.erb_filterbank(w, co)
    eo = oracle.extract_envelope(go, lpf, cutoff)
    gfb = r["gfb"].cpu().numpy().reshape(C, n)
    env = r["env"].cpu().numpy().reshape(C, n)
    # scale: channel RMS, floored at 1 % of the loudest channel (stop-band channels of tonal inputs)
    sg = np.sqrt(np.mean(go ** 2, axis=1)); sg = np.maximum(sg, 0.01 * sg.max())
    se = np.sqrt(np.mean(eo ** 2, axis=1)); se = np.maximum(se, 0.01 * se.max())
    tag = dict(case=case, fs=fs, C=C, low=low, width=width, n=n, kind=kind, dtype=dtype.__name__, lpf=lpf,
               cutoff=cutoff, step=step, target=target, items=b.num_items)
    for name, got, want, sc in (("gfb", gfb, go, sg), ("env", env, eo, se), ("dec", dec, eo[:, ::step], se)):
        e = np.max(np.abs(got - want), axis=1) / sc
        v = float(e.max()) / 1e-4
        if v > worst[name][0]:
            worst[name] = (v, dict(tag, channel=int(e.argmax())))
print("cases", cases, "seconds %.0f" % (time.time() - t0))
for k, (v, tag) in worst.items():
    print("%s worst = %.3f of the 1e-4 bar at %s" % (k, v, tag))
