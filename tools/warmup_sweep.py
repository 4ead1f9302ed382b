"""Development sweep: truncated-history lengths (w_imag, w_edge) vs error against the float64 oracle
and time, on one 3 s utterance (error) and a 592-utterance batch (time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import engine, synth
from f2cnn_b200.gammatone import filters
from oracle import oracle

co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
plan = engine.Plan(co)  # private plan: set_warmup is refused on the cached ones
print("defaults", plan.get_warmup())
cases = {"white": synth.white_noise_i16(48000, seed=0), "speech": synth.speech_like_i16(40000, seed=1)}
ref = {k: oracle.utterance(w, co, True, 50)[1] for k, w in cases.items()}
lens = synth.corpus_lengths(592, seed=1)
flat, _ = synth.corpus_waves_i16(lens, seed=1)
fd = torch.from_numpy(flat).cuda()
for wi, we in ((1536, 2048), (1280, 1536), (1280, 1280), (1024, 1280), (1024, 1024), (768, 1024), (768, 768), (512, 512)):
    plan.set_warmup(wi, we, 2048)
    errs = []
    for k, w in cases.items():
        b = plan.batch([len(w)], target_items=1)
        env = b.run(torch.from_numpy(w).cuda(), lpf=True, cutoff=50, env=torch.float64)["env"].cpu().numpy().reshape(128, -1)
        e = np.max(np.abs(env - ref[k]), axis=1) / np.sqrt(np.mean(ref[k] ** 2, axis=1))
        errs.append("%s %.2e (ch %d)" % (k, e.max(), int(e.argmax())))
    bb = plan.batch(lens)
    dec = torch.empty((bb.total_frames, 128), device="cuda")
    for _ in range(2):
        bb.run(fd, lpf=True, cutoff=50, out={"dec": dec})
    torch.cuda.synchronize()
    a, z = engine.DeviceEvent(), engine.DeviceEvent()
    a.record()
    for _ in range(5):
        bb.run(fd, lpf=True, cutoff=50, out={"dec": dec})
    z.record()
    torch.cuda.synchronize()
    print("w_imag %4d w_edge %4d : %s : %.3f ms" % (wi, we, "  ".join(errs), a.elapsed_ms(z) / 5), flush=True)
