"""Corpus pass (pre-pass + fused kernel, decimated frames), a few times: target of ncu one-liners."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import engine, synth
from f2cnn_b200.gammatone import filters
co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
plan = engine.plan_for(co)
lengths = synth.corpus_lengths(4620, 32000, 64000, seed=1)
flat, offs = synth.corpus_waves_i16(lengths, seed=1)
wave = torch.from_numpy(flat).cuda()
b = plan.batch(lengths, target_items=1)
dec = torch.empty((b.total_frames, 128), dtype=torch.float32, device="cuda")
ev = [(engine.DeviceEvent(), engine.DeviceEvent()) for _ in range(4)]
for i in range(4):
    b.run(wave, lpf=True, cutoff=50, out={"dec": dec}, fused_events=ev[i])
torch.cuda.synchronize()
print("fused ms", [round(x.elapsed_ms(y), 3) for x, y in ev])
