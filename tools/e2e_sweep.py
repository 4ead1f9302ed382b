"""How many sub-batches should the host-to-host window pipeline use?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import engine, synth
from f2cnn_b200.gammatone import filters
co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
lengths = synth.corpus_lengths(4620, seed=1)
flat, offsets = synth.corpus_waves_i16(lengths, seed=1)
plan = engine.plan_for(co)
nwin = np.maximum((lengths / 160 - 12).astype(np.int64), 0)
wave_host = torch.from_numpy(flat).pin_memory()
out_host = torch.empty((int(nwin.sum()), 11, 128), dtype=torch.float32, pin_memory=True)
for n_sub in (4, 8, 12, 16, 24, 48):
    pipe = engine.WindowPipeline(plan, lengths, [np.arange(k, dtype=np.int64) for k in nwin], n_sub=n_sub)
    for _ in range(2):
        pipe.run(wave_host, out_host)
    torch.cuda.synchronize()
    a, b = engine.DeviceEvent(), engine.DeviceEvent()
    a.record()
    for _ in range(4):
        pipe.run(wave_host, out_host)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_ms(b) / 4
    print("n_sub %2d: %.1f ms/step, D2H-equivalent %.1f GB/s" % (n_sub, ms, out_host.numel() * 4 / ms / 1e6), flush=True)
    del pipe
