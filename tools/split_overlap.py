"""Resident corpus pass as P parts on P streams: the HBM-bound pre-pass of one part under the FMA-bound
fused kernel of another."""
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
for parts in (1, 2, 3, 4, 8):
    cuts = [int(round(len(lengths) * k / parts)) for k in range(parts + 1)]
    batches = [plan.batch(lengths[a:b], target_items=1) for a, b in zip(cuts[:-1], cuts[1:])]
    streams = [torch.cuda.Stream() for _ in batches]
    outs = []
    for b in batches:
        g, rows = b.grid_windows(11)
        outs.append((g, torch.empty((rows, 11, 128), dtype=torch.float32, device="cuda")))
    def step():
        cur = torch.cuda.current_stream()
        for (a, z), b, s, (g, w) in zip(zip(cuts[:-1], cuts[1:]), batches, streams, outs):
            s.wait_stream(cur)
            b.run(wave[offs[a]:offs[z]], lpf=True, cutoff=50, windows=(g, 11, w), stream=s)
        for s in streams:
            cur.wait_stream(s)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record(); torch.cuda.synchronize()
    print("parts %d: %.3f ms per corpus pass" % (parts, e0.elapsed_time(e1) / 5), flush=True)
    del outs, batches
    torch.cuda.empty_cache()
