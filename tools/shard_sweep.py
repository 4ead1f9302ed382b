"""How the fused path behaves on shard-sized batches (config 3: 4620 utterances over 1/2/4/8 GPUs):
device time of prep + fused (decimated frames) for U utterances of the seed-1 corpus, whole
utterances (target_items=1) against the automatic time-chunking policy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import engine, synth
from f2cnn_b200.gammatone import filters

co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
plan = engine.plan_for(co)
lengths_all = synth.corpus_lengths(4620, 32000, 64000, seed=1)
flat_all, offs = synth.corpus_waves_i16(lengths_all, seed=1)
for world in [int(w) for w in os.environ.get('SHARD_WORLDS', '1,2,4,8,16,32,64').split(',')]:
    idx = engine.shard_utterances(lengths_all, world)[0]
    lengths = lengths_all[idx]
    wave = torch.from_numpy(np.concatenate([flat_all[offs[u]:offs[u + 1]] for u in idx])).cuda()
    for target in ((1,) if os.environ.get('SHARD_WHOLE_ONLY') else (1, 0)):
        b = plan.batch(lengths, target_items=target)
        dec = torch.empty((b.total_frames, 128), dtype=torch.float32, device="cuda")
        ev = [(engine.DeviceEvent(), engine.DeviceEvent()) for _ in range(5)]
        for _ in range(2):
            b.run(wave, lpf=True, cutoff=50, out={"dec": dec})
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(5):
            b.run(wave, lpf=True, cutoff=50, out={"dec": dec}, fused_events=ev[i])
        z.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(z) / 5
        fused = np.mean([x.elapsed_ms(y) for x, y in ev])
        cs = 128.0 * lengths.sum()
        print("shard 1/%-2d %4d utts target_items=%d items %5d: step %.3f ms fused %.3f ms  %.3e ch-samples/s  (x%d = %.3e)"
              % (world, len(idx), target, b.num_items, ms, fused, cs / ms * 1e3, world, world * cs / ms * 1e3), flush=True)
