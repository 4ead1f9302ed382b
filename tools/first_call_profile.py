"""Where the FIRST api.features_to_windows call of a process spends its time (cProfile, top entries)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
from f2cnn_b200 import api, engine, synth
from f2cnn_b200.gammatone import filters

co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
lengths = synth.corpus_lengths(4620, 32000, 64000, seed=1)
flat, _ = synth.corpus_waves_i16(lengths, seed=1)
t = time.perf_counter(); wave = torch.from_numpy(flat).pin_memory(); print("pin 444 MB of waves: %.1f ms" % ((time.perf_counter() - t) * 1e3))
nwin = np.maximum((lengths / 160 - 12).astype(np.int64), 0)
centers = np.concatenate([800 + 160 * np.arange(k, dtype=np.int64) for k in nwin])
t = time.perf_counter(); out = engine.host_empty((int(nwin.sum()), 11, 128), np.float32); print("host_empty 7.5 GB: %.1f ms" % ((time.perf_counter() - t) * 1e3))
pr = cProfile.Profile()
t = time.perf_counter()
pr.enable()
api.features_to_windows((wave, lengths), co, centers, True, 50, 5, 160, out=out, counts=nwin)
pr.disable()
print("first call: %.1f ms" % ((time.perf_counter() - t) * 1e3))
t = time.perf_counter()
api.features_to_windows((wave, lengths), co, centers, True, 50, 5, 160, out=out, counts=nwin)
print("second call: %.1f ms" % ((time.perf_counter() - t) * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
