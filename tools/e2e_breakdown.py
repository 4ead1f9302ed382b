"""Where the end-to-end time of a corpus-sized api.features_to_windows call goes: run detection,
GPU chain alone (H2D + kernels + D2H of the frames, no placement), placement alone (frames already on
the host), and the overlapped whole, by sub-batch count and placement-thread count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import api, engine, synth
from f2cnn_b200.gammatone import filters

co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
plan = engine.plan_for(co)
lengths = synth.corpus_lengths(4620, 32000, 64000, seed=1)
flat, offs = synth.corpus_waves_i16(lengths, seed=1)
wave_host = torch.from_numpy(flat).pin_memory()
nwin = np.maximum((lengths / 160 - 12).astype(np.int64), 0)
centers = np.concatenate([800 + 160 * np.arange(k, dtype=np.int64) for k in nwin])
N = int(nwin.sum())
out = engine.host_empty((N, 11, 128), np.float32)
out[:] = 0


def clock(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


class NoPlacer:
    threads = 0
    def submit(self, *a, **k): pass
    def wait(self): pass


for n_sub in (None, 4, 8, 16, 24, 32):
    pipe = engine.WindowPipeline(plan, lengths, n_sub=n_sub, placer=NoPlacer())
    t_runs = clock(lambda: engine.window_runs(centers, nwin, lengths, pipe.frame_offsets))
    runs, _, _ = engine.window_runs(centers, nwin, lengths, pipe.frame_offsets)
    t_gpu = clock(lambda: pipe.run(wave_host, runs, out))
    print("n_sub=%s -> %d sub-batches: window_runs %.2f ms, GPU chain alone (H2D+kernels+D2H frames) %.2f ms" %
          (n_sub, len(pipe.subs), t_runs, t_gpu), flush=True)
    frames = pipe.frames_host
    for threads in (12, 15, 16):
        placer = engine.Placer(threads)
        t_place = clock(lambda: (placer.submit(frames, runs, out, after_stream=False), placer.wait()))
        pipe.placer = placer
        t_all = clock(lambda: pipe.run(wave_host, runs, out))
        print("   %2d placement threads: placement alone %.2f ms (%.0f GB/s), whole pipeline %.2f ms" %
              (threads, t_place, out.nbytes / t_place / 1e6, t_all), flush=True)
        pipe.placer = NoPlacer()
        del placer
# a shard of an 8-GPU run
idx = engine.shard_utterances(lengths, 8)[0]
cum = np.concatenate([[0], np.cumsum(lengths)])
mine = np.zeros(len(lengths), bool); mine[idx] = True
for n_sub in (1, 2, 3, 4):
    pipe = engine.WindowPipeline(plan, lengths[idx], n_sub=n_sub, src_offsets=cum[idx], placer=engine.Placer(16))
    runs, _, _ = engine.window_runs(centers[np.repeat(mine, nwin)], nwin[idx], lengths[idx], pipe.frame_offsets,
                                    row_offsets=(np.cumsum(nwin) - nwin)[idx])
    print("1/8 shard, n_sub=%d -> %d sub-batches: %.2f ms" % (n_sub, len(pipe.subs), clock(lambda: pipe.run(wave_host, runs, out))), flush=True)
t_api = clock(lambda: api.features_to_windows((wave_host, lengths), co, centers, True, 50, out=out, counts=nwin))
print("api.features_to_windows warm: %.2f ms" % t_api)
import gc
def fresh():
    r = api.features_to_windows((wave_host, lengths), co, centers, True, 50, counts=nwin)
    del r
    gc.collect()
print("api.features_to_windows warm, fresh output each call (pooled block): %.2f ms" % clock(fresh))
