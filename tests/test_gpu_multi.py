"""Utterance sharding over the GPUs of one box (needs >= 2 devices; skipped otherwise): the
single-process MultiGpuWindowPipeline and, under torchrun, `bench.py --gpus N` (which checks the
tensor the ranks assembled against the 1-GPU result itself).  The host-side logic of both runs on
CPU in tests/test_host_side_cpu.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _corpus(n_utts=64, seed=9):
    from f2cnn_b200 import synth
    lengths = synth.corpus_lengths(n_utts, lo=8000, hi=20000, seed=seed)
    flat, _ = synth.corpus_waves_i16(lengths, seed=seed)
    cent = [synth.label_grid(int(n)) for n in lengths]
    return lengths, flat, np.concatenate(cent), np.asarray([len(c) for c in cent], dtype=np.int64)


def test_multi_gpu_window_pipeline_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from f2cnn_b200 import engine
    from f2cnn_b200.gammatone import filters
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
    lengths, flat, centers, counts = _corpus()
    wave_host = torch.from_numpy(flat).pin_memory()
    n_win = int(counts.sum())
    one = engine.WindowPipeline(engine.plan_for(co, 0), lengths, n_sub=4)
    runs, _, _ = engine.window_runs(centers, counts, lengths, one.frame_offsets)
    out1 = np.zeros((n_win, 11, 128), dtype=np.float32)
    one.run(wave_host, runs, out1)
    multi = engine.MultiGpuWindowPipeline(co, lengths, n_sub=2)
    assert len(multi.parts) == torch.cuda.device_count()
    out2 = np.zeros((n_win, 11, 128), dtype=np.float32)
    multi.run(wave_host, multi.window_runs(centers, counts), out2)
    assert np.array_equal(out1, out2)


def test_api_shards_fill_one_tensor_like_the_single_call():
    """features_to_windows(shard=(r, w)) for every r, each on its own device, into one array."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from f2cnn_b200 import api
    from f2cnn_b200.gammatone import filters
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
    lengths, flat, centers, counts = _corpus(48, seed=4)
    whole = np.zeros((int(counts.sum()), 11, 128), np.float32)
    api.features_to_windows((flat, lengths), co, centers, True, 50, counts=counts, out=whole, shard=(0, 1))
    parts = np.zeros_like(whole)
    world = torch.cuda.device_count()
    for r in range(world):
        with torch.cuda.device(r):
            api.features_to_windows((flat, lengths), co, centers, True, 50, counts=counts, out=parts, shard=(r, world))
    assert np.array_equal(whole, parts)


def test_bench_under_torchrun_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    port = 29710 + os.getpid() % 100
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2",
                        "--warmup", "1", "--utts", "256", "--no-cpu"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    assert line["e2e"]["equals_single_gpu_result"] is True
