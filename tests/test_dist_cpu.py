"""World-size-2 gloo test of the multi-process sharding logic used by bench.py (no GPU):
every rank owns a corpus shard, there is no data-path collective, the only exchange is the
MAX-reduce of the per-rank time and the row offsets are a prefix sum known before launch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from f2cnn_b200 import synth
    lengths = synth.corpus_lengths(40, seed=1 + rank)  # bench.py: per-rank shard, seed 1 + rank
    nwin = np.maximum((lengths / 160 - 12).astype(np.int64), 0)
    # row offsets of this rank's block in a gathered tensor: exclusive prefix sum of window counts
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([int(nwin.sum())]))
    offset = int(sum(int(c) for c in counts[:rank]))
    # the timing reduction bench.py does
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ret[rank] = (int(lengths.sum()), int(nwin.sum()), offset, float(t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29650 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert set(ret.keys()) == {0, 1}
    (s0, w0, o0, t0), (s1, w1, o1, t1) = ret[0], ret[1]
    assert s0 != s1                      # different shards
    assert o0 == 0 and o1 == w0          # contiguous, non-overlapping row blocks
    assert t0 == t1 == 11.0              # max over ranks


def test_bench_reference_arm_other_ranks_exit_quietly():
    import subprocess
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
