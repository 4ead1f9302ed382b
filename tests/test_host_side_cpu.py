"""Host half of the window stage (f2_host.cpp) and the sharding logic of the multi-GPU path, on
CPU: none of this needs a device -- placement copies float32 bit patterns, run detection and
sharding are integer work.  Reference semantics: scripts/processing/InputGenerator.py:73-80 (rows),
LabelDataGenerator.py:44-50 (grid)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames(total, C=128, seed=0):
    return np.random.default_rng(seed).standard_normal((total, C)).astype(np.float32)


def _reference_rows(frames, frame_offsets, centers, counts, radius, step, phase=0):
    """Row by row, the way InputGenerator.py:73-80 indexes (on decimated frames)."""
    rows = []
    pos = 0
    for u, k in enumerate(counts):
        for c in centers[pos:pos + k]:
            idx = [(c + step * (j - radius) - phase) // step for j in range(2 * radius + 1)]
            rows.append(frames[frame_offsets[u] + np.asarray(idx)])
        pos += k
    return np.stack(rows) if rows else np.zeros((0, 2 * radius + 1, frames.shape[1]), np.float32)


def _corpus(lengths, step=160, radius=5, phase=0):
    from f2cnn_b200 import synth
    lengths = np.asarray(lengths, dtype=np.int64)
    n_dec = np.where(lengths > phase, (lengths - phase + step - 1) // step, 0)
    fo = np.concatenate([[0], np.cumsum(n_dec)]).astype(np.int64)
    cent = [synth.label_grid(int(n), radius, step) + phase for n in lengths]
    cent = [c[c + radius * step < n] for c, n in zip(cent, lengths)]
    return lengths, fo, cent


def test_grid_runs_and_placement_equal_the_reference_loop():
    from f2cnn_b200 import engine
    lengths, fo, cent = _corpus([48000, 1000, 3000, 64000, 1759, 1761, 0, 33333])
    counts = [len(c) for c in cent]
    centers = np.concatenate(cent)
    runs, phase, n_rows = engine.window_runs(centers, counts, lengths, fo)
    assert phase == 0 and n_rows == sum(counts)
    assert runs.shape[0] == sum(1 for k in counts if k)  # the label grid: ONE run per utterance
    frames = _frames(fo[-1])
    want = _reference_rows(frames, fo, centers, counts, 5, 160)
    for threads in (1, 3, 0):
        out = np.full((n_rows, 11, 128), np.nan, dtype=np.float32)
        engine.place_windows(frames, runs, out, 11, threads)
        assert np.array_equal(out, want)


def test_placement_other_shapes_and_misaligned_output():
    from f2cnn_b200 import engine
    rng = np.random.default_rng(3)
    for C, dots in ((128, 11), (40, 5), (1, 3), (7, 21), (256, 11)):
        frames = _frames(500, C, seed=C)
        k = int(rng.integers(1, 12))
        runs = []
        row = 0
        for _ in range(k):
            cnt = int(rng.integers(0, 40))
            first = int(rng.integers(0, 500 - cnt - dots + 1))
            runs.append((first, row, cnt))
            row += cnt
        runs = np.asarray(runs, dtype=np.int64)
        # output that starts 4 bytes off any vector alignment: head / tail handling of the stores
        backing = np.zeros(row * dots * C + 1, dtype=np.float32)
        out = backing[1:].reshape(row, dots, C)
        engine.place_windows(frames, runs, out, dots, 2)
        want = np.stack([frames[f + i:f + i + dots] for f, r, c in runs for i in range(c)]) if row else out
        assert np.array_equal(out, want) and backing[0] == 0
    with pytest.raises(IndexError):
        engine.place_windows(_frames(20), np.asarray([[15, 0, 2]]), np.zeros((2, 11, 128), np.float32))
    with pytest.raises(IndexError):
        engine.place_windows(_frames(40), np.asarray([[0, 1, 2]]), np.zeros((2, 11, 128), np.float32))
    with pytest.raises(TypeError):
        engine.place_windows(_frames(40).astype(np.float64), np.asarray([[0, 0, 2]]), np.zeros((2, 11, 128), np.float32))


def test_runs_follow_csv_order_and_python_index_semantics():
    from f2cnn_b200 import engine
    lengths, fo, cent = _corpus([20000, 9000])
    rng = np.random.default_rng(5)
    shuffled = [rng.permutation(c) for c in cent]      # a hand-made CSV: any order, repeats allowed
    shuffled[1] = np.concatenate([shuffled[1], shuffled[1][:3]])
    counts = [len(c) for c in shuffled]
    centers = np.concatenate(shuffled)
    runs, phase, n_rows = engine.window_runs(centers, counts, lengths, fo)
    assert runs is not None and runs[:, 2].sum() == n_rows
    frames = _frames(fo[-1], seed=2)
    out = np.zeros((n_rows, 11, 128), np.float32)
    engine.place_windows(frames, runs, out)
    assert np.array_equal(out, _reference_rows(frames, fo, centers, counts, 5, 160))
    # off-grid but consistent phase
    lengths, fo, cent = _corpus([20000, 9000], phase=37)
    centers, counts = np.concatenate(cent), [len(c) for c in cent]
    runs, phase, _ = engine.window_runs(centers, counts, lengths, fo)
    assert phase == 37 and runs.shape[0] == 2
    out = np.zeros((sum(counts), 11, 128), np.float32)
    engine.place_windows(_frames(fo[-1], seed=4), runs, out)
    assert np.array_equal(out, _reference_rows(_frames(fo[-1], seed=4), fo, centers, counts, 5, 160, 37))
    # a window that wraps like a negative Python index, or mixed phases: legal, but not runs
    assert engine.window_runs([100], [1], [48000], [0])[0] is None
    assert engine.window_runs([800, 961], [2], [48000], [0])[0] is None
    # beyond the end / below -n: the reference raises IndexError (InputGenerator.py:76)
    with pytest.raises(IndexError):
        engine.window_runs([47900], [1], [48000], [0])
    with pytest.raises(IndexError):
        engine.window_runs([400], [1], [350], [0])
    with pytest.raises(ValueError):
        engine.window_runs([800, 960], [1], [48000], [0])


def test_async_placer_without_a_stream_and_huge_page_arrays():
    from f2cnn_b200 import engine
    lengths, fo, cent = _corpus([30000] * 40)
    counts = [len(c) for c in cent]
    runs, _, n_rows = engine.window_runs(np.concatenate(cent), counts, lengths, fo)
    frames = _frames(fo[-1], seed=9)
    want = np.zeros((n_rows, 11, 128), np.float32)
    engine.place_windows(frames, runs, want, threads=1)
    placer = engine.Placer(3)
    assert placer.threads == 3
    out = engine.host_empty((n_rows, 11, 128), np.float32)
    assert out.flags["C_CONTIGUOUS"] and out.flags["WRITEABLE"]
    for lo in range(0, runs.shape[0], 7):          # many small jobs queued back to back
        placer.submit(frames, runs[lo:lo + 7], out, after_stream=False)
    placer.wait()
    assert np.array_equal(out, want)
    placer.wait()                                   # idempotent
    keep = out[5].copy()
    del placer
    assert np.array_equal(out[5], keep)


def test_shard_utterances_is_a_balanced_partition():
    from f2cnn_b200 import engine, synth
    lengths = synth.corpus_lengths(4620, seed=1)
    for world in (1, 2, 4, 8):
        shards = engine.shard_utterances(lengths, world)
        assert len(shards) == world
        allidx = np.concatenate(shards)
        assert np.array_equal(np.sort(allidx), np.arange(4620))        # a partition
        assert all(np.all(np.diff(s) > 0) for s in shards)             # ascending: rows stay in corpus order
        sums = np.asarray([lengths[s].sum() for s in shards])
        assert sums.max() - sums.min() <= lengths.max()                # length-sorted round-robin
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    assert [list(s) for s in engine.shard_utterances([5, 9, 7], 8)][:3] == [[1], [2], [0]]


# ---- world-size-2 run of the multi-process path (gloo, no GPU) ------------------------------------
def _rank_main(rank, world, port, path, ok):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from f2cnn_b200 import engine, hostmem, synth
    lengths = synth.corpus_lengths(60, lo=4000, hi=12000, seed=1)       # every rank: the SAME corpus
    step, radius = 160, 5
    cent = [synth.label_grid(int(n)) for n in lengths]
    counts = np.asarray([len(c) for c in cent], dtype=np.int64)
    rows = np.concatenate([[0], np.cumsum(counts)])                    # prefix sum known before launch
    n_dec = (lengths + step - 1) // step
    fo_all = np.concatenate([[0], np.cumsum(n_dec)])
    frames_all = (np.arange(fo_all[-1] * 128, dtype=np.float32).reshape(-1, 128) * 0.25)  # stands for the kernel's output
    shared = hostmem.SharedArray(path, (int(rows[-1]), 11, 128), create=(rank == 0)) if rank == 0 else None
    dist.barrier()
    if shared is None:
        shared = hostmem.SharedArray(path, (int(rows[-1]), 11, 128))
    mine = engine.shard_utterances(lengths, world)[rank]
    # this rank's frames, numbered over ITS utterances (what its WindowPipeline downloads)
    fo = np.concatenate([[0], np.cumsum(n_dec[mine])])
    frames = np.concatenate([frames_all[fo_all[u]:fo_all[u + 1]] for u in mine])
    runs, _, n_rows = engine.window_runs(np.concatenate([cent[u] for u in mine]), counts[mine], lengths[mine], fo,
                                         radius, step, 0, row_offsets=rows[mine])
    assert n_rows == counts[mine].sum()
    engine.place_windows(frames, runs, shared.array, 11, 2)             # rows land at their FINAL offsets
    dist.barrier()
    if rank == 0:
        runs1, _, _ = engine.window_runs(np.concatenate(cent), counts, lengths, fo_all, radius, step, 0)
        want = np.zeros_like(shared.array)
        engine.place_windows(frames_all, runs1, want, 11, 2)
        ok.value = int(np.array_equal(shared.array, want))
        shared.unlink()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_place_their_shards_into_one_shared_buffer(tmp_path):
    import torch.multiprocessing as mp
    from f2cnn_b200 import hostmem
    d = hostmem.shm_dir() or str(tmp_path)
    path = os.path.join(d, "f2cnn_b200_test_%d.bin" % os.getpid())
    ok = mp.get_context("spawn").Value("i", 0)
    port = 29650 + os.getpid() % 200
    mp.spawn(_rank_main, args=(2, port, path, ok), nprocs=2, join=True)
    assert ok.value == 1
    assert not os.path.exists(path)


def test_bench_reference_arm_other_ranks_exit_quietly():
    import subprocess
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_frame_windows_views_getitem_and_materialize_on_the_host():
    """api.FrameWindows is host logic over a frames array: utterance blocks as strided views, global rows,
    and materialize() through the library's placement pool -- against plain numpy indexing."""
    from f2cnn_b200 import api, engine
    rng = np.random.default_rng(11)
    n_frames = np.array([40, 12, 11, 0, 25], dtype=np.int64)          # 12 frames: 2 windows; 11: exactly 1
    counts = np.maximum(n_frames - 10, 0)
    counts[1] = 1                                                       # fewer windows than the frames allow
    frame_offsets = np.concatenate([[0], np.cumsum(n_frames)]).astype(np.int64)
    frames = rng.standard_normal((int(frame_offsets[-1]), 128)).astype(np.float32)
    fw = api.FrameWindows(frames, frame_offsets, counts, 11, np.arange(5), engine.Placer(3))
    want = np.concatenate([np.stack([frames[frame_offsets[u] + k:frame_offsets[u] + k + 11] for k in range(int(c))])
                           for u, c in enumerate(counts) if c > 0])
    assert len(fw) == want.shape[0] == int(counts.sum())
    row = 0
    for u, c in enumerate(counts):
        v = fw.windows(u)
        assert v.shape == (c, 11, 128) and np.array_equal(v, want[row:row + int(c)])
        assert c == 0 or np.shares_memory(v, frames)
        row += int(c)
    for r in list(range(len(fw))) + [-1, -len(fw)]:
        assert np.array_equal(fw[r], want[r])
    for bad in (len(fw), -len(fw) - 1):
        with pytest.raises(IndexError):
            fw[bad]
    assert np.array_equal(fw.materialize(), want)
    out = np.zeros_like(want)
    assert fw.materialize(out) is out and np.array_equal(out, want)
    with pytest.raises(ValueError):
        fw.windows(0)[0, 0, 0] = 1.0                                    # views are read-only
