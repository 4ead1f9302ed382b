"""Utterance sharding over the GPUs of one box from a single process (needs >= 2 devices;
skipped otherwise).  The multi-process path is exercised by `bench.py --gpus N` under torchrun
and, on CPU, by tests/test_dist_cpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_multi_gpu_window_pipeline_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from f2cnn_b200 import engine, synth
    from f2cnn_b200.gammatone import filters
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
    lengths = synth.corpus_lengths(64, lo=8000, hi=20000, seed=9)
    flat, _ = synth.corpus_waves_i16(lengths, seed=9)
    bases = [np.arange(max(int(n / 160 - 12), 0), dtype=np.int64) for n in lengths]
    wave_host = torch.from_numpy(flat).pin_memory()
    n_win = sum(len(b) for b in bases)
    one = engine.WindowPipeline(engine.plan_for(co, 0), lengths, bases, n_sub=4)
    out1 = torch.empty((n_win, 11, 128), dtype=torch.float32, pin_memory=True)
    one.run(wave_host, out1)
    torch.cuda.synchronize(0)
    multi = engine.MultiGpuWindowPipeline(co, lengths, bases, n_sub=2)
    assert len(multi.parts) == torch.cuda.device_count() and multi.n_windows == n_win
    out2 = torch.zeros((n_win, 11, 128), dtype=torch.float32, pin_memory=True)
    multi.run(wave_host, out2)
    assert torch.equal(out1, out2)
