"""Just the tcgen05 CNN forward on one 3 s utterance (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import cnn, engine, synth
from f2cnn_b200.gammatone import filters
co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
plan = engine.plan_for(co)
n = 48000
w = synth.speech_like_i16(n, seed=3).astype(np.float64) + np.random.default_rng(1).normal(0, 30, n)
env_t = plan.batch([n]).run(torch.from_numpy(w).cuda(), lpf=True, cutoff=50, env_t=True)["env_t"]
tc = cnn.TensorCoreCNN(cnn.seeded_model(seed=0))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    tc.predict_envelope(env_t, 160)
torch.cuda.synchronize()
