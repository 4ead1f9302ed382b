"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import api, engine, synth
from f2cnn_b200.gammatone import filters
co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 40, 100))
w = synth.white_noise_i16(9000, 1)
g, e = api.filterbank_envelope(w, co, True, 50, with_gfb=True)
api.extract_envelope_from_matrix(g[:8], True, 50)
api.features_to_windows([w, synth.white_noise_i16(100, 2), synth.white_noise_i16(70000, 3)], co,
                        [np.array([800, 960]), np.zeros(0, dtype=np.int64), synth.label_grid(70000)[:20]], True, 50)
api.features_to_windows([w], co, [np.array([801, 1000])], False)
api.dense_frames(w, co, True, 50, normalize=True, frames=(0, 64))
plan = engine.plan_for(co)
b = plan.batch([9000, 3000, 40000], target_items=64)
b.run(torch.from_numpy(np.concatenate([w, w[:3000], synth.white_noise_i16(40000, 5)])).cuda(), lpf=True, cutoff=20, dec=True,
      env_t=True)
os.environ["F2_USE_LANES"] = "1"
torch.cuda.synchronize()
print("sanitize case OK")
