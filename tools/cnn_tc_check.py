"""tcgen05 CNN forward against the float32 PyTorch oracle, layer by layer, and its time per 3 s utterance."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from f2cnn_b200 import cnn, engine, synth
from f2cnn_b200.gammatone import filters

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
plan = engine.plan_for(co)
n = 48000
w = synth.speech_like_i16(n, seed=3).astype(np.float64) + np.random.default_rng(1).normal(0, 30, n)
env_t = plan.batch([n]).run(torch.from_numpy(w).cuda(), lpf=True, cutoff=50, env_t=True)["env_t"]
model = cnn.seeded_model(seed=0)
tc = cnn.TensorCoreCNN(model)
nb = n - 1760

def oracle(i0, i1):
    idx = torch.arange(i0, i1, device="cuda")[:, None] + 160 * torch.arange(11, device="cuda")[None, :]
    fr = env_t[idx].double()                                   # (m, 11, 128)
    lo = fr.amin(dim=(1, 2), keepdim=True).log(); hi = fr.amax(dim=(1, 2), keepdim=True).log()
    x = ((fr.log() - lo) / (hi - lo)).float().unsqueeze(1)
    a1 = F.relu(model.c1(x)); p2 = F.max_pool2d(F.relu(model.c2(a1)), 2)
    a3 = F.relu(model.c3(p2)); p4 = F.max_pool2d(F.relu(model.c4(a3)), 2)
    feat = p4.permute(0, 2, 3, 1).reshape(p4.shape[0], -1)
    return p2, feat, F.softmax(model.d2(F.relu(model.d1(feat))), dim=1)

m = 600
with torch.no_grad():
    p2, feat, sc = oracle(1000, 1000 + m)
got = tc.predict_envelope(env_t, 160, frames=(1000, 1000 + m))
torch.cuda.synchronize()
gp2, gfeat = tc.intermediates(m)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
print("pooled conv2: max err / max %.3e   features: %.3e   scores: max abs diff %.3e" % (rel(gp2, p2), rel(gfeat, feat), float((got - sc).abs().max())))
print("argmax agreement %.4f (margin > 0.05: %.4f)" % (float((got.argmax(1) == sc.argmax(1)).float().mean()),
      float(((got.argmax(1) == sc.argmax(1)) | ((sc[:, 0] - sc[:, 1]).abs() < 0.05)).float().mean())))
# whole utterance, timed
for _ in range(2):
    s_all = tc.predict_envelope(env_t, 160)
torch.cuda.synchronize()
a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    s_all = tc.predict_envelope(env_t, 160)
z.record(); torch.cuda.synchronize()
ms = a.elapsed_time(z) / 5
print("tcgen05 forward, %d frames: %.3f ms  (%.1f TFLOP/s of the network's 2 x 21.0 MMAC per frame)" % (nb, ms, nb * 42.0e6 / ms / 1e9))
with torch.no_grad():
    ref_all = torch.cat([oracle(i, min(i + 4096, nb))[2] for i in range(0, nb, 4096)])
print("all frames: scores max abs diff %.3e, argmax agreement %.4f" % (float((s_all - ref_all).abs().max()),
      float((s_all.argmax(1) == ref_all.argmax(1)).float().mean())))
# the cuDNN path it replaces (bf16, channels-last: the fastest setting of round 1)
from f2cnn_b200 import api
frames = torch.from_numpy(api.dense_frames(w, co, True, 50, dtype=np.float32)).cuda()
mb = cnn.seeded_model(seed=0)
for _ in range(2):
    cnn.predict(mb, frames, autocast_dtype=torch.bfloat16, channels_last=True)
torch.cuda.synchronize()
a.record()
for _ in range(3):
    cnn.predict(mb, frames, autocast_dtype=torch.bfloat16, channels_last=True)
z.record(); torch.cuda.synchronize()
print("cuDNN bf16 NHWC forward on materialised frames: %.3f ms" % (a.elapsed_time(z) / 3))
