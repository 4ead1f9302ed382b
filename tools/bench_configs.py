"""Device-side timings of the BASELINE.json configs that are not the bench.py headline:
config 1 (single 3 s utterance, latency), config 4 (600 s stream, 256 channels, cut-off sweep),
config 5 (evalnoise front end: dense framing + normalisation + CNN forward)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import api, cnn, engine, synth
from f2cnn_b200.gammatone import filters

def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = engine.DeviceEvent(), engine.DeviceEvent()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_ms(b) / reps

out = {}
co128 = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
co256 = filters.make_erb_filters(16000, filters.centre_freqs(16000, 256, 100))
# ---- config 1 ----
w = synth.white_noise_i16(48000, 0)
plan = engine.plan_for(co128)
wd = torch.from_numpy(w).cuda()
for target, tag in ((1, "one_item"), (0, "default_chunks")):
    b = plan.batch([48000], target_items=target)
    dec = torch.empty((b.total_frames, 128), device="cuda")
    ms = timed(lambda: b.run(wd, lpf=True, cutoff=50, out={"dec": dec}), reps=20)
    out["config1_dec_%s" % tag] = {"items": b.num_items, "device_ms": ms, "channel_samples_per_s": 128 * 48000 / ms * 1e3}
t = time.perf_counter(); centers = synth.label_grid(48000)
for _ in range(5):
    api.features_to_windows([w], co128, [centers], True, 50)
out["config1_api_windows_host_ms"] = (time.perf_counter() - t) / 5 * 1e3
t = time.perf_counter()
for _ in range(3):
    api.filterbank_envelope(w, co128, True, 50, with_gfb=True)
out["config1_api_gfb_env_float64_host_ms"] = (time.perf_counter() - t) / 3 * 1e3
# ---- config 4 ----
n = 9_600_000
w4 = torch.from_numpy(synth.white_noise_i16(n, seed=2)).cuda()
plan4 = engine.plan_for(co256)
b4 = plan4.batch([n])
dec4 = torch.empty((b4.total_frames, 256), device="cuda")
for cut in (20, 50, 100):
    ms = timed(lambda: b4.run(w4, lpf=True, cutoff=cut, out={"dec": dec4}), reps=3, warm=1)
    out["config4_cutoff%d" % cut] = {"items": b4.num_items, "device_ms": ms, "channel_samples_per_s": 256.0 * n / ms * 1e3}
# ---- config 5 ----
base = synth.speech_like_i16(48000, seed=31)
wave = base + np.random.default_rng(3).normal(scale=300.0, size=48000)
model = cnn.seeded_model(0)
def eval_path(cnn_dtype=None, with_cnn=True, channels_last=False):
    b5 = plan.batch([48000])
    res = b5.run(torch.from_numpy(wave).cuda(), lpf=True, cutoff=50, env_t=True)
    frames, flag = engine.dense_frames(res["env_t"], 11, 160, 0, 48000 - 1760, normalize=True, out_dtype=torch.float32)
    return cnn.predict(model, frames, autocast_dtype=cnn_dtype, channels_last=channels_last) if with_cnn else frames
out["config5_eval_frontend_only"] = {"frames": 48000 - 1760, "device_ms": timed(lambda: eval_path(with_cnn=False), reps=5)}
out["config5_eval_frontend_plus_cnn_fp32"] = {"device_ms": timed(eval_path, reps=5),
                                              "note": "reference: 25 s Python framing + 0.9 s normalise per utterance before Keras"}
out["config5_eval_frontend_plus_cnn_bf16"] = {"device_ms": timed(lambda: eval_path(torch.bfloat16), reps=5)}
out["config5_eval_frontend_plus_cnn_bf16_nhwc"] = {"device_ms": timed(lambda: eval_path(torch.bfloat16, channels_last=True), reps=5)}
# ---- full-rate output modes (HBM-bound per SURVEY.md 8d): 256 corpus utterances ----
lengths = synth.corpus_lengths(256, seed=1)
flat, _ = synth.corpus_waves_i16(lengths, seed=1)
fd = torch.from_numpy(flat).cuda()
bq = plan.batch(lengths)
cs = 128.0 * float(lengths.sum())
for tag, kw, bytes_per_cs in (("env_t_f32_time_major", dict(env_t=True), 4.04),
                              ("gfb_env_f32_reference_layout", dict(gfb=torch.float32, env=torch.float32), 8.04),
                              ("gfb_env_f64_reference_layout", dict(gfb=torch.float64, env=torch.float64), 16.04)):
    res = bq.run(fd, lpf=True, cutoff=50, **kw)
    ms = timed(lambda: bq.run(fd, lpf=True, cutoff=50, out=res, **kw), reps=3, warm=1)
    out["fullrate_" + tag] = {"device_ms": ms, "channel_samples_per_s": cs / ms * 1e3,
                              "algorithmic_GBps": bytes_per_cs * cs / ms / 1e6}
    del res
# ---- label generation: the corpus label grid, one launch ----
lens = synth.corpus_lengths(4620, seed=1)
rng = np.random.default_rng(5)
tracks = [1500.0 + np.cumsum(rng.normal(0, 8.0, int(n) // 160 + 3)) for n in lens]
centers = [synth.label_grid(int(n)) for n in lens]
firsts = [c // 160 - 5 for c in centers]
offs = np.concatenate([[0], np.cumsum([len(t) for t in tracks])]).astype(np.int64)
first_d = torch.from_numpy(np.concatenate([f + o for f, o in zip(firsts, offs)])).cuda()
center_d = torch.from_numpy(np.concatenate(centers).astype(np.int32)).cuda()
flat_d = torch.from_numpy(np.concatenate(tracks)).cuda()
ms = timed(lambda: engine.label_fit(flat_d, first_d, center_d, 11, 160), reps=10)
t = time.perf_counter()
fit = api.label_fit(tracks, firsts, centers)
host_ms = (time.perf_counter() - t) * 1e3
out["label_fit_corpus_grid"] = {"timepoints": int(first_d.shape[0]), "device_ms": ms, "api_host_ms_incl_copies": host_ms,
                                "note": "reference: numpy.linalg.lstsq + scipy.stats.pearsonr per timepoint, "
                                        "~0.16 s per file (SURVEY.md 8f)"}
print(json.dumps(out, indent=1))
