"""Where does the host time of one small API call go?  (3 s utterance -> windows on the label grid)"""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from f2cnn_b200 import api, engine, synth
from f2cnn_b200.gammatone import filters
co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
w = synth.white_noise_i16(48000, 0)
centers = synth.label_grid(48000)
for _ in range(5):
    api.features_to_windows([w], co, [centers], True, 50)
t = time.perf_counter()
for _ in range(50):
    api.features_to_windows([w], co, [centers], True, 50)
print("features_to_windows: %.3f ms per call" % ((time.perf_counter() - t) / 50 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    api.features_to_windows([w], co, [centers], True, 50)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
