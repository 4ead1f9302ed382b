"""GPU implementation behind the names of the reference's scripts/processing/LabelDataGenerator.py.

Contract kept from the reference (file:line there):
  * ExtractLabel(wavFile, config) (:22-77): rows [TEST|TRAIN, region, speaker, sentence, phoneme,
    timepoint, round(slope, 5), round(p, 5), 1 if slope > 0 else 0] for the timepoints
    START + k*STEP, k < nb = int(nf/(FRAMERATE*SAMPLING_PERIOD*1e-6) - DOTSPERINPUT - 1), whose
    phoneme is not silent ('pau', 'epi', 'h#') and whose slope p-value is below RISK; the scan
    stops at the first empty phoneme string (:58-59); None when nothing is kept or the .FB file
    is missing.
  * GenerateLabelData() (:80-122): every resources/f2cnn/*/*.WAV in sorted order ->
    trainingData/label_data.csv (lineterminator '\\n').
The slope / p-value of every timepoint of every file come from ONE kernel launch
(libf2cnn_b200.so f2_label_fit: closed-form least squares + incomplete beta in float64) instead of
one numpy.linalg.lstsq and one scipy.stats.pearsonr per timepoint, and the sample count is read
from the WAV header instead of decoding the file.
"""
import csv
import glob
import os
import time
from configparser import ConfigParser

import numpy

from .FBFileReader import GetFormantFrequencies, formant_window_start
from .PHNFileReader import SILENTS, ExtractPhonemes, phoneme_at


def wav_shape(wavFile):
    """(framerate, number of samples) from the header of a RIFF or NIST SPHERE file."""
    from ...ingest import wav_layout
    layout = wav_layout(wavFile)
    return layout.rate, layout.samples


class _Geometry:
    def __init__(self, config):
        self.radius = config.getint('CNN', 'RADIUS')
        self.risk = config.getfloat('CNN', 'RISK')
        self.formant = config.getint('CNN', 'FORMANT')
        self.period = config.getint('CNN', 'SAMPLING_PERIOD')
        self.dots = 2 * self.radius + 1


def _candidates(wavFile, geo):
    """Everything ExtractLabel knows before the regression: identity columns, the formant track,
    and the (timepoint, phoneme) pairs that survive the phoneme filter.  None without .FB file."""
    stem = os.path.splitext(wavFile)[0]
    track, _ = GetFormantFrequencies(stem + '.FB', geo.formant)
    if track is None:
        return None
    phonemes = ExtractPhonemes(stem + '.PHN')
    rate, nf = wav_shape(wavFile)
    to_formant = rate * geo.period * (1.0 / 1000000)
    nb = int(nf / to_formant - geo.dots - 1)
    step = int(to_formant)
    points = step * geo.radius + step * numpy.arange(max(nb, 0), dtype=numpy.int64)
    names = phoneme_at(phonemes, points)
    kept = []
    for t, name in zip(points.tolist(), names):
        if name in SILENTS:
            continue
        if not name:
            break
        kept.append((t, name))
    centers = numpy.asarray([t for t, _ in kept], dtype=numpy.int64)
    first = formant_window_start(centers, geo.radius, to_formant)
    last = (centers / to_formant + geo.radius).astype(numpy.int64) + 1
    bad = numpy.nonzero((first < 0) | (last >= len(track)))[0]
    if bad.size:  # FBFileReader.py:80-88
        i = int(bad[0])
        print("ERROR: WRONG RANGE IN GETFORMANTFREQUENCIESAROUND IN ARRAY OF LEN:\n", len(track),
              "\nAT TIME AND RADIUS", int(centers[i]), geo.radius, "START", int(first[i]), "END", int(last[i]))
        print("INF" if first[i] < 0 else "SUP")
        print(len(track))
        exit(-1)
    region, speaker, sentence = os.path.split(stem)[1].split(".")
    split = os.path.split(os.path.split(stem)[0])[1]
    return dict(ident=[split, region, speaker, sentence], track=numpy.ascontiguousarray(track, dtype=numpy.float64),
                kept=kept, centers=centers, first=first, step=step)


def _rows(cand, fit, risk):
    rows = []
    for (t, name), (slope, _, _, p) in zip(cand['kept'], fit.tolist()):
        if p < risk:  # NaN (flat track) compares false, as in the reference
            rows.append(cand['ident'] + [name, t, round(slope, 5), round(p, 5), 1 if slope > 0 else 0])
    return rows


def _fit_all(cands, geo):
    """One launch for all files; returns the per-file slices of the (N, 4) result."""
    from ... import api
    steps = {c['step'] for c in cands}
    out = [None] * len(cands)
    for step in steps:  # one launch per distinct STEP (one, unless the files differ in frame rate)
        group = [i for i, c in enumerate(cands) if c['step'] == step]
        fit = api.label_fit([cands[i]['track'] for i in group], [cands[i]['first'] for i in group],
                            [cands[i]['centers'] for i in group], geo.radius, step)
        at = 0
        for i in group:
            n = len(cands[i]['kept'])
            out[i] = fit[at:at + n]
            at += n
    return out


def ExtractLabel(wavFile, config):
    geo = _Geometry(config)
    cand = _candidates(wavFile, geo)
    if cand is None:
        return None
    rows = _rows(cand, _fit_all([cand], geo)[0], geo.risk) if cand['kept'] else []
    return rows if rows else None


def GenerateLabelData():
    started = time.time()
    config = ConfigParser()
    config.read('configF2CNN.conf')
    found = glob.glob(os.path.join("resources", "f2cnn", "*", "*.WAV"))
    # the reference formats found[0] into the banner before looking at the list (:89-90)
    print("\n###############################\nGenerating Label Data from files in '{}' into 2 classes.".format(
        os.path.split(os.path.split(found[0])[0])[0]))
    found = sorted(found)
    if not found:
        print("NO FILES FOUND")
        exit(-1)
    print(len(found), "files found")

    geo = _Geometry(config)
    cands = []
    for i, name in enumerate(found):
        print("Reading:\t{:<50}\t{}/{}".format(name, i, len(found)))
        cands.append(_candidates(name, geo))
        print("\t\t{:<50}\tdone !".format(name))
    live = [c for c in cands if c is not None and c['kept']]
    lines = []
    for cand, fit in zip(live, _fit_all(live, geo) if live else []):
        lines.extend(_rows(cand, fit, geo.risk))

    target = os.path.join("trainingData", "label_data.csv")
    print("Saving {} lines in '{}'.".format(len(lines), target))
    os.makedirs(os.path.split(target)[0], exist_ok=True)
    with open(target, "w") as handle:
        writer = csv.writer(handle, lineterminator='\n')
        writer.writerows(lines)
    print("Generated Label Data CSV of", len(lines), "lines.")
    print('                Total time:', time.time() - started)
    print('')
