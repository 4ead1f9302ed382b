"""Drop-in for the reference's scripts/processing/InputGenerator.py.

GenerateInputData keeps the reference's signature, files and row order: rows follow the
sorted file keys, CSV order within a file (:50, :67-82), values are the float64 envelope
samples at center + STEP*(j - RADIUS), cast to float32 at the very end (:83).  The gather
itself runs on the GPU (float64 -> float32 conversion on the device rounds to nearest even
exactly like numpy.astype).  If a listed .ENV1.npy file is missing but the utterance's .WAV
is present, the window rows are produced by the fused filterbank+envelope kernel straight
from the waveform (the reference would stop with FileNotFoundError).
"""
import csv
import os
import time
from configparser import ConfigParser

import numpy


def GetListOfEnvelopeFilesAndTimepoints(labelFilename):
    """{"<TEST|TRAIN>/<region>.<speaker>.<sentence>.ENV1.npy": [timepoints in CSV order]}.
    Reference :9-25."""
    output = dict()
    with open(labelFilename, 'r') as labelFile:
        for (testOrTrain, region, speaker, sentence, phoneme, timepoint, slope, pvalue, sign) in csv.reader(labelFile):
            key = os.path.join(testOrTrain, '.'.join((region, speaker, sentence, 'ENV1.npy')))
            output.setdefault(key, []).append(int(timepoint))
    return output


def GenerateInputData(labelFile=None, inputFile=None, LPF=False, CUTOFF=100):
    """Label CSV + envelope files -> (N, 2*RADIUS+1, NCHANNELS) float32 .npy.  Reference :28-93."""
    from ... import api
    TotalTime = time.time()

    if not os.path.isdir("trainingData"):
        print("LABEL GENERATION SHOULD BE DONE PRIOR TO INPUT...")
        exit(-1)
    csvFilename = labelFile or os.path.join("trainingData", "label_data.csv")
    filesAndTimepointsDict = GetListOfEnvelopeFilesAndTimepoints(csvFilename)

    print("\n###############################\nGenerating Input Data from files with '{}'.".format(csvFilename))
    if LPF:
        print("Using Low Pass Filtering with a cutoff at {}Hz".format(CUTOFF))
    else:
        print("Not using Low Pass Filtering")
    if not filesAndTimepointsDict:
        print("NO ENV1.npy FILES FOUND, PLEASE GENERATE ENVELOPES")
        exit(-1)
    files = sorted(filesAndTimepointsDict.keys())
    totalTimePoints = sum(len(data) for data in filesAndTimepointsDict.values())
    print(len(files), "files found along with their", totalTimePoints, "entry timepoints.")

    config = ConfigParser()
    config.read('configF2CNN.conf')
    RADIUS = config.getint('CNN', 'RADIUS')
    SAMPPERIOD = config.getint('CNN', 'SAMPLING_PERIOD')
    FRAMERATE = config.getint('FILTERBANK', 'FRAMERATE')
    NCHANNELS = config.getint('FILTERBANK', 'NCHANNELS')
    DOTSPERINPUT = RADIUS * 2 + 1
    STEP = int(FRAMERATE * SAMPPERIOD / 1000000)

    inputData = numpy.zeros((totalTimePoints, DOTSPERINPUT, NCHANNELS), dtype=numpy.float32)
    print("Output shape:", inputData.shape)
    currentEntry = 0
    fused = []  # (first row, wav path, timepoints): utterances without a saved envelope
    for currentFileIndex, file in enumerate(files):
        timepoints = filesAndTimepointsDict[file]
        path = os.path.join('resources', 'f2cnn', file)
        print("Reading:\t{}".format(path))
        if os.path.isfile(path):
            envelopes = numpy.load(path)
            if envelopes.shape[0] != NCHANNELS:
                raise ValueError("could not broadcast input array from shape ({},) into shape ({},)".format(
                    envelopes.shape[0], NCHANNELS))
            inputData[currentEntry:currentEntry + len(timepoints)] = api.gather_windows_from_matrix(
                envelopes, timepoints, RADIUS, STEP)
        else:
            wavPath = path[:-len('.ENV1.npy')] + '.WAV'
            if not os.path.isfile(wavPath):
                raise FileNotFoundError(2, 'No such file or directory', path)
            fused.append((currentEntry, wavPath, timepoints))
        currentEntry += len(timepoints)
        print("\t\t{:<50} done !  {}/{} Files".format(path, currentFileIndex + 1, len(files)))
    if fused:
        # one batched launch sequence for every utterance that has no .ENV1.npy: waveform ->
        # filterbank -> envelope -> windows without the 98 MB/utterance intermediates
        from concurrent.futures import ThreadPoolExecutor
        from .GammatoneFiltering import GetArrayFromWAV
        from ...gammatone import filters
        low = config.getint('FILTERBANK', 'LOW_FREQ')
        coefs = filters.make_erb_filters(FRAMERATE, filters.centre_freqs(FRAMERATE, NCHANNELS, low))
        with ThreadPoolExecutor(max_workers=8) as pool:
            wavs = [w for _, w in pool.map(GetArrayFromWAV, [f[1] for f in fused])]
        rows = api.features_to_windows(wavs, coefs, [f[2] for f in fused], LPF, CUTOFF, RADIUS, STEP)
        pos = 0
        for first, _, tps in fused:
            inputData[first:first + len(tps)] = rows[pos:pos + len(tps)]
            pos += len(tps)
    print('Generated Input Matrix of shape {}.'.format(inputData.shape))

    savePath = inputFile or (
        os.path.join('trainingData', 'input_data_LPF{}.npy'.format(CUTOFF) if LPF else 'input_data_NOLPF.npy'))
    print("Saving as {}...".format(savePath))
    os.makedirs(os.path.split(savePath)[0], exist_ok=True)
    numpy.save(savePath, inputData)
    numpy.save(os.path.join('trainingData', 'last_input_data.npy'), inputData)
    print('                Total time:', time.time() - TotalTime)
    print('')
