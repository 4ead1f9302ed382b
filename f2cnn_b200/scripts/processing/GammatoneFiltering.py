"""Drop-in for the reference's scripts/processing/GammatoneFiltering.py.

Same public names and signatures (reference file:line in each docstring); the filterbank
itself runs on the B200 through libf2cnn_b200.so.  The reference fans files out over a
multiprocessing.Pool (:121-125); a CUDA context must not be forked, so the batch driver here
keeps one host thread that feeds the GPU and hands the finished matrices to a small pool of
writer threads (numpy.save releases the GIL while writing).
"""
import glob
import os
import struct
import time
from concurrent.futures import ThreadPoolExecutor
from configparser import ConfigParser
from multiprocessing import Value

import numpy

from ...gammatone import filters

counter = None
FILTERBANK_COEFFICIENTS = None


def _read_nist_sphere(filename):
    """NIST SPHERE reader for TIMIT's uncompressed 16-bit PCM files (the reference goes through
    the third-party `sphfile` package and copies sample by sample, :33-38)."""
    with open(filename, 'rb') as f:
        magic = f.readline()
        if not magic.startswith(b'NIST_1A'):
            raise ValueError("{} is neither RIFF nor NIST SPHERE".format(filename))
        header_size = int(f.readline().strip())
        f.seek(0)
        header = f.read(header_size).decode('ascii', 'replace')
        fields = {}
        for line in header.splitlines()[2:]:
            parts = line.split(None, 2)
            if not parts or parts[0] == 'end_head':
                break
            if len(parts) == 3:
                fields[parts[0]] = parts[2]
        coding = fields.get('sample_coding', 'pcm')
        if coding != 'pcm':
            raise ValueError("unsupported SPHERE sample_coding '{}' in {}".format(coding, filename))
        if int(fields.get('sample_n_bytes', 2)) != 2 or int(fields.get('channel_count', 1)) != 1:
            raise ValueError("only mono 16-bit SPHERE files are supported: {}".format(filename))
        count = int(fields['sample_count'])
        order = '<' if fields.get('sample_byte_format', '01') == '01' else '>'
        data = numpy.frombuffer(f.read(2 * count), dtype=order + 'i2').astype(numpy.int16)
    return int(fields['sample_rate']), data


def GetArrayFromWAV(filename):
    """(framerate, int16 array) from a RIFF WAVE or a NIST SPHERE file.  Reference :28-39."""
    with open(filename, 'rb') as wavFile:
        header = wavFile.read(4)
    if header == b'RIFF':
        from scipy.io import wavfile as WavFileTool
        framerate, wavArray = WavFileTool.read(filename)
    else:
        framerate, wavArray = _read_nist_sphere(filename)
    return framerate, wavArray


def GetFilteredOutputFromArray(array, FILTERBANK_COEFFICIENTS):
    """(C, n) float64 filterbank output of a 1-D array.  Reference :42-47."""
    return filters.erb_filterbank(array, FILTERBANK_COEFFICIENTS)


def GetFilteredOutputFromFile(filename, FILTERBANK_COEFFICIENTS):
    """(matrix, framerate) for a WAV file.  Reference :50-58."""
    framerate, wavArray = GetArrayFromWAV(filename)
    return GetFilteredOutputFromArray(wavArray, FILTERBANK_COEFFICIENTS), framerate


def saveGFBMatrix(filename, matrix):
    """numpy.save(filename, matrix) -> filename + '.npy'.  Reference :61-62."""
    numpy.save(filename, matrix)


def loadGFBMatrix(filename):
    """Reference :65-66."""
    return numpy.load(filename + '.npy')


def _count_done(label, n):
    global counter
    if counter is None:
        return
    with counter.get_lock():
        counter.value += 1
        print("\t\t{:<50} done ! {}/{} Files.".format(label, counter.value, n))


def GammatoneFiltering(wavFile, n):
    """Filter one WAV file and save <base>.GFB.npy.  Reference :69-83."""
    gfbFilename = os.path.splitext(wavFile)[0] + '.GFB'
    print("Filtering:\t{}".format(wavFile))
    outputMatrix, _ = GetFilteredOutputFromFile(wavFile, FILTERBANK_COEFFICIENTS)
    print("Saving:\t\t{}.npy".format(gfbFilename))
    saveGFBMatrix(gfbFilename, outputMatrix)
    _count_done(wavFile, n)


def InitProcesses(FBCOEFS, cn):
    """Reference :86-90 (kept for callers; no worker processes are forked here)."""
    global FILTERBANK_COEFFICIENTS
    global counter
    counter = cn
    FILTERBANK_COEFFICIENTS = FBCOEFS


def FilterAllOrganisedFiles():
    """Filter every resources/f2cnn/**/*.WAV into .GFB.npy.  Reference :93-129."""
    TotalTime = time.time()
    wavFiles = glob.glob(os.path.join("resources", "f2cnn", "**", "*.WAV"))

    # the reference indexes wavFiles[0] before its emptiness check (:100-103): same IndexError
    print("\n###############################\nApplying FilterBank to files in '{}'.".format(
        os.path.split(wavFiles[0])[0]))
    if not wavFiles:
        print("NO WAV FILES FOUND, PLEASE ORGANIZE FILES")
        exit(-1)
    print(len(wavFiles), "files found")

    config = ConfigParser()
    config.read('configF2CNN.conf')
    framerate = config.getint('FILTERBANK', 'FRAMERATE')
    nchannels = config.getint('FILTERBANK', 'NCHANNELS')
    lowcutoff = config.getint('FILTERBANK', 'LOW_FREQ')
    CENTER_FREQUENCIES = filters.centre_freqs(framerate, nchannels, lowcutoff)
    COEFS = filters.make_erb_filters(framerate, CENTER_FREQUENCIES)

    InitProcesses(COEFS, Value('i', 0))
    nfiles = len(wavFiles)

    def save(wavFile, matrix):
        gfbFilename = os.path.splitext(wavFile)[0] + '.GFB'
        print("Saving:\t\t{}.npy".format(gfbFilename))
        saveGFBMatrix(gfbFilename, matrix)
        _count_done(wavFile, nfiles)

    with ThreadPoolExecutor(max_workers=4) as writers:
        pending = []
        for wavFile in wavFiles:
            print("Filtering:\t{}".format(wavFile))
            matrix, _ = GetFilteredOutputFromFile(wavFile, COEFS)
            pending.append(writers.submit(save, wavFile, matrix))
            while len(pending) > 8:  # bound the host memory held by queued 49 MB matrices
                pending.pop(0).result()
        for p in pending:
            p.result()

    print("Filtered and Saved all files.")
    print('                Total time:', time.time() - TotalTime)
    print('')
