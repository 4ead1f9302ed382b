"""GPU implementation behind the names of the reference's scripts/processing/GammatoneFiltering.py.

Kept from the reference (file:line there): GetArrayFromWAV (:28-39) returns (framerate, int16
samples) for RIFF and NIST SPHERE files; GetFilteredOutputFromArray / ...FromFile (:42-58) return
the (C, n) float64 filterbank matrix (and the frame rate); saveGFBMatrix / loadGFBMatrix
(:61-66) are numpy.save / numpy.load(name + '.npy'); GammatoneFiltering (:69-83) writes
<base>.GFB.npy for one file; InitProcesses (:86-90) sets the module globals; and
FilterAllOrganisedFiles (:93-129) filters every resources/f2cnn/**/*.WAV with the bank described
by configF2CNN.conf.

The filterbank runs on the B200 (libf2cnn_b200.so).  The reference fans the files out over a
multiprocessing.Pool (:121-125); a CUDA context must not be forked, so the driver keeps one
thread feeding the GPU and hands finished matrices to writer threads (numpy.save releases the
GIL while it writes 49 MB per utterance).  Opt-in: F2CNN_B200_NPY_FLOAT32=1 in the environment
makes the two file drivers (this one and ExtractAllEnvelopes) store float32 matrices -- half the
PCIe and disk traffic; every reader of this package takes either dtype.  The array functions keep
the reference's float64.
"""
import glob
import os
import time
from concurrent.futures import ThreadPoolExecutor
from configparser import ConfigParser
from multiprocessing import Value

import numpy

from ...gammatone import filters

counter = None
FILTERBANK_COEFFICIENTS = None


# ---- reading ----------------------------------------------------------------------------------
def _sphere_fields(raw_header):
    """key -> value of a NIST SPHERE header ('name -type value' lines up to end_head)."""
    fields = {}
    for line in raw_header.splitlines()[2:]:
        tokens = line.split(None, 2)
        if tokens and tokens[0] == 'end_head':
            break
        if len(tokens) == 3:
            fields[tokens[0]] = tokens[2]
    return fields


def _read_nist_sphere(filename):
    """TIMIT's uncompressed mono 16-bit SPHERE files (the reference uses the third-party `sphfile`
    package and copies sample by sample, :33-38)."""
    with open(filename, 'rb') as handle:
        if not handle.readline().startswith(b'NIST_1A'):
            raise ValueError("{} is neither RIFF nor NIST SPHERE".format(filename))
        header_bytes = int(handle.readline().strip())
        handle.seek(0)
        fields = _sphere_fields(handle.read(header_bytes).decode('ascii', 'replace'))
        if fields.get('sample_coding', 'pcm') != 'pcm':
            raise ValueError("unsupported SPHERE sample_coding '{}' in {}".format(fields['sample_coding'], filename))
        if int(fields.get('sample_n_bytes', 2)) != 2 or int(fields.get('channel_count', 1)) != 1:
            raise ValueError("only mono 16-bit SPHERE files are supported: {}".format(filename))
        little = fields.get('sample_byte_format', '01') == '01'
        payload = handle.read(2 * int(fields['sample_count']))
    samples = numpy.frombuffer(payload, dtype='<i2' if little else '>i2').astype(numpy.int16)
    return int(fields['sample_rate']), samples


def GetArrayFromWAV(filename):
    with open(filename, 'rb') as handle:
        magic = handle.read(4)
    if magic != b'RIFF':
        return _read_nist_sphere(filename)
    from scipy.io import wavfile
    rate, samples = wavfile.read(filename)
    return rate, samples


# ---- array / file functions ---------------------------------------------------------------------
def GetFilteredOutputFromArray(array, FILTERBANK_COEFFICIENTS):
    return filters.erb_filterbank(array, FILTERBANK_COEFFICIENTS)


def GetFilteredOutputFromFile(filename, FILTERBANK_COEFFICIENTS):
    rate, samples = GetArrayFromWAV(filename)
    return GetFilteredOutputFromArray(samples, FILTERBANK_COEFFICIENTS), rate


def saveGFBMatrix(filename, matrix):
    numpy.save(filename, matrix)


def loadGFBMatrix(filename):
    return numpy.load(filename + '.npy')


def npy_dtype():
    """float64 like the reference, or float32 when F2CNN_B200_NPY_FLOAT32 is set to a non-zero value."""
    return numpy.float32 if os.environ.get('F2CNN_B200_NPY_FLOAT32', '0') not in ('', '0') else numpy.float64


def _gfb_name(wavFile):
    return os.path.splitext(wavFile)[0] + '.GFB'


def _store(wavFile, matrix, total):
    name = _gfb_name(wavFile)
    print("Saving:\t\t{}.npy".format(name))
    saveGFBMatrix(name, matrix)
    if counter is not None:
        with counter.get_lock():
            counter.value += 1
            print("\t\t{:<50} done ! {}/{} Files.".format(wavFile, counter.value, total))


def GammatoneFiltering(wavFile, n):
    print("Filtering:\t{}".format(wavFile))
    matrix, _ = GetFilteredOutputFromFile(wavFile, FILTERBANK_COEFFICIENTS)
    _store(wavFile, matrix, n)


def InitProcesses(FBCOEFS, cn):
    global FILTERBANK_COEFFICIENTS, counter
    FILTERBANK_COEFFICIENTS, counter = FBCOEFS, cn


# ---- batch driver ---------------------------------------------------------------------------------
def _configured_bank():
    cfg = ConfigParser()
    cfg.read('configF2CNN.conf')
    rate = cfg.getint('FILTERBANK', 'FRAMERATE')
    centre = filters.centre_freqs(rate, cfg.getint('FILTERBANK', 'NCHANNELS'), cfg.getint('FILTERBANK', 'LOW_FREQ'))
    return filters.make_erb_filters(rate, centre)


def FilterAllOrganisedFiles():
    started = time.time()
    found = glob.glob(os.path.join("resources", "f2cnn", "**", "*.WAV"))
    # the reference prints the directory of found[0] BEFORE testing for emptiness (:100-103), so
    # an empty tree ends in IndexError there as well
    print("\n###############################\nApplying FilterBank to files in '{}'.".format(os.path.split(found[0])[0]))
    if not found:
        print("NO WAV FILES FOUND, PLEASE ORGANIZE FILES")
        exit(-1)
    print(len(found), "files found")

    InitProcesses(_configured_bank(), Value('i', 0))
    # Many utterances per launch sequence: the matrices of a batch land in one of two pinned host slots and
    # the writer threads save straight from there while the next batch is read, filtered and downloaded.
    from ... import api
    stream = api.MatrixStream(FILTERBANK_COEFFICIENTS, "filterbank", dtype=npy_dtype())
    channels = FILTERBANK_COEFFICIENTS.shape[0]
    with ThreadPoolExecutor(max_workers=8) as readers, ThreadPoolExecutor(max_workers=8) as writers:
        pending = []
        chunk = 64   # files decoded ahead, regrouped into slot-sized batches below
        for lo in range(0, len(found), chunk):
            names = found[lo:lo + chunk]
            waves = [samples for _, samples in readers.map(GetArrayFromWAV, names)]
            for group in stream.batches([channels * len(w) for w in waves]):
                for i in group:
                    print("Filtering:\t{}".format(names[i]))
                matrices = stream.process([waves[i] for i in group])
                jobs = [writers.submit(_store, names[i], m, len(found)) for i, m in zip(group, matrices)]
                stream.retire(jobs)
                pending.extend(jobs)
        for job in pending:
            job.result()
    print("Filtered and Saved all files.")
    print('                Total time:', time.time() - started)
    print('')
