"""GPU implementation behind the names of the reference's scripts/processing/EnvelopeExtraction.py.

Kept from the reference (file:line there): paddedHilbert (:20-36) = analytic signal computed on
the signal zero-padded to 2^ceil(log2 n) and cut back; lowPassFilter (:39-48) = first-order
Butterworth at freq/8000 (the 8 kHz Nyquist is fixed whatever the configured frame rate),
zero initial state, along axis 0; ExtractEnvelopeFromMatrix (:51-67) = |paddedHilbert(row)| then
the low-pass iff LPF; ExtractEnvelope / SaveEnvelope / ExtractAndSaveEnvelope (:70-117) = the
.GFB.npy -> .ENV1.npy file step (METHOD 1); ExtractAllEnvelopes (:125-153) = that step for every
resources/f2cnn/*/*.GFB.npy.

Hilbert transform, magnitude and low-pass run on the B200: a hand-written FFT per row and a
chunked warp-shuffle scan for the first-order recurrence (libf2cnn_b200.so, f2_rows_op).
"""
from __future__ import division

import glob
import time
from concurrent.futures import ThreadPoolExecutor
from multiprocessing import Value
from os.path import join, split, splitext

import numpy

from .GammatoneFiltering import npy_dtype

counter = None


def paddedHilbert(signal):
    from ... import api
    samples = numpy.asarray(signal)
    quadrature = api.hilbert_imag_rows(samples.reshape(1, -1))[0]
    return samples.astype(numpy.float64) + 1j * quadrature


def lowPassFilter(signal, freq):
    from ... import api
    samples = numpy.asarray(signal)
    if samples.ndim == 1:
        return api.lowpass_rows(samples.reshape(1, -1), freq)[0]
    # lfilter(..., axis=0) on an N-d array: filter every column
    columns_last = numpy.moveaxis(samples, 0, -1)
    rows = numpy.ascontiguousarray(columns_last).reshape(-1, samples.shape[0])
    filtered = api.lowpass_rows(rows, freq).reshape(columns_last.shape)
    return numpy.ascontiguousarray(numpy.moveaxis(filtered, -1, 0))


def ExtractEnvelopeFromMatrix(matrix, LPF=False, CUTOFF=100):
    from ... import api
    return api.extract_envelope_from_matrix(matrix, LPF, CUTOFF)


def ExtractEnvelope(gfbFileName, LPF=False, CUTOFF=100):
    print("File:\t{}".format(gfbFileName))
    return ExtractEnvelopeFromMatrix(numpy.load(gfbFileName), LPF, CUTOFF)


def _envelope_name(gfbFileName, method=1):
    return splitext(splitext(gfbFileName)[0])[0] + ".ENV" + str(method)


def SaveEnvelope(matrix, gfbFileName, nbf):
    target = _envelope_name(gfbFileName)
    numpy.save(target, matrix)
    if counter is not None:
        with counter.get_lock():
            counter.value += 1
            print("\t{:<50} done ! {}/{} Files.".format(target, counter.value, nbf))


def ExtractAndSaveEnvelope(gfbFileName, nbf, LPF=False, CUTOFF=100):
    SaveEnvelope(ExtractEnvelope(gfbFileName, LPF, CUTOFF), gfbFileName, nbf)


def InitProcesses(cn):
    global counter
    counter = cn


def ExtractAllEnvelopes(LPF=False, CUTOFF=100):
    started = time.time()
    found = glob.glob(join("resources", "f2cnn", "*", "*.GFB.npy"))
    # as in the reference (:132 before :138) the banner indexes the list first: IndexError if empty
    print("\n###############################\nExtracting Envelopes from files in '{}'.".format(split(found[0])[0]))
    print("Using Low Pass Filtering with a cutoff at {}Hz".format(CUTOFF) if LPF else "Not using Low Pass Filtering")
    if not found:
        print("ERROR: NO .GFB.npy FILES FOUND, PLEASE GENERATE FILTERED OUTPUTS")
        exit(-1)
    print(len(found), ".GFB.npy files found")

    InitProcesses(Value('i', 0))
    # one thread drives the GPU; matrix loads run a few files ahead, the envelopes of a batch land in one of
    # two pinned host slots and the writer threads save straight from there
    from ... import api
    stream = api.MatrixStream(None, "envelope", LPF=LPF, CUTOFF=CUTOFF, dtype=npy_dtype())
    lookahead = 8
    with ThreadPoolExecutor(max_workers=4) as loaders, ThreadPoolExecutor(max_workers=8) as writers:
        loading, submitted = {}, 0

        def matrix_of(i):
            nonlocal submitted
            while submitted < min(len(found), i + lookahead):
                loading[submitted] = loaders.submit(numpy.load, found[submitted])
                submitted += 1
            return loading[i].result()

        pending = []
        i = 0
        while i < len(found):
            group, used = [], 0
            while i < len(found):
                matrix = matrix_of(i)
                need = matrix.size * stream.per_value
                if group and used + need > stream.slot_bytes:
                    break
                print("File:\t{}".format(found[i]))
                group.append((found[i], matrix))
                used += need
                del loading[i]
                i += 1
            envelopes = stream.process([m for _, m in group])
            jobs = [writers.submit(SaveEnvelope, e, name, len(found)) for (name, _), e in zip(group, envelopes)]
            stream.retire(jobs)
            pending.extend(jobs)
        for job in pending:
            job.result()
    print("Extracted Envelopes from all files.")
    print('              Total time:', time.time() - started)
    print('')
