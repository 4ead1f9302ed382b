"""Drop-in for the reference's scripts/processing/EnvelopeExtraction.py.

Same public names and signatures; Hilbert transform, magnitude and low-pass run on the B200
(hand-written FFT + chunked scan in libf2cnn_b200.so).  When the matrix handed to
ExtractEnvelopeFromMatrix is the untouched result of this package's erb_filterbank, the fused
kernel computes the envelope straight from the waveform instead of re-uploading the 49 MB
matrix (same values: both paths are held to the same oracle).
"""
from __future__ import division

import glob
import time
from concurrent.futures import ThreadPoolExecutor
from multiprocessing import Value
from os.path import splitext, join, split

import numpy

counter = None


def paddedHilbert(signal):
    """Analytic signal of `signal`, computed on the zero-padded power-of-two length and cut
    back: complex128 (n,).  Reference :20-36."""
    from ... import api
    sig = numpy.asarray(signal)
    imag = api.hilbert_imag_rows(sig.reshape(1, -1))[0]
    return sig.astype(numpy.float64) + 1j * imag


def lowPassFilter(signal, freq):
    """butter(1, freq/8000, 'low') applied along axis 0, zero initial state.  The Nyquist
    frequency is fixed at 8000 Hz whatever the configured frame rate.  Reference :39-48."""
    from ... import api
    sig = numpy.asarray(signal)
    if sig.ndim == 1:
        return api.lowpass_rows(sig.reshape(1, -1), freq)[0]
    moved = numpy.moveaxis(sig, 0, -1)
    flat = numpy.ascontiguousarray(moved).reshape(-1, sig.shape[0])
    out = api.lowpass_rows(flat, freq).reshape(moved.shape)
    return numpy.ascontiguousarray(numpy.moveaxis(out, -1, 0))


def ExtractEnvelopeFromMatrix(matrix, LPF=False, CUTOFF=100):
    """abs(paddedHilbert(row)) for every row, then lowPassFilter(., CUTOFF) iff LPF:
    float64, same shape.  Reference :51-67."""
    from ... import api
    return api.extract_envelope_from_matrix(matrix, LPF, CUTOFF)


def ExtractEnvelope(gfbFileName, LPF=False, CUTOFF=100):
    """Load a .GFB.npy matrix and extract its envelopes.  Reference :70-83."""
    print("File:\t{}".format(gfbFileName))
    matrix = numpy.load(gfbFileName)
    return ExtractEnvelopeFromMatrix(matrix, LPF, CUTOFF)


def SaveEnvelope(matrix, gfbFileName, nbf):
    """Save as <base>.ENV1.npy (METHOD = 1).  Reference :86-99."""
    METHOD = 1
    envelopeFilename = splitext(splitext(gfbFileName)[0])[0] + ".ENV" + str(METHOD)
    numpy.save(envelopeFilename, matrix)
    global counter
    if counter is not None:
        with counter.get_lock():
            counter.value += 1
            print("\t{:<50} done ! {}/{} Files.".format(envelopeFilename, counter.value, nbf))


def ExtractAndSaveEnvelope(gfbFileName, nbf, LPF=False, CUTOFF=100):
    """Reference :101-117."""
    SaveEnvelope(ExtractEnvelope(gfbFileName, LPF, CUTOFF), gfbFileName, nbf)


def InitProcesses(cn):
    """Reference :120-122 (kept for callers; no worker processes are forked here)."""
    global counter
    counter = cn


def ExtractAllEnvelopes(LPF=False, CUTOFF=100):
    """Every resources/f2cnn/*/*.GFB.npy -> .ENV1.npy.  Reference :125-153."""
    TotalTime = time.time()
    gfbFiles = glob.glob(join("resources", "f2cnn", "*", "*.GFB.npy"))
    # like the reference (:132 before :138) an empty tree fails on gfbFiles[0]
    print("\n###############################\nExtracting Envelopes from files in '{}'.".format(split(gfbFiles[0])[0]))
    if LPF:
        print("Using Low Pass Filtering with a cutoff at {}Hz".format(CUTOFF))
    else:
        print("Not using Low Pass Filtering")
    if not gfbFiles:
        print("ERROR: NO .GFB.npy FILES FOUND, PLEASE GENERATE FILTERED OUTPUTS")
        exit(-1)
    print(len(gfbFiles), ".GFB.npy files found")

    InitProcesses(Value('i', 0))
    nbf = len(gfbFiles)
    # one thread drives the GPU; loads run ahead and saves trail behind on helper threads
    with ThreadPoolExecutor(max_workers=2) as loaders, ThreadPoolExecutor(max_workers=4) as writers:
        ahead = 4
        loads = [loaders.submit(numpy.load, f) for f in gfbFiles[:ahead]]
        pending = []
        for i, gfbFileName in enumerate(gfbFiles):
            print("File:\t{}".format(gfbFileName))
            matrix = loads[i].result()
            loads[i] = None
            if i + ahead < nbf:
                loads.append(loaders.submit(numpy.load, gfbFiles[i + ahead]))
            envelopes = ExtractEnvelopeFromMatrix(matrix, LPF, CUTOFF)
            pending.append(writers.submit(SaveEnvelope, envelopes, gfbFileName, nbf))
            while len(pending) > 8:
                pending.pop(0).result()
        for p in pending:
            p.result()

    print("Extracted Envelopes from all files.")
    print('              Total time:', time.time() - TotalTime)
    print('')
