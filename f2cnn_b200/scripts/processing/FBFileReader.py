"""VTR formant (.FB) files, behind the names of the reference's scripts/processing/FBFileReader.py.

File layout (reference :19-49): big-endian header int32 nFrame, int32 sampPeriod, int16 sampSize,
int16 fileType, then nFrame records of 8 big-endian float32 (F1..F4, B1..B4 in kHz).  The
reference ignores the stored sampPeriod and always returns 10000 (:26-27, one VTR file carries
100000), converts to Hz with round(value*1000, 2) and returns (matrix float64 (nFrame, 8), 10000),
or (None, 0) after printing when the file is missing.  Here the records are decoded by one
numpy.frombuffer instead of a struct.unpack per frame.
"""
import numpy

_HEADER = numpy.dtype([('nFrame', '>i4'), ('sampPeriod', '>i4'), ('sampSize', '>i2'), ('fileType', '>i2')])


def ExtractFBFile(fbFilename, verbose=False):
    try:
        with open(fbFilename, 'rb') as handle:
            raw = handle.read()
    except FileNotFoundError:
        print("No .FB formant data file.")
        return None, 0
    head = numpy.frombuffer(raw, dtype=_HEADER, count=1)[0]
    frames = int(head['nFrame'])
    period = 10000
    if verbose:
        print('N_SAMPLES=', frames)
        print('SAMP_PERIOD=', period)
        print('SAMP_SIZE=', int(head['sampSize']))
        print('NUM_COMPS=', int(head['sampSize']) / 4)
        print('FILE_TYPE=', int(head['fileType']))
    body = numpy.frombuffer(raw, dtype='>f4', count=frames * 8, offset=_HEADER.itemsize)
    # float32 kHz -> float64 Hz, 2 decimals: numpy.round agrees with the builtin round here because a
    # float32 times 1000 is at least 3e-8 away from any ...5 tie at the third decimal
    hz = numpy.round(body.astype(numpy.float64) * 1000, 2)
    return hz.reshape(frames, 8), period


def GetFormantFrequencies(fbFilename, formant):
    """Column `formant` (1..4) of the file in Hz and the sample period, or (None, None)."""
    matrix, period = ExtractFBFile(fbFilename)
    if matrix is None:
        return None, None
    return matrix[:, formant - 1], period


def formant_window_start(timepoint, radius, wavToFormant):
    """First frame of the window around `timepoint` (reference :77-78); also valid on arrays."""
    return numpy.asarray(numpy.asarray(timepoint) / wavToFormant - radius).astype(numpy.int64)


def GetFromantFrequenciesAround(array, timepoint, radius, wavToFormant):
    """The 2*radius+1 frames centred on timepoint/wavToFormant; print + exit(-1) outside the
    track, as the reference does (:80-88)."""
    start = int(timepoint / wavToFormant - radius)
    end = int(timepoint / wavToFormant + radius) + 1
    if start < 0 or end >= len(array):
        print("ERROR: WRONG RANGE IN GETFORMANTFREQUENCIESAROUND IN ARRAY OF LEN:\n", len(array),
              "\nAT TIME AND RADIUS", timepoint, radius, "START", start, "END", end)
        print("INF" if start < 0 else "SUP")
        print(len(array))
        exit(-1)
    return array[start:end]
