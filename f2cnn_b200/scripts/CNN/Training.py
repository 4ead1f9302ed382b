"""Drop-in for the part of the reference's scripts/CNN/Training.py that belongs to the feature
path: normalizeInput (reference :13-28).  Training the Keras model itself is out of scope and
stays in the reference tree.  The batched, on-device version of this function is fused into
the dense-framing kernel (f2_dense_frames with normalize=1)."""
import numpy


def normalizeInput(matrix: numpy.ndarray):
    """Per-frame log min-max normalisation.  Reference :13-28: raises ValueError when a value
    is <= 0, zero-fills (in place) and returns a constant frame, otherwise returns
    (log m - log min) / (log max - log min) as a new array."""
    minvalue, maxvalue = matrix.min(), matrix.max()
    if minvalue > maxvalue:
        raise ValueError("minvalue must be less than or equal to maxvalue")
    elif minvalue <= 0:
        print(matrix.shape)
        raise ValueError("values must all be positive")
    elif minvalue == maxvalue:
        matrix.fill(0)
        return matrix
    low, high = numpy.log(minvalue), numpy.log(maxvalue)
    out = numpy.log(matrix)
    out -= low
    out /= high - low
    return out
