"""TIMIT phoneme (.PHN) files, behind the names of the reference's scripts/processing/PHNFileReader.py.

A .PHN line is "<first sample> <last sample> <phoneme>" (space separated); ExtractPhonemes returns
[(phoneme, first, last)] in file order, or None after printing when the file is missing (reference
:20-30); GetPhonemeFromArrayAt returns the FIRST entry whose closed interval holds the timepoint
and 'h#' when none does (:33-37).  phoneme_at() is the same lookup for an array of timepoints.
"""
import numpy

# TIMIT phone classes, same members and order as the reference's lists (:10-17)
STOPS = 'b d g p t k dx q'.split()
AFFRICATIVES = 'jh ch'.split()
FRICATIVES = 's sh w wh f th v dh'.split()
NASALS = 'm n ng em en eng nx'.split()
SEMIVOWELS_AND_GLIDES = 'l r w y hh hv el'.split()
VOWELS = 'iy ih eh ey ae aa aw ay ah ao oy ow uh uw ux er ax ix axr ax-h'.split()
SILENTS = 'pau epi h#'.split()


def ExtractPhonemes(phnFilename):
    try:
        with open(phnFilename, 'r') as handle:
            lines = handle.read().split('\n')
    except FileNotFoundError:
        print("No .PHN phoneme data file.")
        return None
    segments = []
    for line in lines:
        if not line:
            continue  # csv.reader skips empty lines too
        first, last, phoneme = line.split(' ')[:3]
        segments.append((phoneme, int(first), int(last)))
    return segments


def GetPhonemeFromArrayAt(phonemes, timepoint):
    for phoneme, first, last in phonemes:
        if first <= timepoint <= last:
            return phoneme
    return 'h#'


def phoneme_at(phonemes, timepoints):
    """GetPhonemeFromArrayAt for every timepoint: list of phoneme strings."""
    t = numpy.asarray(timepoints)
    which = numpy.full(t.shape, -1, dtype=numpy.int64)
    for k, (_, first, last) in enumerate(phonemes):
        hit = (which < 0) & (first <= t) & (t <= last)
        which[hit] = k
    return [phonemes[k][0] if k >= 0 else 'h#' for k in which]


def GetPhonemeAt(phnFilename, timepoint):
    return GetPhonemeFromArrayAt(ExtractPhonemes(phnFilename), timepoint)
