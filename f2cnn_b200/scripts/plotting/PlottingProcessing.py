"""`plot gtg` behind the names of the reference's scripts/plotting/PlottingProcessing.py, fed from the
DECIMATED envelope.

The reference (file:line there) filters the file, takes the full-rate envelope -- (128, n) float64, 49 MB
for 3 s -- repeats every channel round(ERB / ERB of the lowest channel) times (:27-60: a (~1500, n) float64
image, 0.6 GB) and hands that to imshow, which resamples it to a few thousand screen columns (:63-78).
Here the fused kernel stores every `hop`-th envelope sample only (`api.gammatonegram`; the stored samples
ARE the full-rate ones, bit for bit), so that the image has about MAX_COLUMNS columns before it leaves
the device; the time axis keeps its seconds through FRAMERATE / hop.

Same names, arguments and return values as the reference; `axis=None` stands for pyplot (the reference
binds the pyplot module as the default at import time, and matplotlib is imported lazily here)."""
import os
from configparser import ConfigParser

import numpy

MAX_COLUMNS = 4096   # image columns that leave the device for a whole-file plot


def ERBScale(f):
    """Moore & Glasberg's equivalent rectangular bandwidth at centre frequency f in Hz (:18-24)."""
    return 24.7 * (4.37 * f * 0.001 + 1)


def GetNewHeightERB(matrix, CENTER_FREQUENCIES):
    """Image height if every channel is as many pixels high as its bandwidth is multiples of the lowest
    channel's (rounded), and those multiples (:27-42)."""
    base = ERBScale(CENTER_FREQUENCIES[-1])
    ratios = [int(round(ERBScale(CENTER_FREQUENCIES[i]) / base)) for i in range(len(matrix))]
    return sum(ratios), ratios


def ReshapeEnvelopesForSpectrogram(envelopes, CENTER_FREQUENCIES, start=0, end=None):
    """(:45-60) rows repeated by their ERB ratio, columns [start:end].  A ratio of 0 leaves one zero row,
    as the reference's loop does."""
    h, ratios = GetNewHeightERB(envelopes, CENTER_FREQUENCIES)
    image = numpy.zeros([h, envelopes.shape[1]])
    i = 0
    for line, ratio in zip(envelopes, ratios):
        image[i:i + ratio] = line
        i += max(ratio, 1)
    return image[:, start:end] if end is not None else image[:, start:]


def PlotEnvelopeSpectrogram(matrix, CENTER_FREQUENCIES, axis=None, LOW_FREQ=100, FRAMERATE=16000, start=0, end=None,
                            NYQUIST=None):
    """(:63-78) imshow of the reshaped envelopes with logarithmic colours; returns the image height.
    NYQUIST: top of the frequency axis when FRAMERATE is the rate of a decimated envelope."""
    from matplotlib.colors import LogNorm
    if axis is None:
        import matplotlib.pyplot as axis
    image = ReshapeEnvelopesForSpectrogram(matrix, CENTER_FREQUENCIES, start, end)
    top = int(FRAMERATE / 2) if NYQUIST is None else int(NYQUIST)
    axis.imshow(image, norm=LogNorm(), aspect="auto", extent=[start, len(image[0]) / FRAMERATE, LOW_FREQ, top])
    return len(image)


def decimated_envelopes(filename, max_columns=None):
    """-> ((C, ceil(n / hop)) float64 envelope samples at t = 0, hop, 2*hop, ..., centre frequencies, framerate,
    hop) for the 128-channel bank `plot gtg` uses (:101-105), computed by the fused kernel."""
    from ... import api
    from ...gammatone.filters import centre_freqs, make_erb_filters
    from ..processing.GammatoneFiltering import GetArrayFromWAV
    framerate, wave = GetArrayFromWAV(filename)
    cfs = centre_freqs(framerate, 128, 100)
    hop = max(1, -(-len(wave) // int(max_columns or MAX_COLUMNS)))
    return api.gammatonegram(wave, make_erb_filters(framerate, cfs), hop), cfs, framerate, hop


def PlotEnvelopesAndFormantsFromFile(filename, start=0, end=None, formantToPlot=5):
    """(:81-133) gammatonegram of a .wav file with the VTR formants of the .FB file next to it on top."""
    import matplotlib.pyplot as plt
    from ..processing.FBFileReader import ExtractFBFile
    config = ConfigParser()
    config.read('configF2CNN.conf')
    LOW_FREQ = config.getint('FILTERBANK', 'LOW_FREQ')
    sampPeriod = config.getint('CNN', 'SAMPLING_PERIOD')
    matrix, cfs, framerate, hop = decimated_envelopes(filename)
    # the reference slices image columns with start / end: sample indices there, decimated columns here
    PlotEnvelopeSpectrogram(matrix, CENTER_FREQUENCIES=cfs, LOW_FREQ=LOW_FREQ, FRAMERATE=framerate / hop,
                            start=start // hop, end=None if end is None else -(-end // hop), NYQUIST=framerate / 2)
    formants, _ = ExtractFBFile(os.path.splitext(filename)[0] + '.FB')
    print(len(formants))
    print(sampPeriod)
    if formants is not None:
        tracks = numpy.asarray(formants)[:, :4]
        t = [i * sampPeriod * 1.0 / 1000000 for i in range(len(tracks))]
        if 0 < formantToPlot < 5:
            plt.plot(t, tracks[:, formantToPlot - 1], label='F{} Frequencies (Hz)'.format(formantToPlot))
        else:
            for i in range(4):
                plt.plot(t, tracks[:, i], label='F{} Frequencies (Hz)'.format(i + 1))
        plt.legend()
        plt.text(t[-1] / 2, -700, "File:" + filename)
    plt.title("'Spectrogram' like representation of envelopes, and formants")
    plt.show()
