"""Drop-in for the reference's gammatone/filters.py: same names, signatures and results.

Coefficient design (erb_point / erb_space / centre_freqs / make_erb_filters) stays on the
host in float64 -- it is 0.2 ms of work whose output is the kernels' parameter block, and
it reproduces the reference bit for bit (tests/test_host_cpu.py::test_filters_bit_exact_vs_reference_golden compares with `==`
against tests/golden/coefs.npz).  erb_filterbank runs on the B200 through
libf2cnn_b200.so; there is no CPU implementation of it in this package.

Reference citations are file:line under tictacmenthe/F2CNN.
"""
import numpy as np

DEFAULT_FILTER_NUM = 100        # gammatone/filters.py:16
DEFAULT_LOW_FREQ = 100          # gammatone/filters.py:17
DEFAULT_HIGH_FREQ = 44100 / 4   # gammatone/filters.py:18

# Glasberg & Moore ERB-scale parameters (gammatone/filters.py:35-37, 136-138)
_EAR_Q = 9.26449
_MIN_BW = 24.7
_ORDER = 1


def erb_point(low_freq, high_freq, fraction):
    """Point on the ERB scale between high_freq (fraction 0) and low_freq (fraction 1).
    gammatone/filters.py:21-52."""
    offset = _EAR_Q * _MIN_BW
    log_span = -np.log(high_freq + offset) + np.log(low_freq + offset)
    return -offset + np.exp(fraction * log_span) * (high_freq + offset)


def erb_space(low_freq=DEFAULT_LOW_FREQ, high_freq=DEFAULT_HIGH_FREQ, num=DEFAULT_FILTER_NUM):
    """`num` frequencies uniformly spaced on the ERB scale, descending from just below
    high_freq to low_freq.  gammatone/filters.py:55-71."""
    return erb_point(low_freq, high_freq, np.arange(1, num + 1) / num)


def centre_freqs(fs, num_freqs, cutoff):
    """Centre frequencies for make_erb_filters: ERB-spaced between fs/2 and cutoff.
    gammatone/filters.py:74-86."""
    return erb_space(cutoff, fs / 2, num_freqs)


def make_erb_filters(fs, centre_freqs, width=1.0):
    """Slaney's 4th-order gammatone as four second-order sections sharing one denominator.
    Returns (C,10) float64 rows [A0, A11, A12, A13, A14, A2, B0, B1, B2, gain].
    gammatone/filters.py:89-192 (expression order kept so that results are bit-identical)."""
    T = 1 / fs
    erb = width * ((centre_freqs / _EAR_Q) ** _ORDER + _MIN_BW ** _ORDER) ** (1 / _ORDER)
    B = 1.019 * 2 * np.pi * erb
    arg = 2 * centre_freqs * np.pi * T
    vec = np.exp(2j * arg)

    B1 = -2 * np.cos(arg) / np.exp(B * T)
    B2 = np.exp(-2 * B * T)

    rt_pos = np.sqrt(3 + 2 ** 1.5)
    rt_neg = np.sqrt(3 - 2 ** 1.5)
    common = -T * np.exp(-(B * T))
    cos_a, sin_a = np.cos(arg), np.sin(arg)
    k = [cos_a + rt_pos * sin_a, cos_a - rt_pos * sin_a, cos_a + rt_neg * sin_a, cos_a - rt_neg * sin_a]
    A1 = [common * kk for kk in k]

    gain_arg = np.exp(1j * arg - B * T)
    gain = np.abs(
        (vec - gain_arg * k[0]) * (vec - gain_arg * k[1]) * (vec - gain_arg * k[2]) * (vec - gain_arg * k[3])
        * (T * np.exp(B * T) / (-1 / np.exp(B * T) + 1 + vec * (1 - np.exp(B * T)))) ** 4)

    ones = np.ones_like(centre_freqs)
    return np.column_stack([T * ones, A1[0], A1[1], A1[2], A1[3], 0 * ones, 1 * ones, B1, B2, gain])


def erb_filterbank(wave, coefs):
    """Filter a 1-D waveform with the gammatone bank: (C, len(wave)) float64, one channel
    per row, zero initial state.  gammatone/filters.py:195-239 -- on the GPU."""
    from .. import api
    return api.erb_filterbank(wave, coefs)
