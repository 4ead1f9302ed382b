"""Batched WAV ingest: many RIFF / NIST SPHERE files -> ONE pinned
int16 buffer + lengths, ready for a single H2D copy.

The reference decodes file by file (`GetArrayFromWAV`, scripts/processing/GammatoneFiltering.py:28-39:
scipy.io.wavfile for RIFF, an interpreter loop over `sphfile` samples otherwise) and label
generation decodes every file a second time just to learn its length
(LabelDataGenerator.py:38-40).  Here the headers are parsed once (`wav_layout`), the payloads are
read by a few threads straight into their slices of one page-locked buffer (`readinto`, no
intermediate arrays; big-endian SPHERE payloads are byte-swapped in place), and the buffer goes to
the device as one copy.
"""
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


class WavLayout:
    __slots__ = ("rate", "samples", "offset", "sample_bytes", "channels", "big_endian", "pcm")

    def __init__(self, rate, samples, offset, sample_bytes, channels, big_endian, pcm):
        self.rate, self.samples, self.offset = int(rate), int(samples), int(offset)
        self.sample_bytes, self.channels = int(sample_bytes), int(channels)
        self.big_endian, self.pcm = bool(big_endian), bool(pcm)

    @property
    def is_int16_mono(self):
        return self.pcm and self.sample_bytes == 2 and self.channels == 1


def wav_layout(path):
    """Header of a RIFF/WAVE or NIST SPHERE file: where the samples are and what they look like."""
    with open(path, 'rb') as f:
        head = f.read(12)
        if head[:4] == b'RIFF':
            rate = block = channels = fmt_tag = None
            while True:
                chunk = f.read(8)
                if len(chunk) < 8:
                    raise ValueError("no data chunk in {}".format(path))
                tag, size = chunk[:4], struct.unpack('<I', chunk[4:])[0]
                if tag == b'fmt ':
                    fmt = f.read(size + (size & 1))
                    fmt_tag, channels, rate = struct.unpack('<HHI', fmt[:8])
                    block = struct.unpack('<H', fmt[12:14])[0]
                    if fmt_tag == 0xFFFE and len(fmt) >= 26:  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID
                        fmt_tag = struct.unpack('<H', fmt[24:26])[0]
                elif tag == b'data':
                    if rate is None:
                        raise ValueError("data chunk before fmt chunk in {}".format(path))
                    return WavLayout(rate, size // block, f.tell(), block // channels, channels, False, fmt_tag == 1)
                else:
                    f.seek(size + (size & 1), 1)
        if not head.startswith(b'NIST_1A'):
            raise ValueError("{} is neither RIFF nor NIST SPHERE".format(path))
        f.seek(0)
        f.readline()
        header_bytes = int(f.readline().strip())
        text = f.read(header_bytes - 16).decode('ascii', 'replace')
    fields = {}
    for line in text.splitlines():
        tokens = line.split(None, 2)
        if tokens and tokens[0] == 'end_head':
            break
        if len(tokens) == 3:
            fields[tokens[0]] = tokens[2]
    return WavLayout(fields['sample_rate'], fields['sample_count'], header_bytes, fields.get('sample_n_bytes', 2),
                     fields.get('channel_count', 1), fields.get('sample_byte_format', '01') != '01',
                     fields.get('sample_coding', 'pcm') == 'pcm')


def read_corpus(paths, threads=8, pinned=True):
    """(flat int16 torch tensor of all samples, int64 lengths, frame rates).  Files must be mono 16-bit
    PCM (RIFF or SPHERE), which is what TIMIT and the reference's OrganiseFiles produce."""
    layouts = [wav_layout(p) for p in paths]
    for p, lay in zip(paths, layouts):
        if not lay.is_int16_mono:
            raise ValueError("{}: read_corpus takes mono 16-bit PCM; decode it with GetArrayFromWAV".format(p))
    lengths = np.asarray([lay.samples for lay in layouts], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lengths)])
    total = int(offsets[-1])
    flat = torch.empty(max(total, 1), dtype=torch.int16, pin_memory=bool(pinned) and torch.cuda.is_available())[:total]
    view = flat.numpy()

    def load(i):
        lay = layouts[i]
        dst = view[offsets[i]:offsets[i + 1]]
        with open(paths[i], 'rb') as f:
            f.seek(lay.offset)
            got = f.readinto(memoryview(dst).cast('B'))
        if got != 2 * lay.samples:
            raise ValueError("{}: header promises {} samples, file holds {}".format(paths[i], lay.samples, got // 2))
        if lay.big_endian:
            dst.byteswap(inplace=True)

    if len(paths) > 1 and threads > 1:
        with ThreadPoolExecutor(max_workers=int(threads)) as pool:
            list(pool.map(load, range(len(paths))))
    else:
        for i in range(len(paths)):
            load(i)
    return flat, lengths, [lay.rate for lay in layouts]
