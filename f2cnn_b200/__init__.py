"""f2cnn_b200 -- B200 (sm_100a) implementation of F2CNN's feature-extraction hot path:
ERB gammatone filterbank -> ENV1 Hilbert envelope (+ Butterworth low-pass) -> windowing.

Layout:
  csrc/                     hand-written CUDA kernels + the C ABI (include/f2cnn_b200.h)
  _native.py, engine.py     ctypes binding, plans / batches (PyTorch = memory + streams only)
  api.py                    numpy-in / numpy-out entry points
  gammatone/, scripts/      drop-in modules with the reference's names and signatures
  dropin.py                 install() -> the reference's own f2cnn.py runs on this path
Importing the package needs neither the shared library nor a GPU (coefficient design and
the host-side drivers are plain Python); any call that computes on signals does, and raises
instead of falling back when they are missing."""

__version__ = "0.1.0"
