"""`prepare filter` + `prepare envelope` on N synthetic 3 s files in a temporary tree (tmpfs when
available): wall time of the drop-in drivers, and of the CPU arm doing the same job -- the float64 C port
of the reference algorithm (oracle/), one process per core over files like the reference's Pool
(GammatoneFiltering.py:121-125, EnvelopeExtraction.py:144-149), .npy output included."""
import os, sys, time, shutil, tempfile
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
CONF = ("[FILTERBANK]\nFRAMERATE=16000\nNCHANNELS=128\nLOW_FREQ=100\n"
        "[CNN]\nFORMANT=2\nCENTERED=True\nRADIUS=5\nBATCH_SIZE=32\nEPOCHS=20\nRISK=0.05\nSAMPLING_PERIOD=10000\n")


def cpu_one(path):
    from oracle import oracle as orc
    from scipy.io import wavfile
    from f2cnn_b200.gammatone import filters
    orc.set_num_threads(1)
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
    _, w = wavfile.read(path)
    g = orc.erb_filterbank(w, co)
    np.save(os.path.splitext(path)[0] + ".CPU_GFB", g)
    e = orc.extract_envelope(g, True, 50)
    np.save(os.path.splitext(path)[0] + ".CPU_ENV1", e)
    return len(w)


def main():
    from scipy.io import wavfile
    from f2cnn_b200 import synth
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="f2drv_", dir=base)
    try:
        os.chdir(tmp)
        open("configF2CNN.conf", "w").write(CONF)
        d = os.path.join("resources", "f2cnn", "TRAIN")
        os.makedirs(d)
        lengths = synth.corpus_lengths(N, seed=1)
        for i, n in enumerate(lengths):
            wavfile.write(os.path.join(d, "DR1.S%04d.SX1.WAV" % i), 16000, synth.white_noise_i16(int(n), seed=i))
        total = 128.0 * float(lengths.sum())
        from f2cnn_b200 import dropin
        dropin.install()
        from scripts.processing import EnvelopeExtraction, GammatoneFiltering
        import io, contextlib
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            GammatoneFiltering.FilterAllOrganisedFiles()      # first call: context, plan, pinned slots
            for f in os.listdir(d):
                if f.endswith(".npy"):
                    os.remove(os.path.join(d, f))
            t0 = time.perf_counter()
            GammatoneFiltering.FilterAllOrganisedFiles()
            t1 = time.perf_counter()
            EnvelopeExtraction.ExtractAllEnvelopes(True, 50)
            t2 = time.perf_counter()
        print("drop-in drivers, %d files (%.1f GB of .npy per stage) in %s: prepare filter %.2f s, prepare envelope %.2f s "
              "-> %.3e channel-samples/s over both stages" % (N, total * 8 / 1e9, tmp, t1 - t0, t2 - t1, total / (t2 - t0)))
        cores = len(os.sched_getaffinity(0))
        files = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".WAV"))
        sub = files[:max(cores * 2, 32)]
        t0 = time.perf_counter()
        with ProcessPoolExecutor(max_workers=cores) as pool:
            done = sum(pool.map(cpu_one, sub))
        t = time.perf_counter() - t0
        print("CPU arm (C port of the reference, %d processes over files, .npy output): %d files in %.2f s -> %.3e channel-samples/s"
              % (cores, len(sub), t, 128.0 * done / t))
    finally:
        os.chdir("/")
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
