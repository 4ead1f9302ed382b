"""Make the reference's own code run on this package.

    import f2cnn_b200.dropin as dropin
    dropin.install()            # before `import f2cnn` / `from scripts... import ...`

install() registers this package's drop-in modules in sys.modules under the reference's
module paths:
    gammatone, gammatone.filters
    scripts.processing.GammatoneFiltering / EnvelopeExtraction / InputGenerator
    scripts.processing.LabelDataGenerator / FBFileReader / PHNFileReader
    scripts.CNN.Evaluating   (GPU front end; model prediction and plots stay the reference's)
Everything else under `scripts` (OrganiseFiles, CNN training, plotting ...) keeps
resolving to the reference tree, which must be importable (on sys.path) if those are used.
Modules of the reference that were imported BEFORE install() and bound hot-path functions by
name (`from ... import GetFilteredOutputFromArray`, Evaluating.py:19-21,
PlottingProcessing.py:13-15) are re-bound in place.
"""
import importlib
import sys

_MODULES = {
    "gammatone": "f2cnn_b200.gammatone",
    "gammatone.filters": "f2cnn_b200.gammatone.filters",
    "scripts.processing.GammatoneFiltering": "f2cnn_b200.scripts.processing.GammatoneFiltering",
    "scripts.processing.EnvelopeExtraction": "f2cnn_b200.scripts.processing.EnvelopeExtraction",
    "scripts.processing.InputGenerator": "f2cnn_b200.scripts.processing.InputGenerator",
    "scripts.processing.FBFileReader": "f2cnn_b200.scripts.processing.FBFileReader",
    "scripts.processing.PHNFileReader": "f2cnn_b200.scripts.processing.PHNFileReader",
    "scripts.processing.LabelDataGenerator": "f2cnn_b200.scripts.processing.LabelDataGenerator",
    "scripts.CNN.Evaluating": "f2cnn_b200.scripts.CNN.Evaluating",
    "scripts.plotting.PlottingProcessing": "f2cnn_b200.scripts.plotting.PlottingProcessing",
}

# names that reference modules bind with `from X import name`
_REBIND = {
    # scripts/plotting/PlottingProcessing.py:12-15
    "scripts.plotting.PlottingProcessing": {
        "centre_freqs": ("gammatone.filters", "centre_freqs"),
        "make_erb_filters": ("gammatone.filters", "make_erb_filters"),
        "ExtractEnvelopeFromMatrix": ("scripts.processing.EnvelopeExtraction", "ExtractEnvelopeFromMatrix"),
        "ExtractFBFile": ("scripts.processing.FBFileReader", "ExtractFBFile"),
        "GetFilteredOutputFromFile": ("scripts.processing.GammatoneFiltering", "GetFilteredOutputFromFile"),
        "GetArrayFromWAV": ("scripts.processing.GammatoneFiltering", "GetArrayFromWAV"),
    },
    # f2cnn.py:4-9
    "f2cnn": {
        "FilterAllOrganisedFiles": ("scripts.processing.GammatoneFiltering", "FilterAllOrganisedFiles"),
        "ExtractAllEnvelopes": ("scripts.processing.EnvelopeExtraction", "ExtractAllEnvelopes"),
        "GenerateInputData": ("scripts.processing.InputGenerator", "GenerateInputData"),
        "GenerateLabelData": ("scripts.processing.LabelDataGenerator", "GenerateLabelData"),
        "EvaluateOneWavFile": ("scripts.CNN.Evaluating", "EvaluateOneWavFile"),
        "EvaluateRandom": ("scripts.CNN.Evaluating", "EvaluateRandom"),
        "EvaluateWithNoise": ("scripts.CNN.Evaluating", "EvaluateWithNoise"),
        "PlotEnvelopesAndFormantsFromFile": ("scripts.plotting.PlottingProcessing", "PlotEnvelopesAndFormantsFromFile"),
    },
    # scripts/plotting/PlottingCNN.py:12
    "scripts.plotting.PlottingCNN": {
        "ReshapeEnvelopesForSpectrogram": ("scripts.plotting.PlottingProcessing", "ReshapeEnvelopesForSpectrogram"),
        "PlotEnvelopeSpectrogram": ("scripts.plotting.PlottingProcessing", "PlotEnvelopeSpectrogram"),
    },
}

_installed = {}


def _ensure_parent_packages():
    """`scripts` and `scripts.processing` come from the reference tree when it is importable
    (so that LabelDataGenerator, CNN, plotting ... keep resolving); otherwise this package's
    own `scripts` packages stand in, so that the hot-path modules import on their own."""
    for ref_pkg, ours in (("scripts", "f2cnn_b200.scripts"), ("scripts.processing", "f2cnn_b200.scripts.processing"),
                          ("scripts.CNN", "f2cnn_b200.scripts.CNN"), ("scripts.plotting", "f2cnn_b200.scripts.plotting")):
        if ref_pkg in sys.modules:
            continue
        try:
            importlib.import_module(ref_pkg)
        except ImportError:
            mod = importlib.import_module(ours)
            sys.modules[ref_pkg] = mod
            _installed.setdefault(ref_pkg, None)
            parent, _, leaf = ref_pkg.rpartition(".")
            if parent and parent in sys.modules:
                setattr(sys.modules[parent], leaf, mod)


def install():
    """Idempotent.  Returns the dict {reference module path: drop-in module}."""
    _ensure_parent_packages()
    displaced = {}
    for ref_name, ours in _MODULES.items():
        mod = importlib.import_module(ours)
        prev = sys.modules.get(ref_name)
        if prev is not None and prev is not mod:
            _installed.setdefault(ref_name, prev)
            displaced[ref_name] = prev
        sys.modules[ref_name] = mod
        parent, _, leaf = ref_name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], leaf, mod)
    for importer, names in _REBIND.items():
        # also the module object this call displaced: whoever imported it earlier still holds its functions
        for m in (sys.modules.get(importer), displaced.get(importer)):
            if m is None:
                continue
            for attr, (src, name) in names.items():
                if hasattr(m, attr):
                    setattr(m, attr, getattr(sys.modules[src], name))
    return {k: sys.modules[k] for k in _MODULES}


def uninstall():
    """Restore whatever install() displaced (used by tests)."""
    for ref_name in list(_MODULES) + ["scripts.plotting", "scripts.CNN", "scripts.processing", "scripts"]:
        prev = _installed.pop(ref_name, None)
        if prev is not None:
            sys.modules[ref_name] = prev
        elif sys.modules.get(ref_name) is not None and sys.modules[ref_name].__name__.startswith("f2cnn_b200"):
            del sys.modules[ref_name]
