"""INTEGRATION.md section 2 is what a maintainer of the reference would copy: the ctypes block there
is executed VERBATIM.  CPU: it loads the library and its RunArgs equals the binding's, field for
field (a short struct would be read past its end -- ABI 6 rejects it by struct_size).  GPU: the
function it defines reproduces gammatone/filters.py:195-239 + EnvelopeExtraction.py:51-67."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _snippet_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    block = [b for b in blocks if "class RunArgs" in b]
    assert len(block) == 1
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)  # the snippet opens the library by its path relative to the repository root
    try:
        exec(compile(block[0], "INTEGRATION.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    return ns


def test_snippet_struct_equals_the_binding():
    from f2cnn_b200 import _native
    ns = _snippet_namespace()
    theirs, ours = ns["RunArgs"], _native.RunArgs
    assert [(n, t) for n, t in theirs._fields_] == [(n, t) for n, t in ours._fields_]
    assert ctypes.sizeof(theirs) == ctypes.sizeof(ours)
    assert ours().struct_size == ctypes.sizeof(ours)
    assert ns["L"].f2_abi_version() == _native.ABI_VERSION
    # the header and the binding agree on the layout: offsets of a C compile of the header
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "f2cnn_b200.h"\nint main(void){printf("%zu", sizeof(f2_run_args));' + \
          "".join('printf(" %%zu", offsetof(f2_run_args, %s));' % n for n, _ in ours._fields_) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        nums = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert nums[0] == ctypes.sizeof(ours)
    assert nums[1:] == [getattr(ours, n).offset for n, _ in ours._fields_]


@pytest.mark.gpu
def test_snippet_runs_and_matches_the_oracle(oracle):
    from f2cnn_b200 import synth
    from f2cnn_b200.gammatone import filters
    ns = _snippet_namespace()
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 128, 100))
    w = synth.white_noise_i16(12000, seed=3)
    gfb, env = ns["filterbank_and_envelope"](w, co, True, 50)
    go = oracle.erb_filterbank(w, co)
    eo = oracle.extract_envelope(go, True, 50)
    rel = lambda got, want: np.max(np.max(np.abs(got - want), axis=1) / np.sqrt(np.mean(want ** 2, axis=1)))
    assert rel(gfb, go) <= 1e-4 and rel(env, eo) <= 1e-4


@pytest.mark.gpu
def test_short_run_args_struct_is_rejected():
    """An ABI-5 caller (no struct_size, 15 fields): its first member is a pointer, so struct_size reads
    as garbage or 0 -- either way the library must answer F2_ERR_INVALID, not read past the end."""
    import torch
    from f2cnn_b200 import _native, engine
    from f2cnn_b200.gammatone import filters
    co = filters.make_erb_filters(16000, filters.centre_freqs(16000, 32, 100))
    plan = engine.Plan(co)
    batch = plan.batch([4000])
    a = _native.RunArgs()
    a.struct_size = ctypes.sizeof(_native.RunArgs) - 8
    a.wave = torch.zeros(4000, dtype=torch.int16, device="cuda").data_ptr()
    ws = plan.workspace(batch.workspace_bytes())
    rc = _native.lib().f2_batch_run(batch._h, ctypes.byref(a), ctypes.c_void_p(ws.data_ptr()), ws.numel(), None)
    assert rc == _native.F2_ERR_INVALID
    assert b"struct_size" in _native.lib().f2_last_error()
