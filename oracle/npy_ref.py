"""TEST INFRASTRUCTURE ONLY — the reference's numpy known-answer functions.

Loads ``npy_seg_*`` / ``grad_seg_max_pool`` / ``rand_indptr`` straight out of
/root/reference/seg_ops_cuda/mxnet_op/test_seg_ops.py:6-99.  That file imports ``mxnet`` at
module top (absent here), so an empty stand-in module is installed for the duration of the
import; the functions themselves are pure numpy.  Only usable where /root/reference exists
(fixture generation in the authoring container) — the GPU box uses tests/golden/*.npz.
"""
import importlib.util
import os
import sys
import types

REF_TEST = "/root/reference/seg_ops_cuda/mxnet_op/test_seg_ops.py"


def available():
    return os.path.exists(REF_TEST)


def load():
    stubs = {}
    for name in ("mxnet", "mxnet.ndarray", "mxnet.test_utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            stubs[name] = m
            sys.modules[name] = m
    sys.modules["mxnet.test_utils"].assert_almost_equal = None
    sys.modules["mxnet"].ndarray = sys.modules["mxnet.ndarray"]
    try:
        spec = importlib.util.spec_from_file_location("_ref_test_seg_ops", REF_TEST)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name in stubs:
            sys.modules.pop(name, None)
    return mod
