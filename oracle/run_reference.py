"""Import the reference's OWN source (convgp/*.py under /root/reference, unmodified, read in place) on top
of the TF1 / GPflow stand-ins in oracle/ref_shim/.  TEST INFRASTRUCTURE ONLY -- used by
tests/golden/make_ref_golden.py to generate golden vectors from the reference's code, and by
tests/test_reference_source.py to run the reference's own unit tests in the build container.
/root/reference does not exist on the GPU box: nothing that runs there imports this module."""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "ref_shim")
REFERENCE = os.environ.get("CONVGP_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE, "convgp", "convkernels.py"))


def load():
    """-> (GPflow stand-in, convgp.convkernels, convgp.misvgp, convgp.svsumgp, convgp.svconvgp) of the reference."""
    if not available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE)
    for p in (REFERENCE, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tensorflow as tf
    if not getattr(tf, "__file__", "").startswith(SHIM):
        raise RuntimeError("a real tensorflow is importable; the stand-in must not shadow it silently")
    GPflow = importlib.import_module("GPflow")
    ck = importlib.import_module("convgp.convkernels")
    mi = importlib.import_module("convgp.misvgp")
    ss = importlib.import_module("convgp.svsumgp")
    sc = importlib.import_module("convgp.svconvgp")
    for m in (ck, mi, ss, sc):
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(REFERENCE)), m.__file__
    return GPflow, ck, mi, ss, sc
