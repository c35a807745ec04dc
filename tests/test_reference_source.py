"""Build-container check: the reference's OWN unit tests (/root/reference/testing/test_convkerns.py, imported
unmodified) pass when its OWN kernel code (/root/reference/convgp/*.py) runs on the stand-ins in oracle/ref_shim/.
That is the evidence that the stand-ins give the reference's code the semantics its authors tested for, and hence
that tests/golden/ref/*.npz are the reference's outputs.  Skipped where /root/reference is absent (the GPU box)."""
import importlib
import sys
import unittest

import pytest

from oracle import run_reference as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference sources not present")


def test_reference_unit_tests_pass_on_the_stand_ins():
    R.load()
    mod = importlib.import_module("testing.test_convkerns")
    assert mod.__file__.startswith(R.REFERENCE)
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    assert suite.countTestCases() == 9
    result = unittest.TextTestRunner(stream=sys.stderr, verbosity=0).run(suite)
    assert result.wasSuccessful(), (result.failures, result.errors)


def test_fixtures_are_reproducible_from_the_reference_source():
    """Regenerate one fixture from the reference's code and compare with the committed file bitwise-close."""
    import os
    import numpy as np
    from tests.golden import make_ref_golden as G
    GP, ck, mi, ss, sc = R.load()
    case = G.prepare([c for c in G.kernel_cases() if c.name == "ref_cfg3_mnist_wconv"][0])
    out = G.run_kernel_case(case, GP, ck, mi)
    ref = np.load(os.path.join(G.OUT, case.name + ".npz"))
    for key in ("kuf", "kdiag", "kuu", "dZ", "dW", "dvar", "dls"):
        assert np.allclose(out[key], ref[key], rtol=1e-13, atol=0), key
