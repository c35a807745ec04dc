"""Every kernel family of each dtype (float64: FP64-MMA / scalar FMA; float32: tcgen05 kind::f16 two-term "umma16" (the
default) / tcgen05 3xTF32 "umma" / 3xTF32 mma.sync / scalar FMA) must pass parity whatever the library's default selection is: re-run the parity, golden
and identity tests in subprocesses with the selection forced."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("family", ["umma16", "umma", "umma-dwpass", "mma", "fma"])
def test_forced_kernel_family(family):
    env = dict(os.environ)
    opts = []
    if family == "umma-dwpass":  # dW of Kuf from the stand-alone row-sum pass instead of the dZ kernel's column sums
        opts.append("kuf_dw_pass=1")
        family = "umma"
    f64 = {"umma16": 1, "umma": 1, "mma": 1, "fma": 0}[family]  # float64 has no UMMA family (no FP64 tcgen05 kind)
    f32 = {"umma16": 3, "umma": 2, "mma": 1, "fma": 0}[family]
    for k in ("kuf_fwd", "kuf_bwd", "kdiag_fwd", "kdiag_bwd"):
        opts += ["%s=%d" % (k, f64), "%s_f32=%d" % (k, f32)]
    env["CGP_TEST_OPTIONS"] = ",".join(opts)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "tests/test_gpu_parity.py",
                        "tests/test_golden.py", "tests/test_gpu_identities.py", "-p", "no:cacheprovider",
                        "-rf", "--tb=short"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    # the tail names every failing inner test (-rf) so a red variant run can be read from the outer report
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-2000:]
