"""Every kernel family of each dtype (float64: FP64-MMA / scalar FMA; float32: tcgen05 UMMA / 3xTF32 mma.sync /
scalar FMA) must pass parity whatever the library's default selection is: re-run the parity, golden
and identity tests in subprocesses with the selection forced."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("family", ["umma", "umma-dwpass", "mma", "fma"])
def test_forced_kernel_family(family):
    env = dict(os.environ)
    if family == "umma-dwpass":  # dW of Kuf from the stand-alone row-sum pass instead of the dZ kernel's column sums
        env["CGP_KUF_DW_PASS"] = "1"
        family = "umma"
    for k in ("CGP_KUF_FWD", "CGP_KUF_BWD", "CGP_KDIAG_FWD", "CGP_KDIAG_BWD"):
        env[k] = "mma" if family == "umma" else family  # float64 has no UMMA family (no FP64 tcgen05 kind)
        env[k + "_F32"] = family
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "tests/test_gpu_parity.py",
                        "tests/test_golden.py", "tests/test_gpu_identities.py", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
