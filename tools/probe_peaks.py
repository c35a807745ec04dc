"""Print the pipe-throughput probes (roofline denominators) measured on the current GPU."""
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from convgp_b200 import _lib  # noqa: E402

NAMES = {0: "fp64_fma_per_s", 1: "fp32_fma_per_s", 2: "mufu_ex2_per_s", 3: "f64_exp_per_s",
         4: "fp64_mma_fma_per_s", 5: "fp64_mma_plus_fma_per_s",
         6: "fp64_fma_single_chain_per_s", 7: "fp64_fma_with_int_per_s", 8: "fp64_fma_with_lds_per_s",
         9: "fp64_fma_3_distinct_operands_per_s", 10: "fp64_fma_shared_multiplier_per_s",
         11: "dmma_1chain_1warp_fma_per_s", 12: "dmma_2chain_1warp_fma_per_s", 13: "dmma_4chain_1warp_fma_per_s",
         14: "dmma_distinct_A_fma_per_s", 15: "dmma_distinct_AB_fma_per_s",
         16: "tf32_mma_sync_m16n8k8_fma_per_s"}


def probe_all():
    torch.cuda.init()
    L = _lib.lib()
    out = {}
    for kind, name in NAMES.items():
        v = ctypes.c_double()
        _lib.check(L.cgp_peak_probe(kind, ctypes.byref(v)))
        out[name] = v.value
    return out


if __name__ == "__main__":
    print(json.dumps(probe_all()))
