import ctypes, sys
sys.path.insert(0, '/root/repo')
from convgp_b200 import _lib
for kind, name in ((2, "ex2 alone, 16 warps/SMSP"), (18, "ex2 + 6 FFMA, 16 warps/SMSP"), (21, "ex2 + 2 FFMA, 16 warps/SMSP"), (19, "ex2 + 6 FFMA, 2 warps/SMSP"), (20, "ex2 + 6 FFMA, 3 warps/SMSP"), (1, "FFMA alone")):
    v = ctypes.c_double()
    _lib.check(_lib.lib().cgp_peak_probe(kind, ctypes.byref(v)))
    print("%-32s %.3e /s  = %.2f per clk per SM at 1.965 GHz -> %.2f cycles per warp-op per SMSP" % (name, v.value, v.value / 148 / 1.965e9, 32 * 4 / (v.value / 148 / 1.965e9)))
