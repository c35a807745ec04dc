"""Generate the golden fixtures under tests/golden/ from the float64 oracle (values) and its torch
twin (gradients).  Run from the repo root:  python tests/golden/make_golden.py

The reference itself cannot run in this image (TensorFlow 1.x + GPflow-inter-domain) and
ships no golden vectors, so these are outputs of OUR restatement of its algorithm with fixed seeds:
they pin the oracle against regressions and give the GPU tests a fixture that does not depend on
the CPU maths libraries of the GPU box.  Shapes are the five BASELINE configs scaled down in N / M
(the image / patch geometry and kernel structure are the real ones)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import Case  # noqa: E402

S = np.s_


def golden_cases():
    return [
        # cfg 1: rectangles.py -k conv  (28x28, 3x3, Conv(RBF(9)))
        Case("cfg1_rectangles_conv", "conv", 28, 28, 1, 3, 3, [(1.0, 1.0, None, False)], weighted=False, N=3, M=16,
             seed=101),
        # cfg 2/3: mnist01.py / mnist.py -k wconv (28x28, 5x5, WeightedConv(RBF(25)))
        Case("cfg3_mnist_wconv", "conv", 28, 28, 1, 5, 5, [(1.0, 1.0, None, False)], N=3, M=24, seed=103),
        Case("cfg3_mnist_wconv_l2p34", "conv", 28, 28, 1, 5, 5, [(0.001, 2.34, None, False)], N=2, M=12, seed=104,
             xdist="randn"),
        # cfg 5: cifar.py -k addwconv (32x32x3, 5x5x3 colour patches, one RBF per channel)
        Case("cfg5_cifar_addwconv", "colour", 32, 32, 3, 5, 5,
             [(1.004, 0.5, S[0::3], False), (0.991, 0.5, S[1::3], False), (1.013, 0.5, S[2::3], False)],
             N=2, M=16, seed=105, xdist="uint8"),
        # cifar.py -k wconv / multi (32x32x3, per-channel 5x5 patches)
        Case("cifar_wconv_c3", "conv", 32, 32, 3, 5, 5, [(1.0, 1.0, None, False)], N=2, M=12, seed=106, xdist="uint8"),
        Case("cifar_multi_c3", "multi", 32, 32, 3, 5, 5, [(1.0, 1.0, None, False)], N=2, M=12, seed=107, xdist="uint8"),
    ]


def rectangles_rows(n):
    """First n training images of datasets/rectangles.zip (the only dataset the reference ships;
    values in {0,1}, datasets/process_rectangles.py:1-20).  Read in the build container only."""
    import zipfile
    z = zipfile.ZipFile("/root/reference/datasets/rectangles.zip")
    rows = []
    with z.open("rectangles_train.amat") as f:
        for _ in range(n):
            rows.append([float(v) for v in f.readline().split()][:784])
    return np.array(rows)


def save_case(case, path):
    ok = case.oracle_kernel()
    ref = case.twin_grads()
    kuf, kd, kuu = ok.Kzx(case.Z, case.X), ok.Kdiag(case.X), ok.Kzz(case.Z)
    assert np.max(np.abs(ref["kuf"] - kuf)) < 1e-13 * max(1.0, np.max(np.abs(kuf)))
    np.savez_compressed(
        path, X=case.X, Z=case.Z, W=np.ones(case.wshape()) if case.Wv is None else case.Wv, G=case.G, gd=case.gd,
        Gu=case.Gu, kuf=kuf, kdiag=kd, kuu=kuu, dZ=ref["dZ"], dW=np.zeros(0) if ref["dW"] is None else ref["dW"],
        dvar=np.concatenate([np.ravel(v) for v in ref["dvar"]]),
        dls=np.concatenate([np.ravel(v) for v in ref["dls"]]))
    return kuf


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    # reference outputs of every parametrised parity case: the GPU tests then do not depend on the
    # CPU maths libraries of the GPU box (observed to drift by ~1e-9 on some hosts)
    from tests.test_gpu_parity import CASES as PARITY_CASES, full_size_case
    os.makedirs(os.path.join(out_dir, "parity"), exist_ok=True)
    for case in PARITY_CASES:
        save_case(case, os.path.join(out_dir, "parity", case.name + ".npz"))
        print("parity", case.name)
    big = full_size_case()
    okb = big.oracle_kernel()
    np.savez_compressed(os.path.join(out_dir, "parity", "cfg3_full_subsample.npz"),
                        kuf=okb.Kzx(big.Z, big.X)[::15, ::10], kdiag=okb.Kdiag(big.X)[::10], kuu=okb.Kzz(big.Z)[::15, ::15])
    from tests.test_gpu_parity import full_size_cifar_cases
    for c in full_size_cifar_cases():
        okc = c.oracle_kernel()
        np.savez_compressed(os.path.join(out_dir, "parity", c.name + "_subsample.npz"),
                            kuf=okc.Kzx(c.Z, c.X)[::25, ::2], kdiag=okc.Kdiag(c.X))
        print("parity", c.name)
    cases = golden_cases()
    if os.path.exists("/root/reference/datasets/rectangles.zip"):
        c = cases[0]
        c.X = rectangles_rows(c.N)
        rng = np.random.RandomState(c.seed)
        pat = c.oracle_kernel(np.ones(c.wshape()))._get_patches(c.X).reshape(-1, c.geom.D)
        c.Z = pat[rng.permutation(len(pat))[:c.M]] + 0.001 * rng.rand(c.M, c.geom.D)
    for case in cases:
        kuf = save_case(case, os.path.join(out_dir, case.name + ".npz"))
        print(case.name, kuf.shape, float(np.abs(kuf).max()))


if __name__ == "__main__":
    main()
