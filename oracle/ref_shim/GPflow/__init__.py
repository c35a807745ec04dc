"""A stand-in for the parts of ``GPflow`` (github.com/markvdw/GPflow-inter-domain, GPflow 0.x; un-vendored and
unpinned by the reference) that the reference's convgp/*.py files import.  TEST INFRASTRUCTURE ONLY.

It lets the reference's own source run (see ../tensorflow/__init__.py).  What is restated HERE is the
published GPflow 0.x algorithm of each piece (``RBF.K`` with the expanded squared distance, ``White``,
``Add``, ``conditional``, the whitened Gaussian KL, ``SVGP.build_likelihood``, the Gaussian likelihood) --
those remain *parity unpinned* by anything in /root/reference; everything under convgp/ is the
reference's code itself.
"""
from . import settings, transforms, param, mean_functions, kernels, likelihoods, kullback_leiblers  # noqa: F401
from . import conditionals, svgp  # noqa: F401
