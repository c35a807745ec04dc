"""convgp_b200 -- B200-native implementation of the convgp patch-kernel hot path.

Mirrors the import surface of the reference package (convgp/__init__.py:16-17 imports ``svsumgp``
and ``convkernels``) and adds the kernel/model modules the reference scripts pull in directly."""
from .settings import settings  # noqa: F401
from . import convkernels  # noqa: F401
from . import kernels  # noqa: F401
from . import likelihoods  # noqa: F401,E402
from . import misvgp  # noqa: F401,E402
from . import svgp  # noqa: F401,E402
from . import svsumgp  # noqa: F401,E402
from . import svconvgp  # noqa: F401,E402
