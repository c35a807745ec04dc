"""``GPflow.settings`` as configured for the reference's float64 runs (testing/test_svsumgp.py:9 forces
float64; jitter_level from gpflowrc:11)."""
import torch


class _NS:
    pass


dtypes = _NS()
dtypes.float_type = torch.float64
dtypes.int_type = torch.int32
numerics = _NS()
numerics.jitter_level = 1e-5
numerics.float_type = torch.float64
