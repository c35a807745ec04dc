"""Global numerics settings -- the two values of the reference's ``gpflowrc`` that are part of the
parity contract (gpflowrc:6-11): the float type and the Cholesky jitter.  In the reference the
dtype "has to be adjusted manually in the gpflowrc file" (README.md:34-37); here it is a setting
that kernels and models read at construction."""
import torch


class _Settings:
    def __init__(self):
        self.float_type = torch.float64
        self.jitter_level = 1e-5  # gpflowrc:11
        self.check_cholesky = True  # raise on a non-PD Kmm (costs one host sync per step; off inside CUDA graphs)
        self._device = None

    @property
    def device(self):
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        return self._device

    @device.setter
    def device(self, dev):
        self._device = torch.device(dev)


settings = _Settings()
