"""Import shim (TEST INFRASTRUCTURE): the reference's logger constructs a tensorboardX.SummaryWriter
(improved_diffusion/logger.py:161-162) and only ever calls add_scalar on it.  tensorboardX is not in this image."""


class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag=None, scalar_value=None, global_step=None, **k):
        self.scalars.append((tag, scalar_value, global_step))

    def flush(self):
        pass

    def close(self):
        pass

    Close = close
