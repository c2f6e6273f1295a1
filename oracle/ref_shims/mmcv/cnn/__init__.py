import torch.nn as nn


class ConvModule(nn.Module):
    """conv(no bias) -> BatchNorm2d -> ReLU, attribute names as in mmcv 2.x (conv, bn, activate)."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0, bias=False, groups=1,
                 norm_cfg=None, act_cfg=None, **kw):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, padding=padding, bias=bias, groups=groups)
        self.bn = nn.BatchNorm2d(out_channels)
        self.activate = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.activate(self.bn(self.conv(x)))
