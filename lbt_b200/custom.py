"""The reference's PyTorch-style LeNet (``custom.py:N`` = /root/reference/custom.py:N) over lbt_b200.dfxp.

custom.py is an orphan in the reference (it imports a `.dfxp` module that does not exist there) and is
the only place that shows the PyTorch call signatures of the quantised layers (custom.py:11-12, 29-30); this file is
that module running against the real layers.  Two defects of the original are repaired, both noted in SURVEY App. E-5:
`nn.ReLU(out)` constructs a module instead of applying one (custom.py:36), and with padding=1 a 28x28 input reaches the
classifier with 120*1*1 features only for 28x28 inputs (kept: the flatten adapts to whatever the features produce).
"""
import torch.nn as nn

from . import dfxp
from .dfxp import Conv2d_q, Linear_q

__all__ = ['custom']


def conv5x5(bits, in_channels, out_channels, stride=1, **kw):
    return Conv2d_q(bits, in_channels, out_channels, kernel_size=5, stride=stride, padding=1, bias=False, **kw)   # custom.py:10-12


class CUSTOM_MNIST(nn.Module):
    cfg = {'custom': [6, 'M', 16, 'M', 120]}                                                                      # custom.py:16-21

    def __init__(self, bits, custom_name, seed=0):
        super().__init__()
        self.bits = bits
        self.runtime = dfxp.Runtime(seed)
        self.features = self._make_layers(self.cfg[custom_name])
        self.fc1 = None                     # in_features depends on the input size (120 * h * w): built on first use
        self.fc2 = Linear_q(bits, 84, 10, runtime=self.runtime, name='fc2')                                       # custom.py:30

    def forward(self, x):
        out = self.features(x)
        out = out.permute(0, 2, 3, 1).reshape(out.size(0), -1)
        if self.fc1 is None:
            self.fc1 = Linear_q(self.bits, out.shape[1], 84, runtime=self.runtime, name='fc1').to(out.device)     # custom.py:29
        out = dfxp.ReLU_q()(self.fc1(out))                                                                              # custom.py:35-36
        return self.fc2(out)

    def _make_layers(self, cfg):
        layers, in_channels = [], 1
        for i, x in enumerate(cfg):
            if x == 'M':
                layers.append(dfxp.MaxPool_q(2, 2, 'VALID'))                                                      # custom.py:45
            else:
                layers += [conv5x5(self.bits, in_channels, x, runtime=self.runtime, name='conv%d' % i,
                                   input_signed=(in_channels == 1)), dfxp.ReLU_q()]
                in_channels = x
        return nn.Sequential(*layers)

    def ranges(self):
        return self.runtime.ranges()


def custom(bits, **kw):
    """custom.py:52-53."""
    return CUSTOM_MNIST(bits, 'custom', **kw)
