"""lbt_b200 — the dynamic-fixed-point (DFXP) training hot path of freudh/lbt, built for B200 (sm_100a).

Same names as the reference's Python surface (``dynamic_fixed_point.py``, ``models.py``, ``trainer.py``, ``custom.py``); all
tensor arithmetic runs in ``liblbt_b200.so`` (hand-written CUDA behind the C ABI of ``include/lbt.h``) — importing this
package does not need a GPU, calling into it does (there is no CPU or PyTorch fallback).
"""
from ._lib import LbtError
from .quantizer import overflow_rate, update_range, weight_quantization
from .dfxp import (AvgPool_q, BatchNorm2d_q, BatchNorm_q, Conv2d_pq, Conv2d_q, Dense_q, Dropout_q, Flatten_q, GradientBuffer_q,
                   Linear_q, MaxPool_q, Normalization_q, ReLU_q, Rescale_q, ResidualBlock_q, ResidualBottleneck_q, Runtime,
                   Sequential_q, softmax_cross_entropy)
from .models import (CIFAR10_Model, CIFAR10_Resnet20, CIFAR10_Resnet32, CIFAR10_Resnet44, CIFAR10_Resnet56, CIFAR10_VGG_Model,
                     MNIST_Model, Model, PI_MNIST_Model, Resnet18, Resnet50)
from .trainer import HostFeeder, Trainer
from .data import Pipeline
from .custom import custom

__all__ = [
    'LbtError', 'weight_quantization', 'overflow_rate', 'update_range',
    'Conv2d_q', 'Conv2d_pq', 'Linear_q', 'Dense_q', 'BatchNorm2d_q', 'BatchNorm_q', 'Normalization_q', 'Rescale_q',
    'ResidualBlock_q', 'ResidualBottleneck_q', 'ReLU_q', 'MaxPool_q', 'AvgPool_q', 'Dropout_q', 'Flatten_q', 'Sequential_q',
    'GradientBuffer_q', 'Runtime', 'softmax_cross_entropy',
    'Model', 'PI_MNIST_Model', 'MNIST_Model', 'CIFAR10_Model', 'CIFAR10_VGG_Model', 'CIFAR10_Resnet20', 'CIFAR10_Resnet32',
    'CIFAR10_Resnet44', 'CIFAR10_Resnet56', 'Resnet18', 'Resnet50',
    'Trainer', 'HostFeeder', 'Pipeline', 'custom',
]
