"""Forward time of the convolutional evaluators (csrc/convnet.cu, fp32 EXACT mode) -- a measurement aid.
Usage (on a B200): python profiles/bench_convnet.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from e_alphazero_b200 import _abi, ops
from tests import helpers as H

for name, kind, (Hh, W, Cc, A), B, kw, mode in (("resnet 8x8x2, 64 ch x 5 blocks, EXACT", _abi.CONVNET_RESNET, (8, 8, 2, 65), 4096, {}, 0),
                                                ("resnet 8x8x2, 64 ch x 5 blocks, TENSOR", _abi.CONVNET_RESNET, (8, 8, 2, 65), 4096, {}, 1),
                                                ("resnet 19x19x16, 64 ch x 5 blocks, EXACT", _abi.CONVNET_RESNET, (19, 19, 16, 362), 512, {}, 0),
                                                ("resnet 19x19x16, 64 ch x 5 blocks, TENSOR", _abi.CONVNET_RESNET, (19, 19, 16, 362), 512, {}, 1),
                                                ("minatar 10x10x4, EXACT", _abi.CONVNET_MINATAR, (10, 10, 4, 6), 4096, {}, 0)):
    desc = dict(H.random_convnet(kind, Hh, W, Cc, A, seed=1, **kw), mlp_mode=mode)
    net = ops.ConvNetParams(desc)
    obs = torch.as_tensor((np.random.default_rng(2).random((B, Hh, W, Cc)) < 0.3).astype(np.uint8)).cuda()
    for _ in range(3):
        net.forward(obs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        net.forward(obs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    C_ = desc["num_channels"]
    if kind == _abi.CONVNET_RESNET:
        flops = 2 * B * Hh * W * 9 * (Cc * C_ + 10 * C_ * C_)
    else:
        flops = 2 * B * (2 * Hh * W * 9 * Cc * C_ + 2 * (Hh * W * C_ * 64 + 64 * 64) + 4 * 64 * 64)
    print(f"{name:44s} B={B:5d}  {ms:8.3f} ms / forward  {B / ms * 1e3:10.0f} evaluations/s  {flops / ms / 1e9:7.2f} TFLOP/s fp32")
