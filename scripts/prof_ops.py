"""Tiny driver for ncu captures of the standalone ops (keeps the profiled command short)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from synthesis_in_style_b200.op import upfirdn2d, fused_leaky_relu
dev = torch.device('cuda:0')
k = torch.tensor([1., 3., 3., 1.], device=dev); k = k[None] * k[:, None] / 64 * 4
x = torch.randn(32, 128, 257, 257, device=dev)
s = torch.randn(512, 3, 128, 128, device=dev)
for _ in range(3):
    y = upfirdn2d(x, k, pad=(1, 1)); z = upfirdn2d(s, k, up=2, pad=(2, 1)); w = fused_leaky_relu(y, torch.zeros(128, device=dev))
torch.cuda.synchronize()
print('ok', y.shape, z.shape)
