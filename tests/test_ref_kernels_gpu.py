"""GPU parity against the REFERENCE'S OWN CUDA kernels (oracle/_ref, compiled from /root/reference by
oracle/build_ref.py).  Bit-exact: same arithmetic, same FMA chain order."""
import importlib.util
import os

import pytest
import torch

from oracle import stylegan2_oracle as so
from synthesis_in_style_b200.op import fused_bias_act, upfirdn2d_op

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')


def load_ref(name):
    path = os.path.join(REF_DIR, f'{name}.so')
    if not os.path.exists(path):
        pytest.skip(f'{path} not built (oracle/build_ref.py needs /root/reference)')
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('shape', [(32, 512), (2, 128, 64, 64), (3, 5, 7, 9)])
@pytest.mark.parametrize('code', [(3, 0), (3, 1), (1, 0)])
def test_fused_bias_act_equals_reference_kernel(cuda_device, shape, code):
    ref = load_ref('ref_fused')
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g).to(cuda_device)
    b = torch.randn(shape[1], generator=g).to(cuda_device)
    r = torch.randn(*shape, generator=g).to(cuda_device) if code[1] == 1 else x.new_empty(0)
    want = ref.fused_bias_act(x, b, r, code[0], code[1], 0.2, 2 ** 0.5)
    got = fused_bias_act(x, b, r, code[0], code[1], 0.2, 2 ** 0.5)
    assert torch.equal(got, want)


@pytest.mark.parametrize('cfg', [
    # (major, H, W, up, down, pad, k)  the reference's six modes
    (64, 65, 65, 1, 1, (1, 1), 4), (7, 33, 20, 1, 1, (1, 1), 3), (6, 32, 32, 2, 1, (2, 1), 4),
    (6, 16, 16, 2, 1, (1, 0), 2), (5, 64, 64, 1, 2, (1, 1), 4), (5, 32, 32, 1, 2, (0, 0), 2),
    # planes of <= 32 x 32 outputs: the plane-group kernel (several planes per block, ragged last group, cropping pads)
    (300, 8, 8, 1, 1, (1, 1), 4), (64, 33, 33, 1, 1, (1, 1), 4), (1301, 4, 4, 2, 1, (2, 1), 4), (50, 8, 8, 2, 1, (2, 1), 4),
    (40, 16, 16, 1, 2, (1, 1), 4), (9, 20, 12, 1, 1, (-1, 2), 3), (33, 9, 9, 1, 1, (2, 2), 4)])
def test_upfirdn2d_equals_reference_kernel(cuda_device, cfg):
    ref = load_ref('ref_upfirdn2d')
    major, h, w, up, down, pad, ks = cfg
    g = torch.Generator().manual_seed(2)
    x = torch.randn(major, h, w, 1, generator=g).to(cuda_device)
    k = (torch.randn(ks, ks, generator=g) if ks != 4 else so.make_kernel([1, 3, 3, 1]) * 4).to(cuda_device)
    want = ref.upfirdn2d(x, k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    got = upfirdn2d_op(x, k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    assert got.shape == want.shape and torch.equal(got, want)
