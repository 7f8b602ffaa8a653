"""GPU parity: the two native ops through the C-ABI against the oracle and the golden vectors (bit-exact where the
arithmetic is order-free, 1e-6 where only the summation order differs from ATen's CPU conv)."""
import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as so
from synthesis_in_style_b200.op import FusedLeakyReLU, fused_bias_act, fused_leaky_relu, upfirdn2d, upfirdn2d_op

pytestmark = pytest.mark.gpu

UPFIRDN_CASES = ['blur_mode1', 'up_mode3', 'down_mode5', 'asym_mode1', 'k3_mode2', 'haar_up_mode4', 'haar_down_mode6',
                 'negpad', 'minor3', 'up3_generic']


def test_fused_leaky_relu_golden_bit_exact(golden, cuda_device):
    x = torch.from_numpy(golden['fused/x']).to(cuda_device)
    b = torch.from_numpy(golden['fused/b']).to(cuda_device)
    y = fused_leaky_relu(x, b)
    assert torch.equal(y.cpu(), torch.from_numpy(golden['fused/y']))


@pytest.mark.parametrize('shape', [(32, 512), (2, 128, 256, 256), (3, 5, 7, 9), (1, 3, 1, 1), (4, 6, 2), (0, 4, 8, 8)])
@pytest.mark.parametrize('code', [(3, 0), (3, 1), (1, 0), (3, 2), (1, 1)])
def test_fused_bias_act_modes_bit_exact(cuda_device, shape, code):
    g = torch.Generator().manual_seed(hash((shape, code)) % 1000)
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    ref = torch.randn(*shape, generator=g)
    act, grad = code
    want = so.fused_bias_act(x, b, ref if grad == 1 else x.new_empty(0), act, grad, 0.2, 2 ** 0.5)
    got = fused_bias_act(x.to(cuda_device), b.to(cuda_device), (ref if grad == 1 else x.new_empty(0)).to(cuda_device),
                         act, grad, 0.2, 2 ** 0.5)
    assert got.shape == want.shape and torch.equal(got.cpu(), want)
    # no-bias form (empty bias tensor), as FusedLeakyReLUFunctionBackward uses it
    want = so.fused_bias_act(x, x.new_empty(0), ref, 3, 1, 0.2, 1.0)
    got = fused_bias_act(x.to(cuda_device), x.new_empty(0).to(cuda_device), ref.to(cuda_device), 3, 1, 0.2, 1.0)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize('dtype', [torch.float16, torch.float64])
def test_fused_bias_act_other_dtypes(cuda_device, dtype):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 5, 4, generator=g).to(dtype)
    b = torch.randn(6, generator=g).to(dtype)
    got = fused_leaky_relu(x.to(cuda_device), b.to(cuda_device)).cpu()
    want = so.fused_leaky_relu(x.double(), b.double())
    tol = 2e-3 if dtype == torch.float16 else 1e-7   # fp64: alpha/scale are float constants cast up, as in the reference
    assert got.dtype == dtype
    torch.testing.assert_close(got.double(), want, rtol=tol, atol=tol)


def test_fused_leaky_relu_module_and_autograd(cuda_device):
    m = FusedLeakyReLU(8).to(cuda_device)
    with torch.no_grad():
        m.bias.normal_()
    x = torch.randn(2, 8, 4, 4, device=cuda_device, requires_grad=True)
    y = m(x)
    y.sum().backward()
    xr = x.detach().cpu().requires_grad_(True)
    br = m.bias.detach().cpu().requires_grad_(True)
    yr = torch.nn.functional.leaky_relu(xr + br.view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    yr.sum().backward()
    assert torch.equal(y.detach().cpu(), yr.detach())
    torch.testing.assert_close(x.grad.cpu(), xr.grad)
    torch.testing.assert_close(m.bias.grad.cpu(), br.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('name', UPFIRDN_CASES)
def test_upfirdn2d_golden(golden, cuda_device, name):
    x = torch.from_numpy(golden[f'upfirdn2d/{name}/x']).to(cuda_device)
    k = torch.from_numpy(golden[f'upfirdn2d/{name}/k']).to(cuda_device)
    args = [int(v) for v in golden[f'upfirdn2d/{name}/args']]
    y = upfirdn2d_op(x, k, *args).cpu()
    want = torch.from_numpy(golden[f'upfirdn2d/{name}/y'])
    assert y.shape == want.shape
    torch.testing.assert_close(y, want, rtol=0, atol=2e-6 * float(want.abs().max() + 1))


@pytest.mark.parametrize('cfg', [
    # (B, C, H, W, up, down, pad, ksize)   hot-path uses: Blur after up-conv, ToRGB skip upsample
    (2, 16, 129, 129, 1, 1, (1, 1), 4), (2, 3, 64, 64, 2, 1, (2, 1), 4), (1, 8, 65, 33, 1, 2, (1, 1), 4),
    (3, 5, 31, 47, 1, 1, (1, 1), 3), (2, 4, 16, 16, 2, 1, (1, 0), 2), (2, 4, 32, 32, 1, 2, (0, 0), 2),
    (1, 2, 20, 20, 3, 2, (4, 2), 6), (1, 1, 5, 5, 1, 1, (0, 0), 1),
])
def test_upfirdn2d_against_oracle(cuda_device, cfg):
    b, c, h, w, up, down, pad, ks = cfg
    g = torch.Generator().manual_seed(sum(cfg[:6]))
    x = torch.randn(b, c, h, w, generator=g)
    k = torch.randn(ks, ks, generator=g)
    want = so.upfirdn2d(x, k, up=up, down=down, pad=pad)
    got = upfirdn2d(x.to(cuda_device), k.to(cuda_device), up=up, down=down, pad=pad).cpu()
    assert got.shape == want.shape
    torch.testing.assert_close(got, want, rtol=0, atol=4e-6 * float(want.abs().max() + 1))
    # the kernel follows the reference CUDA kernel's index math exactly (float64 emulation)
    if x.numel() <= 20000:
        emu = so.upfirdn2d_index_emulation(x.reshape(-1, h, w, 1).numpy(), k.numpy(), up, up, down, down, pad[0], pad[1], pad[0], pad[1])
        np.testing.assert_allclose(got.reshape(emu.shape).numpy(), emu, atol=1e-5 * float(want.abs().max() + 1))


def test_upfirdn2d_empty_and_errors(cuda_device):
    k = torch.ones(4, 4, device=cuda_device)
    out = upfirdn2d(torch.zeros(0, 3, 8, 8, device=cuda_device), k, pad=(1, 1))
    assert out.shape == (0, 3, 7, 7)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        upfirdn2d(torch.zeros(1, 1, 8, 8), k)
    with pytest.raises(RuntimeError, match='kernel must be between'):
        upfirdn2d(torch.zeros(1, 1, 64, 64, device=cuda_device), torch.ones(33, 33, device=cuda_device))


def test_upfirdn2d_autograd(cuda_device):
    x = torch.randn(2, 3, 9, 9, device=cuda_device, requires_grad=True)
    k = so.make_kernel([1, 3, 3, 1]).to(cuda_device) * 4
    y = upfirdn2d(x, k, up=2, down=1, pad=(2, 1))
    y.square().sum().backward()
    xr = x.detach().cpu().requires_grad_(True)
    yr = so.upfirdn2d(xr, k.cpu(), up=2, down=1, pad=(2, 1))
    yr.square().sum().backward()
    torch.testing.assert_close(x.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-4)


def test_ops_first_and_second_order_gradients_fp64(cuda_device):
    """Both drop-in ops are differentiable to any order through one self-adjoint Function each: finite-difference
    checks of the first and second derivatives in float64 (the kernels dispatch on dtype like the reference's)."""
    from torch.autograd import gradcheck, gradgradcheck
    torch.manual_seed(0)
    x = torch.randn(2, 3, 5, 5, device=cuda_device, dtype=torch.float64)
    x = (x + 0.3 * x.sign()).requires_grad_(True)          # keep away from the kink at 0
    b = torch.randn(3, device=cuda_device, dtype=torch.float64, requires_grad=True)
    with torch.no_grad():
        shifted = x + b.view(1, -1, 1, 1)
        x.sub_(torch.where(shifted.abs() < 0.1, shifted, torch.zeros_like(shifted)) * 2)
    assert gradcheck(lambda a, c: fused_leaky_relu(a, c), (x, b), eps=1e-6, atol=1e-6)
    assert gradgradcheck(lambda a, c: fused_leaky_relu(a, c) ** 2, (x, b), eps=1e-6, atol=1e-5)
    k = (so.make_kernel([1, 3, 3, 1]) * 4).to(cuda_device, torch.float64)
    for up, down, pad in ((2, 1, (2, 1)), (1, 2, (1, 1)), (1, 1, (1, 1)), (2, 3, (0, 2))):
        xi = torch.randn(1, 2, 7, 6, device=cuda_device, dtype=torch.float64, requires_grad=True)
        # like the reference's kernel (upfirdn2d_kernel.cu:56-57 stages input and taps in `float` shared memory) the fp64
        # path has fp32 products: the op is linear (its square quadratic), so central differences are exact for any step
        # and a large step keeps the fp32 rounding out of the quotient
        assert gradcheck(lambda a: upfirdn2d(a, k, up=up, down=down, pad=pad), (xi,), eps=1e-1, atol=1e-4, rtol=1e-3), (up, down, pad)
        assert gradgradcheck(lambda a: upfirdn2d(a, k, up=up, down=down, pad=pad) ** 2, (xi,), eps=1e-1, atol=1e-3, rtol=1e-3), (up, down, pad)
