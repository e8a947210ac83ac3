"""Direct kernels for the 1/3-channel ends of the dcgan stacks and BatchNorm backward without the activation tensor."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False  # the torch reference convolutions must be exact fp32
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def K():
    from p2pvg_b200._lib import CudaKernels
    return CudaKernels("cuda")


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,H,Ci,Co", [(3, 64, 1, 64), (2, 32, 3, 64), (5, 8, 1, 128)])
def test_conv_thin_in(K, dtype, N, H, Ci, Co):
    torch.manual_seed(0)
    x = torch.randn(N, Ci, H, H, device="cuda").to(dtype)
    w = torch.randn(Co, Ci, 4, 4, device="cuda") * 0.1
    b = torch.randn(Co, device="cuda")
    ref = nhwc(F.conv2d(x.float(), w, b, stride=2, padding=1))
    y = torch.empty(N, H // 2, H // 2, Co, device="cuda", dtype=dtype)
    K.conv_thin_in(nhwc(x), w, b, y, N, H, H, Ci, Co)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert (y.float() - ref).abs().max().item() <= tol * ref.abs().max().item() + tol
    K.conv_thin_in(nhwc(x), w, None, y, N, H, H, Ci, Co)
    ref0 = nhwc(F.conv2d(x.float(), w, None, stride=2, padding=1))
    assert (y.float() - ref0).abs().max().item() <= tol * ref.abs().max().item() + tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,H,Ci,Co", [(4, 32, 64, 1), (2, 16, 64, 3), (6, 8, 128, 1)])
def test_convT_thin_out(K, dtype, N, H, Ci, Co):
    torch.manual_seed(1)
    x = torch.randn(N, Ci, H, H, device="cuda").to(dtype)
    w = torch.randn(Ci, Co, 4, 4, device="cuda") * 0.1
    b = torch.randn(Co, device="cuda")
    ref = nhwc(F.conv_transpose2d(x.float(), w, b, stride=2, padding=1))
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    y32 = torch.empty(N, 2 * H, 2 * H, Co, device="cuda")
    K.convT_thin_out(nhwc(x), w, b, y32, N, H, H, Ci, Co)
    assert (y32 - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-4
    B = 2
    addend = torch.randn(2 * B, 2 * H, 2 * H, Co, device="cuda")
    src = torch.tensor([g % 2 for g in range(N // B)], dtype=torch.int32, device="cuda")
    idx = torch.tensor([(g % 2) * B + i for g in range(N // B) for i in range(B)], device="cuda")
    y = torch.empty(N, 2 * H, 2 * H, Co, device="cuda", dtype=dtype)
    K.convT_thin_out(nhwc(x), w, None, y, N, H, H, Ci, Co, addend=addend, grp_src=src, imgs_per_group=B)
    want = nhwc(F.conv_transpose2d(x.float(), w, None, stride=2, padding=1)) + addend[idx]
    assert (y.float() - want).abs().max().item() <= tol * want.abs().max().item() + tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_bwd_without_activation_tensor(K, dtype):
    torch.manual_seed(2)
    G, R, C = 3, 200, 64
    x = (torch.randn(G, R, C, device="cuda") * 2 + 0.3).to(dtype)
    gamma, beta = torch.randn(C, device="cuda") * 0.1 + 1, torch.randn(C, device="cuda") * 0.1
    st = [torch.zeros(G * C, device="cuda") for _ in range(5)]
    K.bn_fwd_stats(x, G, R, C, gamma, beta, *st)
    y = torch.empty_like(x)
    K.bn_act(x, y, st[3], st[4], G, R, C, 1)
    dy = torch.randn(G, R, C, device="cuda").to(dtype)
    outs = []
    for use_y in (True, False):
        dx = torch.empty_like(x)
        s0, s1 = torch.zeros(G * C, device="cuda"), torch.zeros(G * C, device="cuda")
        if use_y:
            K.bn_bwd(dy, x, y, st[0], st[1], gamma, G, R, C, 1, dx, s0, s1)
        else:
            K.bn_bwd(dy, x, None, st[0], st[1], gamma, G, R, C, 1, dx, s0, s1, scale=st[3], shift=st[4])
        outs.append((dx, s0, s1))
    for a, b in zip(*outs):
        assert torch.equal(a, b) or (a.float() - b.float()).abs().max().item() <= 1e-6
