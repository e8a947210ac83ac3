"""TF32 tensor-core path for fp32 operands and split-K of the CUDA-core GEMM (C ABI: p2pvg_gemm with
flags = P2PVG_GEMM_TF32), against the torch emulation."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def KS():
    from p2pvg_b200._lib import CudaKernels
    from tests.emu_backend import EmuKernels
    return CudaKernels("cuda"), EmuKernels("cuda")


@pytest.mark.parametrize("M,N,K", [(256, 1024, 256), (7424, 256, 264), (100, 10, 256), (256, 128, 256), (300, 1024, 40), (64, 64, 32)])
def test_tf32_gemm(KS, M, N, K):
    Kc, Ke = KS
    torch.manual_seed(3)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    add = torch.randn(M, N, device="cuda")
    C1 = torch.randn(M, N, device="cuda")
    C2 = C1.clone()
    Kc.set_fp32_gemm_mode(1)
    try:
        Kc.gemm(A, B, C1, M, N, K, bias=bias, addend=add, accumulate=True)
    finally:
        Kc.set_fp32_gemm_mode(0)
    Ke.gemm(A, B, C2, M, N, K, bias=bias, addend=add, accumulate=True)
    err = (C1 - C2).abs().max().item()
    assert err <= 1e-2 * K ** 0.5, f"tf32 error {err}"          # 10-bit (truncated) mantissa operands, fp32 accumulation; max over M*N outputs
    assert err > 0 or K < 8                                        # really ran at reduced precision
    C3 = C1.clone()
    Kc.gemm(A, B, C3, M, N, K, bias=bias, addend=add)            # exact mode again
    Ke.gemm(A, B, C2, M, N, K, bias=bias, addend=add)
    assert (C3 - C2).abs().max().item() <= 1e-4 * K ** 0.5


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(10, 256, 7424, True, True), (128, 64, 5000, True, True), (256, 258, 7424, True, True),
                                             (20, 30, 100000, False, False)])
def test_simt_splitk(KS, M, N, K, a_mn, b_mn):
    Kc, Ke = KS
    torch.manual_seed(4)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.1
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.1
    bias = torch.randn(N, device="cuda")
    C1 = torch.randn(M, N, device="cuda")
    C2 = C1.clone()
    Kc.gemm(A, B, C1, M, N, K, a_mn=a_mn, b_mn=b_mn, bias=bias, accumulate=True)
    Ke.gemm(A, B, C2, M, N, K, a_mn=a_mn, b_mn=b_mn, bias=bias, accumulate=True)
    assert torch.allclose(C1, C2, rtol=1e-3, atol=1e-3)
