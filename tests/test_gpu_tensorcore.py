"""tcgen05 / TMEM path: the self-test GEMM that pins operand layout, descriptors and TMEM read-back."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mn", [0, 2])
@pytest.mark.parametrize("N,K", [(32, 32), (128, 32), (256, 64), (32, 128), (64, 16)])
def test_umma_selftest_gemm(N, K, mn):
    import _kc
    torch.manual_seed(N * 1000 + K)
    A = torch.randn(128, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = _kc.lib().kc_umma_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), N, K, mn,
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _kc.check(rc, "kc_umma_selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    err = (D.double() - ref).abs().max().item()
    assert np.isfinite(err) and err < (2e-2 if mn == 0 else 8e-2) * max(1.0, K ** 0.5), err  # tf32 / bf16 operands
    # exact check with tf32-representable operands: products and sums are then exact in fp32
    A2 = torch.randint(-8, 9, (128, K), device="cuda").float()
    B2 = torch.randint(-8, 9, (N, K), device="cuda").float()
    rc = _kc.lib().kc_umma_selftest(C.c_void_p(A2.data_ptr()), C.c_void_p(B2.data_ptr()), C.c_void_p(D.data_ptr()), N, K, mn,
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _kc.check(rc, "kc_umma_selftest")
    torch.cuda.synchronize()
    assert torch.equal(D, A2 @ B2.T)
