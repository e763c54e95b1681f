"""GPU: the tcgen05 split-bf16 GEMM building block against an fp64 torch reference.

Tolerance: split-bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulate) carries ~16 mantissa bits per operand, so the
result must sit within 3e-5 of the fp64 product relative to max|C| (plain bf16 would be ~4e-3, fp32 ~1e-6)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, B):
    from neuralnj_b200 import _lib
    import __graft_entry__ as g
    g.build()
    L = _lib.lib()
    Z, M, K = A.shape
    N = B.shape[1]
    out = torch.full((Z, M, N), float("nan"), device="cuda")
    ws = torch.empty(4 * (A.numel() + B.numel()) + 4096, dtype=torch.uint8, device="cuda")
    rc = L.nnj_gemm_split_bf16(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(out.data_ptr()), Z, M, N, K,
                               C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "nnj_gemm_split_bf16")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("Z,M,N,K", [(1, 128, 128, 64), (2, 256, 128, 128), (3, 1024, 1024, 400), (1, 200, 72, 1000), (2, 384, 400, 1024),
                                     (1, 256, 256, 64), (5, 512, 768, 200), (150, 256, 256, 128)])   # the last three: CTA-pair kernel (k_tc_gemm2), 1 / 30 / 150 tile pairs
def test_split_bf16_gemm_matches_fp64(Z, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(Z * 1000 + M + N + K)
    A = torch.randn(Z, M, K, device="cuda", generator=g)
    B = torch.randn(Z, N, K, device="cuda", generator=g) * 0.5 + 0.1
    got = _gemm(A, B)
    want = torch.einsum("zmk,znk->zmn", A.double(), B.double())
    assert bool(torch.isfinite(got).all())
    err = float((got.double() - want).abs().max() / want.abs().max())
    assert err < 3e-5, err
    # and it must be far better than single-pass bf16
    bf = torch.einsum("zmk,znk->zmn", A.bfloat16().double(), B.bfloat16().double())
    assert err < 0.05 * float((bf - want).abs().max() / want.abs().max())


def test_split_bf16_gemm_rejects_bad_k():
    from neuralnj_b200 import NnjError
    with pytest.raises(NnjError):
        _gemm(torch.randn(1, 128, 12, device="cuda"), torch.randn(1, 128, 12, device="cuda"))


@pytest.mark.parametrize("N", [64, 128])
def test_thread_written_operands_and_mn_major_b(N):
    """A K-major and B MN-major tiles staged into swizzled shared memory by threads (fused-kernel operand path)."""
    from neuralnj_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(128, 64, device="cuda", generator=g)
    B = torch.randn(64, N, device="cuda", generator=g)
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(L.nnj_tc_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), N,
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    want = A.double() @ B.double()
    assert float((D.double() - want).abs().max() / want.abs().max()) < 3e-5


def test_a_operand_from_tensor_memory():
    """A written to TMEM by tcgen05.st (row = lane, two bf16 per column), W K-major in shared memory: D = A W^T."""
    from neuralnj_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(12)
    A = torch.randn(128, 64, device="cuda", generator=g)
    W = torch.randn(64, 64, device="cuda", generator=g)
    D = torch.full((128, 64), float("nan"), device="cuda")
    _lib.check(L.nnj_tc_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(D.data_ptr()), 1064,
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    want = A.double() @ W.double().T
    assert float((D.double() - want).abs().max() / want.abs().max()) < 3e-5
