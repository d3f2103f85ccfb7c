"""Unit test of the tcgen05 building blocks (csrc/umma.cuh) through the debug GEMM entry:
3xTF32 split-precision D = A @ W on one CTA vs an fp64 matmul."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(32, 16), (64, 256), (256, 256), (256, 16), (256, 48), (512, 8)])
def test_umma_gemm_3xtf32(K, N):
    import torch

    from e_alphazero_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N)
    A = torch.randn(128, K, device="cuda", generator=g) * torch.rand(128, 1, device="cuda", generator=g) * 3
    A[:, ::7] = 0.0  # relu-like zeros
    W = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
    D = torch.full((128, N), float("nan"), device="cuda")
    npad = (N + 15) // 16 * 16
    scratch = torch.empty(K * 2 * npad, dtype=torch.float32, device="cuda")
    rc = lib.eaz_debug_umma_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(D.data_ptr()), K, N,
                                 C.c_void_p(scratch.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.eaz_last_error()
    torch.cuda.synchronize()
    ref = (A.double() @ W.double()).cpu().numpy()
    got = D.cpu().numpy()
    scale = np.abs(A.cpu().numpy().astype(np.float64)) @ np.abs(W.cpu().numpy().astype(np.float64))  # magnitude of the summands
    err = np.abs(got - ref) / np.maximum(scale, 1e-30)
    assert np.isfinite(got).all()
    assert err.max() < 2e-6, (err.max(), np.abs(got - ref).max())
    # plain fp32 accumulation error for comparison is ~1e-7 * sqrt(K); 1xTF32 would be ~5e-4
