"""tcgen05/TMA GEMM and CUDA-core GEMM vs numpy on bf16-rounded operands (through the C ABI)."""
import numpy as np
import pytest
import torch

from b200asr.engine import test_gemm as run_gemm

pytestmark = pytest.mark.gpu


def _bf16(x):
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()


def _gelu(x):
    return torch.nn.functional.gelu(torch.from_numpy(x)).numpy()


SHAPES = [
    (128, 256, 64), (128, 64, 128), (256, 512, 320), (400, 1280, 1280), (130, 70, 72), (1, 8, 8),
    (800, 3840, 256), (97, 1000, 1288), (400, 256, 5120),
]


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"{m}x{n}x{k}" for m, n, k in SHAPES])
def test_gemm_plain(impl, shape):
    M, N, K = shape
    rng = np.random.default_rng(M * 131 + N * 7 + K)
    A = _bf16(rng.standard_normal((M, K), dtype=np.float32))
    B = _bf16(rng.standard_normal((N, K), dtype=np.float32))
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    out = run_gemm(M, N, K, A, B, impl=impl)
    # fp32 accumulation of exactly-representable bf16 products (|sum| ~ sqrt(K)): error ~ K^0.5 * eps32 * |sum|
    assert np.max(np.abs(out - ref)) <= 5e-3 * max(1.0, np.sqrt(K / 64.0))


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_gemm_epilogue(impl):
    M, N, K = 300, 520, 384
    rng = np.random.default_rng(5)
    A = _bf16(rng.standard_normal((M, K), dtype=np.float32) * 0.2)
    B = _bf16(rng.standard_normal((N, K), dtype=np.float32) * 0.2)
    bias = rng.standard_normal(N, dtype=np.float32)
    res = rng.standard_normal((M, N), dtype=np.float32)
    ref = _gelu((A.astype(np.float64) @ B.astype(np.float64).T + bias).astype(np.float32)) + res
    out = run_gemm(M, N, K, A, B, bias=bias, residual=res, act=1, impl=impl)
    np.testing.assert_allclose(out, ref, atol=2e-4, rtol=1e-4)


def test_gemm_tc_matches_simt_large():
    M, N, K = 1600, 2560, 1280
    rng = np.random.default_rng(9)
    A = _bf16(rng.standard_normal((M, K), dtype=np.float32))
    B = _bf16(rng.standard_normal((N, K), dtype=np.float32))
    a = run_gemm(M, N, K, A, B, impl="tc")
    b = run_gemm(M, N, K, A, B, impl="simt")
    np.testing.assert_allclose(a, b, atol=2e-3, rtol=1e-4)
