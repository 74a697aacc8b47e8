"""tcgen05 GEMM core vs a plain fp32 matmul of the same fp16 operands (through the C ABI test hooks)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _native():
    from pigan_b200 import native
    return native


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# variant -> (BLOCK_N, ACC_TILES) as compiled in csrc/debug_gemm.cu
@pytest.mark.parametrize("variant,m,n,k", [
    (0, 128, 256, 64),        # single tile, single k-block
    (0, 128, 256, 256),       # k loop wraps the 4-stage ring exactly once
    (0, 1024, 512, 512),      # several tiles per CTA, ring wrap, TMEM double buffering
    (0, 65536, 512, 256),     # G.L1 / D.L1 shape at the benchmark batch
    (0, 1000, 250, 200),      # ragged M, N, K (TMA zero fill + masked stores)
    (1, 4096, 512, 1024),     # two accumulators per unit (LayerNorm layers), single-buffered TMEM
    (1, 4096, 1024, 512),
    (2, 4096, 258, 256),      # F output layer: 2 x 144 columns
    (3, 4096, 384, 512),      # 128-wide tiles
])
def test_gemm_tn(variant, m, n, k):
    native = _native()
    torch.manual_seed(1234 + m + n + k)
    a = (torch.randn(m, k, device="cuda") * 0.5).half()
    b = (torch.randn(n, k, device="cuda") * 0.5).half()
    c = torch.full((m, n), float("nan"), device="cuda", dtype=torch.float32)
    native.check(native.lib.pigan_debug_gemm_tn(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, variant,
                                                native.current_stream()))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    assert torch.isfinite(c).all()
    # fp16 products are exact in fp32; only the accumulation order differs
    assert _rel(c, ref) < 2e-6
    assert (c - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("kd,m,n,splits,wrap", [
    (64, 128, 256, 1, 0),
    (512, 128, 256, 1, 0),
    (4096, 256, 512, 8, 0),
    (65536, 512, 256, 37, 0),     # dW1 shape at the benchmark batch
    (131072, 256, 512, 37, 0),    # dW2 of D over real+fake rows
    (8192, 512, 256, 16, 4096),   # B operand rows wrap (real+fake rows share the spectrum tile)
    (1000, 256, 256, 3, 0),       # ragged reduction length
])
def test_gemm_nt(kd, m, n, splits, wrap):
    native = _native()
    torch.manual_seed(4321 + kd + m + n)
    a = (torch.randn(kd, m, device="cuda") * 0.1).half()
    rows_b = wrap if wrap else kd
    b = (torch.randn(rows_b, n, device="cuda") * 0.1).half()
    c = torch.zeros(m, n, device="cuda", dtype=torch.float32)
    native.check(native.lib.pigan_debug_gemm_nt(a.data_ptr(), b.data_ptr(), c.data_ptr(), kd, m, n, splits, wrap,
                                                native.current_stream()))
    torch.cuda.synchronize()
    bb = b if not wrap else b.repeat(kd // wrap, 1)
    ref = a.float().t() @ bb.float()
    assert _rel(c, ref) < 1e-5
