"""tcgen05 GEMM core vs a plain fp32 matmul of the same fp16 operands (through the C ABI test hooks)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _native():
    from pigan_b200 import native_test as native
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
    native.check(native.lib.pigan_debug_gemm_nt(a.data_ptr(), b.data_ptr(), None, c.data_ptr(), kd, m, n, splits,
                                                wrap, 0, 0, -1, None, native.current_stream()))
    torch.cuda.synchronize()
    bb = b if not wrap else b.repeat(kd // wrap, 1)
    ref = a.float().t() @ bb.float()
    assert _rel(c, ref) < 1e-5


def test_gemm_nt_tail_and_bias_column():
    """dW of D's first layer: dY^T [spectrum | params | 1 1]: fake rows take their last 64 columns from the tail
    tensor, output clipped to 254 columns, the ones column yields the bias gradient."""
    native = _native()
    torch.manual_seed(99)
    half_rows, m, n = 4096, 512, 256
    kd = 2 * half_rows
    a = (torch.randn(kd, m, device="cuda") * 0.1).half()
    b = (torch.randn(half_rows, n, device="cuda") * 0.1).half()
    b[:, 254:] = 1.0
    tail = b[:, 192:].clone()
    tail[:, 58:62] = (torch.randn(half_rows, 4, device="cuda") * 0.1).half()
    c = torch.zeros(m, 254, device="cuda")
    db = torch.zeros(m, device="cuda")
    native.check(native.lib.pigan_debug_gemm_nt(a.data_ptr(), b.data_ptr(), tail.data_ptr(), c.data_ptr(), kd, m, n,
                                                19, half_rows, half_rows, 254, 254, db.data_ptr(),
                                                native.current_stream()))
    torch.cuda.synchronize()
    b_fake = b.clone()
    b_fake[:, 192:] = tail
    full = a.float().t() @ torch.cat([b, b_fake]).float()
    assert _rel(c, full[:, :254]) < 1e-5
    assert _rel(db, full[:, 254]) < 1e-5


@pytest.mark.parametrize("m,n,k,bias,leaky,rs,tail", [
    (128, 256, 64, False, False, False, False),
    (4096, 512, 256, True, False, False, True),    # first layers: last k-block from the tail tensor
    (4096, 1024, 512, True, False, True, False),   # LayerNorm row partials over 4 tiles
    (65536, 512, 256, True, True, False, False),
    (1000, 256, 512, True, False, True, False),    # ragged rows: TMA store clipping
    (4096, 250, 256, False, False, False, False),  # ragged columns
])
def test_linear_epilogue(m, n, k, bias, leaky, rs, tail):
    native = _native()
    torch.manual_seed(77 + m + n + k)
    a = (torch.randn(m, k, device="cuda") * 0.5).half()
    ld = (n + 7) // 8 * 8
    b = torch.zeros(ld, k, device="cuda", dtype=torch.float16)      # weight rows padded to the stored width
    b[:n] = (torch.randn(n, k, device="cuda") * 0.1).half()
    nt = (n + 255) // 256
    bias_t = None
    if bias:
        bias_t = torch.zeros(nt * 256, device="cuda")
        bias_t[:n] = torch.randn(n, device="cuda")
    a_tail = (torch.randn(m, 64, device="cuda") * 0.5).half() if tail else None
    out = torch.full((m, ld), float("nan"), device="cuda", dtype=torch.float16)
    rowst = torch.zeros(m, nt, 2, device="cuda") if rs else None
    native.check(native.lib.pigan_debug_linear(a.data_ptr(), native.ptr(a_tail), b.data_ptr(), native.ptr(bias_t),
                                               out.data_ptr(), native.ptr(rowst), m, ld, k, int(leaky),
                                               native.current_stream()))
    torch.cuda.synchronize()
    a_eff = a.clone()
    if tail:
        a_eff[:, -64:] = a_tail
    ref = a_eff.float() @ b[:n].float().t()
    if bias:
        ref = ref + bias_t[:n]
    if leaky:
        ref = torch.nn.functional.leaky_relu(ref, 0.2)
    got = out[:, :n].float()
    assert torch.isfinite(got).all()
    assert _rel(got, ref) < 6e-4        # one fp16 rounding of the output
    if rs:
        v = torch.zeros(m, nt * 256, device="cuda")
        v[:, :n] = ref
        v = v.view(m, nt, 256)
        assert _rel(rowst[..., 0], v.sum(2)) < 1e-4
        assert _rel(rowst[..., 1], (v * v).sum(2)) < 1e-5


@pytest.mark.parametrize("m,n,k", [(256, 256, 64), (1000, 512, 256), (4096 + 128, 256, 512), (8192, 1024, 512)])
def test_two_cta_linear_matches_one_cta(m, n, k):
    """cta_group::2 kernel (csrc/gemm2_tc.cuh): a CTA pair shares the B tile; same epilogue, same numbers."""
    native = _native()
    torch.manual_seed(5 + m + n + k)
    a = (torch.randn(m, k, device="cuda") * 0.5).half()
    b = (torch.randn(n, k, device="cuda") * 0.1).half()
    bias = torch.randn((n + 255) // 256 * 256, device="cuda")
    outs = []
    for fn in (native.lib.pigan_debug_linear, native.lib.pigan_debug_linear2):
        out = torch.full((m, n), float("nan"), device="cuda", dtype=torch.float16)
        native.check(fn(a.data_ptr(), None, b.data_ptr(), bias.data_ptr(), out.data_ptr(), None, m, n, k, 1,
                        native.current_stream()))
        torch.cuda.synchronize()
        outs.append(out)
    ref = torch.nn.functional.leaky_relu(a.float() @ b.float().t() + bias[:n], 0.2)
    assert torch.isfinite(outs[1].float()).all()
    assert _rel(outs[1].float(), ref) < 6e-4
    assert torch.equal(outs[0], outs[1])      # same accumulation order per output element
