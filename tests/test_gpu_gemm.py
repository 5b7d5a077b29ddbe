"""GPU tests of the tcgen05 GEMM core through the C ABI (dinox_gemm_bf16 / _batched / _splitk):
every operand-layout combination, ragged tiles, CTA-pair super tiles with an odd tile count (a
fully out-of-range tile in the peer CTA), the TMA store / reduce-add / bf16 epilogues with guard
bands around the output, and split-K.  Reference: fp32 matmul of the same bf16 operands (the
contraction itself is exact in fp32 up to summation order: rtol 1e-5 on the relative L2 error)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from dinox_b200 import ops, _ext
    assert _ext.lib().dinox_device_check() == 0, "needs a B200"
    return ops


def _operands(M, N, K, am, bm, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    b = torch.randn(N, K, generator=g).to(torch.bfloat16)
    ref = a.float() @ b.float().t()

    def dev(t, mn):
        t = (t.t().contiguous() if mn else t).to(DEV)
        if t.shape[1] % 8:   # TMA needs 16-byte row pitch: embed in a padded buffer
            p = torch.zeros(t.shape[0], (t.shape[1] + 7) // 8 * 8, dtype=t.dtype, device=DEV)
            p[:, :t.shape[1]] = t
            t = p[:, :t.shape[1]]
        return t
    return dev(a, am), dev(b, bm), ref


@pytest.mark.parametrize("M,N,K,am,bm", [
    (128, 256, 64, 0, 0), (200, 384, 384, 0, 0), (300, 1000, 200, 0, 0), (256, 384, 128, 0, 1),
    (128, 256, 64, 1, 0), (512, 384, 1024, 1, 1), (1000, 384, 520, 1, 1), (384, 200, 4096, 0, 1),
    (640, 4096, 384, 0, 0),      # 5 M tiles: the last CTA pair has a fully out-of-range tile
    (8064, 2048, 384, 0, 0),     # 63 M tiles, several tiles per persistent CTA
    (129, 72, 64, 0, 0),
])
def test_gemm_layouts_and_ragged_tiles(ops, M, N, K, am, bm):
    a, b, ref = _operands(M, N, K, am, bm, seed=M + N + K)
    for _ in range(2):   # twice: staging-buffer races are timing dependent
        out = ops.gemm_bf16(a, b, a_mn_major=bool(am), b_mn_major=bool(bm))
        assert rel(out, ref) < 1e-5
        bad = ((out.cpu() - ref).abs() > 1e-3 * ref.abs().max()).sum().item()
        assert bad == 0, f"{bad} wrong elements"


@pytest.mark.parametrize("M,N,K", [(200, 200, 384), (300, 1000, 200), (129, 72, 64), (640, 520, 128)])
def test_gemm_epilogues_keep_guard_bands(ops, M, N, K):
    a, b, ref = _operands(M, N, K, 0, 0, seed=7)
    ld = (N + 7) // 8 * 8
    buf = torch.full((M + 3, ld + 8), 7.0, device=DEV)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(1)).to(DEV)
    ops.gemm_bf16(a, b, out=buf[:M, :N], alpha=0.5, bias_n=bias)
    assert rel(buf[:M, :N], 0.5 * ref + bias.cpu()) < 1e-5
    assert bool((buf[M:] == 7).all()) and bool((buf[:, N:] == 7).all()), "store wrote outside the matrix"
    # accumulate (TMA reduce-add in L2) with a device-side scale: C += 2 * 0.25 * A B^T
    ops.gemm_bf16(a, b, out=buf[:M, :N], accumulate=True, alpha=2.0, alpha_dev=torch.tensor([0.25], device=DEV))
    assert rel(buf[:M, :N], ref + bias.cpu()) < 1e-5
    assert bool((buf[M:] == 7).all()) and bool((buf[:, N:] == 7).all())
    bb = torch.full((M + 3, ld + 8), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16(a, b, out=bb[:M, :N])
    assert rel(bb[:M, :N], ref) < 3e-3     # bf16 rounding of the output
    assert bool((bb[M:] == 7).all()) and bool((bb[:, N:] == 7).all())
    # bf16 read-modify-write goes through the direct-store epilogue
    ops.gemm_bf16(a, b, out=bb[:M, :N], accumulate=True, alpha=-1.0)
    assert bb[:M, :N].float().abs().max().item() <= 2e-2 * ref.abs().max().item()


def test_gemm_batched_gram_shapes(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 200, 384, generator=g).to(torch.bfloat16)
    ref = torch.bmm(x.float(), x.float().transpose(1, 2))
    xd = x.to(DEV)
    assert rel(ops.gemm_bf16_batched(xd, xd), ref) < 1e-5
    assert rel(ops.gemm_bf16_batched(xd, xd, out_dtype=torch.bfloat16), ref) < 3e-3
    y = torch.randn(5, 200, 200, generator=g).to(torch.bfloat16)
    ref2 = torch.bmm(y.float(), x.float())                    # (B, T, T) @ (B, T, D): B operand MN-major
    assert rel(ops.gemm_bf16_batched(y.to(DEV), xd, b_mn_major=True), ref2) < 1e-5


@pytest.mark.parametrize("M,N,K,S", [(256, 384, 4096, 3), (200, 384, 8000, 7), (384, 384, 8064, None), (640, 384, 2048, 2)])
def test_gemm_split_k(ops, M, N, K, S):
    a, b, ref = _operands(M, N, K, 1, 1, seed=11)
    parts = ops.gemm_bf16_splitk(a, b, a_mn_major=True, b_mn_major=True, splits=S)
    assert parts.shape[1:] == (M, N) and (S is None or parts.shape[0] == S)
    assert rel(parts.sum(0), ref) < 1e-5
    # deterministic: the same launch twice gives bit-identical slabs
    parts2 = ops.gemm_bf16_splitk(a, b, a_mn_major=True, b_mn_major=True, splits=parts.shape[0])
    assert torch.equal(parts, parts2)


def test_gemm_rejects_bad_arguments(ops):
    from dinox_b200 import _ext
    a = torch.zeros(128, 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(_ext.DinoxError):
        ops.gemm_bf16(a, torch.zeros(128, 32, dtype=torch.bfloat16, device=DEV))      # K mismatch
    with pytest.raises(_ext.DinoxError):
        ops.gemm_bf16(a.float(), a.float())                                             # not bf16
    with pytest.raises(_ext.DinoxError):
        ops.gemm_bf16_splitk(a, a, splits=200)                                          # more splits than k-blocks
