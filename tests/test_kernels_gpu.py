"""GPU parity tests of the individual kernels, called through the C ABI (ctypes).
Checkers are plain fp32 torch math on the same seeded inputs (floating-point kernels) and the C oracle
(bit-exact integer output of the VQ search)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from open_muse_b200 import ops  # noqa: E402

DEV = "cuda"


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


GEMM_SHAPES = [
    (128, 128, 64), (128, 256, 128), (256, 512, 512), (304, 200, 136), (1032, 1536, 512), (64, 72, 64),
    (520, 2048, 512), (2056, 512, 2048),
]


@pytest.mark.parametrize("majors", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("shape", GEMM_SHAPES)
def test_gemm_bf16_out(majors, shape):
    M, N, K = shape
    a_mn, b_mn = majors
    A = _rand((K, M) if a_mn else (M, K), 1)
    B = _rand((K, N) if b_mn else (N, K), 2)
    ref = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t())
    ldc = ((N + 7) // 8) * 8
    C = torch.full((M, ldc), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, B, C, M, N, K, A.stride(0), B.stride(0), ldc, a_mn, b_mn, ops.EPI_BF16)
    torch.cuda.synchronize()
    assert _rel(C[:, :N], ref) < 6e-3  # bf16 output rounding (2^-9 relative per element)
    if ldc > N:
        assert bool((C[:, N:] == 7.0).all())  # pad columns untouched
    # independent on-device cross-check: the first-generation mma.sync kernel (tests/xcheck, not in the product library)
    from tests import xcheck

    C2 = torch.full((M, ldc), 7.0, dtype=torch.bfloat16, device=DEV)
    xcheck.gemm(A, B, C2, M, N, K, A.stride(0), B.stride(0), ldc, a_mn, b_mn, ops.EPI_BF16)
    assert _rel(C[:, :N], C2[:, :N]) < 3e-3  # same fp32 products, different accumulation order, bf16 output rounding


def test_gemm_epilogues():
    M, N, K = 384, 320, 192
    A, B = _rand((M, K), 3), _rand((N, K), 4)
    ref = A.float() @ B.float().t()
    C = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(A, B, C, M, N, K, K, K, N, 0, 0, ops.EPI_F32)
    assert _rel(C, ref) < 1e-5
    res = torch.randn(M, N, device=DEV)
    C2 = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(A, B, C2, M, N, K, K, K, N, 0, 0, ops.EPI_RESADD_F32, res=res)
    assert _rel(C2, res + ref.to(torch.bfloat16).float()) < 1e-4  # bf16 rounding of the accumulator may flip
    # split-K atomic accumulate (wgrad shape: short M,N, long K, both operands MN-major)
    T, No, Ki = 4112, 192, 128
    dY, X = _rand((T, No), 5), _rand((T, Ki), 6)
    dW = torch.ones(No, Ki, dtype=torch.float32, device=DEV)
    ops.linear_wgrad(dY, X, dW)
    assert _rel(dW, 1.0 + dY.float().t() @ X.float()) < 1e-4


@pytest.mark.parametrize("H", [128, 512, 2048, 1000])
@pytest.mark.parametrize("rms", [0, 1])
@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_norm_fwd_bwd(H, rms, xdt, ydt):
    rows, eps = 77, 1e-6
    x = _rand((rows, H), 1, dtype=xdt)
    w = (1 + 0.1 * torch.randn(H, generator=torch.Generator().manual_seed(2))).to(DEV)
    res = torch.randn(rows, H, device=DEV) if ydt == torch.float32 else None
    for act in (0, 1):
        y, stats = ops.norm_fwd(x, w, eps, ydt, res=res, act=act, rms=rms)
        xr = x.float().clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        a = torch.nn.functional.gelu(xr) if act else xr
        if rms:
            n = a * torch.rsqrt(a.pow(2).mean(-1, keepdim=True) + eps) * wr
        else:
            n = torch.nn.functional.layer_norm(a, (H,), wr, None, eps)
        yr = n + (res if res is not None else 0)
        tol = 5e-3 if ydt == torch.bfloat16 else 1e-5
        assert _rel(y, yr) < tol
        dy = _rand((rows, H), 3, dtype=torch.bfloat16 if ydt == torch.bfloat16 else torch.float32)
        dres = torch.randn(rows, H, device=DEV)
        dw = torch.zeros(H, device=DEV)
        dx = ops.norm_bwd(dy, x, w, stats, torch.float32, dw=dw, dres=dres, act=act, rms=rms)
        yr.backward(dy.float())
        assert _rel(dx, xr.grad + dres) < 1e-4
        assert _rel(dw, wr.grad) < 1e-4


@pytest.mark.parametrize("B,S,H,V", [(5, 17, 128, 72), (7, 257, 512, 2025), (3, 33, 96, 50), (2, 256, 1024, 8256)])
def test_embed_fwd_bwd(B, S, H, V):
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, V, (B, S), generator=g).to(DEV)
    ids[:, 3] = V - 1  # hot row (mask token): accumulated per CTA before touching global memory
    ids[torch.rand(B, S, generator=g).to(DEV) < 0.5] = V - 1
    word, pos = torch.randn(V, H, generator=g).to(DEV), torch.randn(S + 3, H, generator=g).to(DEV)
    out = ops.embed_fwd(ids, word, pos)
    ref = word[ids] + pos[:S][None]
    assert torch.equal(out.view(B, S, H), ref)
    dx = torch.randn(B * S, H, device=DEV)
    dword, dpos = torch.zeros_like(word), torch.zeros_like(pos)
    ops.embed_bwd(ids, dx, dword, dpos)
    rw = torch.zeros_like(word).index_add_(0, ids.view(-1), dx)
    assert _rel(dword, rw) < 2e-6 and _rel(dword[V - 1], rw[V - 1]) < 2e-6
    assert _rel(dpos[:S], dx.view(B, S, H).sum(0)) < 1e-6 and bool((dpos[S:] == 0).all())


def test_glu_fwd_bwd():
    rows, I = 33, 256
    ab = _rand((rows, 2 * I), 1)
    out = ops.glu_fwd(ab)
    a = ab[:, :I].float().clone().requires_grad_(True)
    b = ab[:, I:].float().clone().requires_grad_(True)
    ref = torch.nn.functional.gelu(a) * b
    assert _rel(out, ref) < 6e-3
    d = _rand((rows, I), 2)
    dab = ops.glu_bwd(ab, d)
    ref.backward(d.float())
    assert _rel(dab[:, :I], a.grad) < 8e-3 and _rel(dab[:, I:], b.grad) < 8e-3


@pytest.mark.parametrize("V,ld,ls", [(2025, 2048, 0.0), (2025, 2048, 0.1), (64, 64, 0.0), (8192, 8192, 0.1)])
def test_cross_entropy(V, ld, ls):
    rows = 203
    g = torch.Generator().manual_seed(0)
    logits = torch.zeros(rows, ld, dtype=torch.bfloat16, device=DEV)
    logits[:, :V] = (torch.randn(rows, V, generator=g) * 2).to(torch.bfloat16).to(DEV)
    labels = torch.randint(0, V, (rows,), generator=g)
    labels[torch.rand(rows, generator=g) < 0.4] = -100
    labels = labels.to(DEV)
    out, ws = ops.ce_fwd(logits, labels, V, ls)
    lr = logits[:, :V].float().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, labels, ignore_index=-100, label_smoothing=ls)
    assert abs(float(out[0]) - float(ref)) < 2e-5 * max(1.0, abs(float(ref)))
    assert int(out[1]) == int((labels != -100).sum())
    dloss = torch.tensor([0.5], device=DEV)
    dl = ops.ce_bwd(logits, labels, ws, dloss, out, V, ls)
    ref.backward(torch.tensor(0.5, device=DEV))
    assert _rel(dl[:, :V], lr.grad) < 6e-3
    assert bool((dl[:, V:] == 0).all())
    assert bool((dl[labels == -100] == 0).all())


def _attn_ref(q, k, v, scale):
    s = (q @ k.transpose(-1, -2)) * scale
    return s.softmax(-1) @ v


@pytest.mark.parametrize("B,nh,Sq,Skv", [(2, 2, 257, 257), (1, 1, 64, 64), (3, 2, 16, 77), (2, 8, 256, 256), (1, 2, 130, 5), (4, 8, 257, 257), (2, 16, 256, 77), (3, 4, 385, 385), (2, 2, 66, 194), (1, 3, 193, 129)])
def test_attention_fwd_bwd(B, nh, Sq, Skv):
    H = nh * 64
    cross = Sq != Skv
    scale = 1.0 / math.sqrt(64)
    if cross:
        qb = _rand((B * Sq, H), 1)
        kvb = _rand((B * Skv, 2 * H), 2)
        q, k, v = qb, kvb[:, :H], kvb[:, H:]
    else:
        qkv = _rand((B * Sq, 3 * H), 1)
        q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    o, lse = ops.attn_fwd(q, k, v, B, nh, Sq, Skv, scale)
    qr = q.float().reshape(B, Sq, nh, 64).transpose(1, 2).clone().requires_grad_(True)
    kr = k.float().reshape(B, Skv, nh, 64).transpose(1, 2).clone().requires_grad_(True)
    vr = v.float().reshape(B, Skv, nh, 64).transpose(1, 2).clone().requires_grad_(True)
    ref = _attn_ref(qr, kr, vr, scale)
    assert _rel(o.view(B, Sq, nh, 64).transpose(1, 2), ref) < 8e-3
    lse_ref = torch.logsumexp((qr @ kr.transpose(-1, -2)) * scale, dim=-1)
    assert _rel(lse, lse_ref) < 1e-4
    do = _rand((B * Sq, H), 3)
    dq = torch.empty_like(q.contiguous()) if cross else None
    if cross:
        dkv = torch.empty_like(kvb)
        ops.attn_bwd(q, k, v, o, do, lse, dq, dkv[:, :H], dkv[:, H:], B, nh, Sq, Skv, scale)
        dk, dv = dkv[:, :H], dkv[:, H:]
    else:
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(q, k, v, o, do, lse, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, nh, Sq, Skv, scale)
        dq, dk, dv = dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:]
    ref.backward(do.float().view(B, Sq, nh, 64).transpose(1, 2))
    assert _rel(dq.reshape(B, Sq, nh, 64).transpose(1, 2), qr.grad) < 1.5e-2
    assert _rel(dk.reshape(B, Skv, nh, 64).transpose(1, 2), kr.grad) < 1.5e-2
    assert _rel(dv.reshape(B, Skv, nh, 64).transpose(1, 2), vr.grad) < 1.5e-2
    # independent on-device cross-check: the first-generation mma.sync attention kernels (tests/xcheck)
    from tests import xcheck

    o2, lse2 = xcheck.attn_fwd(q, k, v, B, nh, Sq, Skv, scale)
    assert _rel(o, o2) < 8e-3 and _rel(lse, lse2) < 1e-4
    dq2, dk2, dv2 = torch.empty_like(dq), torch.empty_like(dk.contiguous()), torch.empty_like(dv.contiguous())
    xcheck.attn_bwd(q, k, v, o2, do, lse2, dq2, dk2, dv2, B, nh, Sq, Skv, scale)
    assert _rel(dq, dq2) < 2e-2 and _rel(dk, dk2) < 2e-2 and _rel(dv, dv2) < 2e-2


@pytest.mark.parametrize("B,nh,Sq,Skv", [(2, 2, 257, 257), (1, 16, 257, 257), (3, 4, 16, 77), (2, 16, 256, 256), (1, 3, 193, 129)])
def test_attention_head_dim_48_fwd_bwd(B, nh, Sq, Skv):
    """head_dim 48 (configs/imagenet.yaml: hidden 768, 16 heads): the TMA boxes stay 64 columns wide (the trailing 16 belong
    to the next head, or to the neighbouring K block of the fused projection, or are zero-filled past the tensor) and must
    never enter a product; the accumulate MMAs run with N = 48.  Forward, LSE and all three gradients vs fp32 torch."""
    hd = 48
    H = nh * hd
    cross = Sq != Skv
    scale = 1.0 / math.sqrt(hd)
    if cross:
        qb = _rand((B * Sq, H), 1)
        kvb = _rand((B * Skv, 2 * H), 2)
        q, k, v = qb, kvb[:, :H], kvb[:, H:]
    else:
        qkv = _rand((B * Sq, 3 * H), 1)
        q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    o, lse = ops.attn_fwd(q, k, v, B, nh, Sq, Skv, scale, head_dim=hd)
    assert o.shape == (B * Sq, H)
    qr = q.float().reshape(B, Sq, nh, hd).transpose(1, 2).clone().requires_grad_(True)
    kr = k.float().reshape(B, Skv, nh, hd).transpose(1, 2).clone().requires_grad_(True)
    vr = v.float().reshape(B, Skv, nh, hd).transpose(1, 2).clone().requires_grad_(True)
    ref = _attn_ref(qr, kr, vr, scale)
    assert _rel(o.view(B, Sq, nh, hd).transpose(1, 2), ref) < 8e-3
    lse_ref = torch.logsumexp((qr @ kr.transpose(-1, -2)) * scale, dim=-1)
    assert _rel(lse, lse_ref) < 1e-4
    do = _rand((B * Sq, H), 3)
    if cross:
        dq = torch.empty_like(q.contiguous())
        dkv = torch.empty_like(kvb)
        ops.attn_bwd(q, k, v, o, do, lse, dq, dkv[:, :H], dkv[:, H:], B, nh, Sq, Skv, scale, head_dim=hd)
        dk, dv = dkv[:, :H], dkv[:, H:]
    else:
        dqkv = torch.full_like(qkv, float("nan"))  # every element of the three gradients must be written
        ops.attn_bwd(q, k, v, o, do, lse, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, nh, Sq, Skv, scale, head_dim=hd)
        dq, dk, dv = dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:]
        assert bool(torch.isfinite(dqkv.float()).all())
    ref.backward(do.float().view(B, Sq, nh, hd).transpose(1, 2))
    assert _rel(dq.reshape(B, Sq, nh, hd).transpose(1, 2), qr.grad) < 1.5e-2
    assert _rel(dk.reshape(B, Skv, nh, hd).transpose(1, 2), kr.grad) < 1.5e-2
    assert _rel(dv.reshape(B, Skv, nh, hd).transpose(1, 2), vr.grad) < 1.5e-2


@pytest.mark.parametrize("n,ncodes,D", [(512, 1024, 256), (300, 128, 64), (1, 64, 16), (4096, 1024, 256)])
def test_vq_argmin_bit_exact_vs_c_oracle(n, ncodes, D):
    from oracle import vq_oracle as VQ

    g = torch.Generator().manual_seed(n)
    z = torch.randn(n, D, generator=g)
    cb = torch.randn(ncodes, D, generator=g) * 0.7
    cb[ncodes // 2] = cb[3]  # exact duplicate code: lowest index must win
    ids, dmin = ops.vq_argmin(z.to(DEV), cb.to(DEV), return_dmin=True)
    ids_o, dmin_o = VQ.argmin(z.numpy(), cb.numpy())
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    assert np.array_equal(dmin.cpu().numpy().view(np.uint32), dmin_o.view(np.uint32))  # distances bit-identical too


def test_vq_golden_ids(golden):
    from oracle import vq_oracle as VQ

    g = golden("vq_quantizer.pt")
    z = torch.from_numpy(VQ.nchw_to_rows(g["z"].numpy())).to(DEV)
    ids = ops.vq_argmin(z, g["codebook"].to(DEV))
    assert torch.equal(ids.cpu().view(2, -1), g["ids"])
    entry = ops.vq_lookup_nchw(g["ids"].to(DEV), g["codebook"].to(DEV))
    assert torch.equal(entry.cpu().view_as(g["entry"]), g["entry"])


def test_vq_soft_code_vs_reference_fixture_and_oracle(golden):
    """get_soft_code (:327-340): soft within fp32 tolerance, deterministic ids exact, stochastic ids exact through
    torch's RNG stream (Exp(1) draws reproduced on the CPU generator)."""
    from oracle import vq_oracle as VQ

    g = golden("vq_soft_code.pt")
    z = torch.from_numpy(VQ.nchw_to_rows(g["z"].numpy())).to(DEV)
    cb = g["codebook"].to(DEV)
    soft, ids = ops.vq_soft_code(z, cb, g["temp"])
    assert torch.equal(ids.cpu().view(2, -1), g["code"])
    torch.testing.assert_close(soft.cpu().view_as(g["soft"]), g["soft"], rtol=2e-4, atol=1e-9)
    torch.testing.assert_close(soft.sum(1).cpu(), torch.ones(soft.shape[0]), rtol=1e-5, atol=1e-5)
    torch.manual_seed(g["seed_s"])
    q = torch.empty(z.shape[0], cb.shape[0]).exponential_()
    soft_s, ids_s = ops.vq_soft_code(z, cb, g["temp_s"], q.to(DEV))
    assert torch.equal(ids_s.cpu().view(2, -1), g["code_s"])
    torch.testing.assert_close(soft_s.cpu().view_as(g["soft_s"]), g["soft_s"], rtol=2e-4, atol=1e-9)
    # full-size shape (1024 codes x 256 dims, ragged row count) against the oracle
    gen = torch.Generator().manual_seed(11)
    z2, cb2 = torch.randn(700, 256, generator=gen), torch.randn(1024, 256, generator=gen) * 0.7
    soft2, ids2 = ops.vq_soft_code(z2.to(DEV), cb2.to(DEV), 100.0)
    soft_o, ids_o = VQ.soft_code(z2.numpy(), cb2.numpy(), 100.0)
    assert np.array_equal(ids2.cpu().numpy(), ids_o)
    np.testing.assert_allclose(soft2.cpu().numpy(), soft_o, rtol=1e-3, atol=1e-8)


@pytest.mark.parametrize("I,rms", [(2048, 0), (128, 0), (512, 1), (4096, 0)])
def test_norm_glu_fused_fwd_bwd(I, rms):
    """act=2: LN(gelu(a) * b) straight from the [a | b] GEMM output; backward emits d[a | b]."""
    rows, eps = 67, 1e-6
    ab = _rand((rows, 2 * I), 1)
    w = (1 + 0.1 * torch.randn(I, generator=torch.Generator().manual_seed(2))).to(DEV)
    y, stats = ops.norm_fwd(ab, w, eps, torch.bfloat16, act=2, rms=rms)
    a = ab[:, :I].float().clone().requires_grad_(True)
    b = ab[:, I:].float().clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    v = torch.nn.functional.gelu(a) * b
    n = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps) * wr if rms else torch.nn.functional.layer_norm(v, (I,), wr, None, eps)
    assert y.shape == (rows, I) and _rel(y, n) < 8e-3
    dy = _rand((rows, I), 3)
    dw = torch.zeros(I, device=DEV)
    dab = ops.norm_bwd(dy, ab, w, stats, torch.bfloat16, dw=dw, act=2, rms=rms)
    n.backward(dy.float())
    assert dab.shape == ab.shape
    assert _rel(dab[:, :I], a.grad) < 1.2e-2 and _rel(dab[:, I:], b.grad) < 1.2e-2
    assert _rel(dw, wr.grad) < 1e-2
    # with the saved forward output the first pass skips the GELU: same result within bf16 noise
    dw2 = torch.zeros(I, device=DEV)
    dab2 = ops.norm_bwd(dy, ab, w, stats, torch.bfloat16, dw=dw2, act=2, rms=rms, y_fwd=y)
    assert _rel(dab2[:, :I], a.grad) < 1.2e-2 and _rel(dab2[:, I:], b.grad) < 1.2e-2
    assert _rel(dw2, wr.grad) < 1e-2 and _rel(dab2, dab.float()) < 4e-3


# ------------------------------------------------------------------------------------------ reproducible reductions
@pytest.mark.parametrize("T,N,K", [(4112, 192, 128), (65792 // 8, 1536, 512), (520, 72, 200), (2056, 2048, 512), (130, 128, 64)])
def test_wgrad_deterministic_splitk(T, N, K):
    """dW = dY^T X through the deterministic split-K epilogue: matches fp32 math, does not need a zeroed output, and two
    runs agree bit for bit (the atomic epilogue does not guarantee that)."""
    dY, X = _rand((T, N), 5), _rand((T, K), 6)
    ref = dY.float().t() @ X.float()
    outs = []
    for _ in range(3):
        dW = torch.full((N, K), float("nan"), dtype=torch.float32, device=DEV)  # every element must be overwritten
        ops.linear_wgrad_det(dY, X, out=dW)
        outs.append(dW)
    torch.cuda.synchronize()
    assert _rel(outs[0], ref) < 1e-5
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("H,act,xdt", [(512, 0, torch.float32), (1024, 1, torch.bfloat16), (2048, 0, torch.bfloat16), (2048, 2, torch.bfloat16)])
def test_norm_bwd_dw_deterministic(H, act, xdt):
    rows = 4099
    x = _rand((rows, 2 * H if act == 2 else H), 1, dtype=xdt)
    w = (1 + 0.1 * torch.randn(H, generator=torch.Generator().manual_seed(2))).to(DEV)
    _, stats = ops.norm_fwd(x, w, 1e-5, torch.bfloat16, act=act)
    dy = _rand((rows, H), 3)
    dw_atomic = torch.zeros(H, device=DEV)
    dx_a = ops.norm_bwd(dy, x, w, stats, torch.bfloat16, dw=dw_atomic, act=act)
    res = [ops.norm_bwd(dy, x, w, stats, torch.bfloat16, act=act, want_dw=True) for _ in range(3)]
    torch.cuda.synchronize()
    assert torch.equal(res[0][0], dx_a)
    assert _rel(res[0][1], dw_atomic) < 1e-5
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][1], res[2][1])


@pytest.mark.parametrize("B,S,H,V", [(5, 17, 128, 72), (64, 257, 512, 2025), (3, 33, 96, 50), (4, 256, 1024, 8256)])
def test_embed_bwd_deterministic(B, S, H, V):
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, V - 10, (B, S), generator=g).to(DEV)  # the last ids stay unused: their rows must come out zero
    ids[torch.rand(B, S, generator=g).to(DEV) < 0.5] = V - 1    # the mask token: one very long segment
    dx = torch.randn(B * S, H, device=DEV)
    outs = [ops.embed_bwd_det(ids, dx, V, S + 3) for _ in range(2)]
    torch.cuda.synchronize()
    rw = torch.zeros(V, H, device=DEV, dtype=torch.float64).index_add_(0, ids.view(-1), dx.double())
    dword, dpos = outs[0]
    assert _rel(dword, rw) < 1e-6 and bool((dword[V - 10:V - 1] == 0).all())
    assert _rel(dpos[:S], dx.view(B, S, H).sum(0)) < 1e-6 and bool((dpos[S:] == 0).all())
    assert torch.equal(dword, outs[1][0]) and torch.equal(dpos, outs[1][1])


def test_norm_bwd_emits_bf16_copy_of_dx():
    rows, H = 1031, 512
    x = _rand((rows, H), 1, dtype=torch.float32)
    w = (1 + 0.1 * torch.randn(H, generator=torch.Generator().manual_seed(2))).to(DEV)
    _, stats = ops.norm_fwd(x, w, 1e-5, torch.bfloat16)
    dy, dres = _rand((rows, H), 3), torch.randn(rows, H, device=DEV)
    dx, dw = ops.norm_bwd(dy, x, w, stats, torch.float32, dres=dres, want_dw=True, bf16_copy=True)
    ref, _ = ops.norm_bwd(dy, x, w, stats, torch.float32, dres=dres, want_dw=True)
    copy = ops.take_bf16_copy(dx)
    assert torch.equal(dx, ref) and copy is not None and torch.equal(copy, dx.to(torch.bfloat16))
    assert ops.take_bf16_copy(dx) is None  # consumed once


@pytest.mark.parametrize("H,rms1", [(512, 0), (128, 1), (1024, 0), (768, 1)])
def test_fused_double_norm_matches_the_two_single_kernels(H, rms1):
    """post-attention norm + residual followed by the FFN pre-norm, forward and backward, fused vs the two single-norm
    kernels (same formulas; the compiler contracts multiply-adds differently, so fp32 values agree to rounding)."""
    rows, eps = 2051, 1e-6
    a = _rand((rows, H), 1)
    x = _rand((rows, H), 2, dtype=torch.float32)
    g = torch.Generator().manual_seed(3)
    w1 = (1 + 0.1 * torch.randn(H, generator=g)).to(DEV)
    w2 = (1 + 0.1 * torch.randn(H, generator=g)).to(DEV)
    x2_ref, st1 = ops.norm_fwd(a, w1, eps, torch.float32, res=x, rms=rms1)
    h2_ref, st2 = ops.norm_fwd(x2_ref, w2, eps, torch.bfloat16, rms=0)
    x2, h2, st = ops.norm2_fwd(a, x, w1, w2, eps, rms1=rms1, rms2=0)
    assert _rel(x2, x2_ref) < 1e-6 and _rel(h2, h2_ref) < 4e-3  # h2 is bf16: rounding flips on fp32 last-bit differences
    assert _rel(st[1], st1[1]) < 1e-6 and _rel(st[2] + 1.0, st2[0] + 1.0) < 1e-6 and _rel(st[3], st2[1]) < 1e-6
    d_h2, dres = _rand((rows, H), 4), torch.randn(rows, H, device=DEV)
    dx2_ref, dw2_ref = ops.norm_bwd(d_h2, x2_ref, w2, st2, torch.float32, dres=dres, rms=0, want_dw=True)
    da_ref, dw1_ref = ops.norm_bwd(dx2_ref, a, w1, st1, torch.bfloat16, rms=rms1, want_dw=True)
    dx2, d_a, dw1, dw2 = ops.norm2_bwd(d_h2, x2, w2, dres, a, w1, st, rms1=rms1, rms2=0)
    torch.cuda.synchronize()
    assert _rel(dx2, dx2_ref) < 1e-6 and _rel(d_a, da_ref) < 4e-3
    assert _rel(dw2, dw2_ref) < 1e-6 and _rel(dw1, dw1_ref) < 1e-5
    again = ops.norm2_bwd(d_h2, x2, w2, dres, a, w1, st, rms1=rms1, rms2=0)
    assert all(torch.equal(p, q) for p, q in zip((dx2, d_a, dw1, dw2), again))
