"""Host-logic test (no GPU): the autograd plumbing of MaskGitTransformer -- per-layer Functions, packed operand cache,
cross-attention / projected text states, both loss entry points -- with the C-ABI kernels replaced by shape/dtype-checking
stand-ins.  Every parameter must receive an fp32 gradient of its own shape for each wiring the reference configs use; the
numerics are covered by the GPU parity tests."""
import pytest
import torch

from open_muse_b200 import MaskGitTransformer, ops

BF, F32 = torch.bfloat16, torch.float32


def _fake_ops(mp):
    def lin_fwd(x, w, out_dtype=BF, res=None, n_valid=None):
        assert x.dtype == BF and w.dtype == BF and x.shape[1] == w.shape[1], (x.shape, w.shape)
        n = w.shape[0] if n_valid is None else n_valid
        if res is not None:
            assert res.dtype == F32 and res.shape == (x.shape[0], n)
        return torch.zeros(x.shape[0], n, dtype=F32 if res is not None else out_dtype)

    def wgrad(dy, x, out=None):
        assert dy.dtype == BF and x.dtype == BF and dy.shape[0] == x.shape[0]
        return torch.zeros(dy.shape[1], x.shape[1], dtype=F32)

    def embed_bwd(ids, dx, vocab, n_pos):
        assert dx.shape[0] == ids.numel() and dx.dtype == F32
        return torch.zeros(vocab, dx.shape[1]), torch.zeros(n_pos, dx.shape[1])

    def norm_fwd(x, w, eps, out_dtype, res=None, act=0, rms=0, save_stats=True):
        H = x.shape[1] // 2 if act == 2 else x.shape[1]
        assert w is None or w.shape == (H,)
        return torch.zeros(x.shape[0], H, dtype=out_dtype), (torch.zeros(2, x.shape[0]) if save_stats else None)

    def norm_bwd(dy, x, w, stats, dx_dtype, dw=None, dres=None, act=0, rms=0, y_fwd=None, want_dw=False, bf16_copy=False):
        H = x.shape[1] // 2 if act == 2 else x.shape[1]
        assert dy.shape == (x.shape[0], H) and stats is not None
        if dres is not None:
            assert dres.shape == x.shape and dres.dtype == F32
        assert dw is None and want_dw  # the v1 model only uses the reproducible (stored) weight-gradient form
        return torch.zeros(x.shape, dtype=dx_dtype), torch.zeros(H, dtype=F32)

    def norm2_fwd(a, res, w1, w2, eps, rms1=0, rms2=0, save_stats=True):
        assert a.dtype == BF and res.dtype == F32 and a.shape == res.shape and w1.shape == w2.shape == (a.shape[1],)
        return torch.zeros(a.shape, dtype=F32), torch.zeros(a.shape, dtype=BF), (torch.zeros(4, a.shape[0]) if save_stats else None)

    def norm2_bwd(d_h2, x2, w2, dres, a, w1, stats, rms1=0, rms2=0):
        assert d_h2.dtype == BF and x2.dtype == F32 and dres.dtype == F32 and a.dtype == BF and stats.shape[0] == 4
        H = a.shape[1]
        return torch.zeros(a.shape, dtype=F32), torch.zeros(a.shape, dtype=BF), torch.zeros(H), torch.zeros(H)

    def attb(q, k, v, o, do, lse, dq, dk, dv, B, nh, Sq, Skv, sc, head_dim=64):
        assert q.shape[1] == nh * head_dim and do.dtype == BF and do.shape == o.shape and dq.shape == q.shape and dk.shape == k.shape and dv.shape == v.shape

    def gemm(a, b, c, M, N, K, lda, ldb, ldc, a_mn=0, b_mn=0, epi=0, res=None):
        assert a.dtype == BF and b.dtype == BF and c.shape[0] == M and ldc >= N
        return c

    fakes = dict(
        gemm=gemm, linear_fwd=lin_fwd, linear_dgrad=lambda dy, w, out_dtype=BF: torch.zeros(dy.shape[0], w.shape[1], dtype=out_dtype),
        linear_wgrad_det=wgrad, cast_bf16=lambda x: x.to(BF), take_bf16_copy=lambda t: None, pack_bf16=lambda table, n, blocks: None,
        embed_fwd=lambda ids, w, pos: torch.zeros(ids.numel(), w.shape[1]), embed_bwd_det=embed_bwd,
        norm_fwd=norm_fwd, norm_bwd=norm_bwd, norm2_fwd=norm2_fwd, norm2_bwd=norm2_bwd, glu_fwd=lambda ab: torch.zeros(ab.shape[0], ab.shape[1] // 2, dtype=BF),
        glu_bwd=lambda ab, d: torch.zeros_like(ab),
        attn_fwd=lambda q, k, v, B, nh, Sq, Skv, sc, head_dim=64: (torch.zeros(q.shape[0], nh * head_dim, dtype=BF), torch.zeros(B, nh, Sq)),
        attn_bwd=attb, ce_fwd=lambda lg, lab, V, ls: (torch.zeros(2), torch.zeros(2, lg.shape[0])),
        ce_bwd=lambda lg, lab, ws, dl, out, V, ls, row_scale=None: torch.zeros_like(lg))
    for k, v in fakes.items():
        mp.setattr(ops, k, v)
    mp.setattr(torch.Tensor, "is_cuda", property(lambda self: True))  # the model refuses CPU tensors (no fallback)


BASE = dict(vocab_size=72, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
            hidden_dropout=0.0, attention_dropout=0.0, max_position_embeddings=17, codebook_size=64, num_vq_tokens=16)
CFGS = {
    "class-cond-normformer": dict(BASE, num_classes=7),
    "class-cond-head-dim-48": dict(BASE, num_classes=7, hidden_size=96),  # configs/imagenet.yaml: 768 / 16 heads = 48
    "t2i-head-dim-48": dict(BASE, hidden_size=96, max_position_embeddings=16, add_cross_attention=True, encoder_hidden_size=32),
    "t2i-rmsnorm-no-normformer": dict(BASE, max_position_embeddings=16, add_cross_attention=True, encoder_hidden_size=32,
                                      norm_type="rmsnorm", use_normformer=False, use_codebook_size_for_output=True),
    "t2i-projected-text-states": dict(BASE, max_position_embeddings=16, add_cross_attention=True, encoder_hidden_size=32,
                                      project_encoder_hidden_states=True),
}


@pytest.mark.parametrize("name", list(CFGS))
@pytest.mark.parametrize("external_loss", [False, True])
def test_maskgit_transformer_training_plumbing(monkeypatch, name, external_loss):
    _fake_ops(monkeypatch)
    cfg = CFGS[name]
    torch.manual_seed(0)
    m = MaskGitTransformer(**cfg).train()
    S = cfg["max_position_embeddings"]
    ids = torch.randint(0, 64, (3, S))
    lab = torch.randint(0, 64, (3, S))
    kw = dict(encoder_hidden_states=torch.randn(3, 5, 32)) if cfg.get("add_cross_attention") else {}
    if external_loss:  # soft-target style: the script computes its own loss from the returned logits
        logits = m(ids, **kw)
        assert logits.shape == (3, S, m.output_size)
        logits.float().square().mean().backward()
    else:
        logits, loss = m(ids, labels=lab, label_smoothing=0.1, **kw)
        assert logits.shape == (3, S, m.output_size) and loss.dim() == 0
        loss.backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == torch.float32, n
