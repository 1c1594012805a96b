"""GPU parity of the assembled MaskGitTransformer (forward, loss, every gradient) against the outputs of the
unmodified reference stored in tests/golden/ (fp32 CPU) and against the fp32 oracle on larger seeded inputs.

Tolerances (bf16 GEMM operands, fp32 accumulation / statistics, as the reference's own bf16-autocast mode):
  logits: relative L2 error <= 1e-2 (north_star).  Against the fp32 reference fixtures for the 1-2 layer models; for the
          8-layer base-256 and the cc12m-width models, where the reference's OWN bf16-autocast forward is 1.2e-2 away from
          its fp32 forward, the 1e-2 is asserted against the reference recipe executed in the same precision
          mode on the same GPU (oracle ops under torch.autocast(cuda, bf16)), and the distance to fp32 must not exceed
          the reference's own bf16 distance by more than 10 %.
  loss  : relative error <= 2e-3
  grads : relative L2 error <= 6e-2 per parameter (and cosine similarity >= 0.998); a parameter may exceed it only where
          the reference recipe's own bf16 gradients are within 1.5x as noisy (query/key weights of deeper layers at random
          init) -- those parameters are listed in the test output and their number is bounded.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from open_muse_b200.modeling_transformer import MaskGitTransformer  # noqa: E402
from oracle import transformer_oracle as T  # noqa: E402

DEV = "cuda"
LOGIT_TOL, LOSS_TOL, GRAD_TOL = 1e-2, 2e-3, 6e-2


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _bf16_recipe_grad_errors(g, **fwd_kwargs):
    """Calibration: per-parameter gradient error of the REFERENCE RECIPE's own bf16-autocast mode (oracle under
    torch.autocast on CPU) against its fp32 gradients.  Query/key weights of later layers have true gradients ~1e-6 at
    random init (near-uniform attention), five orders below the rest, and sit at ~0.2 relative error in any bf16 path."""
    q = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        _, loss = T.forward(q, g["config"], **fwd_kwargs)
    loss.backward()
    return {k: _rel(q[k].grad, g["grads"][k]) for k in q}


def _same_mode_reference(sd, cfg, input_ids, labels, **kw):
    """The reference recipe in the SAME precision mode on the SAME device: oracle ops (= the reference's torch calls)
    under torch.autocast(cuda, bf16), i.e. what the unmodified reference computes on this GPU.  Checker only."""
    q = {k: v.detach().clone().float().to(DEV).requires_grad_(True) for k, v in sd.items()}
    kw = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in kw.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, loss = T.forward(q, cfg, input_ids.to(DEV), labels=labels.to(DEV), **kw)
    loss.backward()
    return logits.detach().float().cpu(), loss.detach().float().cpu(), {k: v.grad.cpu() for k, v in q.items()}


def _check(model, logits, loss, ref_logits, ref_loss, ref_grads, report, recipe_err=None, logit_tol=LOGIT_TOL,
           max_hatch=0):
    r = _rel(logits, ref_logits)
    report.append(f"logits rel-L2 {r:.3e}")
    assert r < logit_tol
    lr = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
    report.append(f"loss rel {lr:.3e}")
    assert lr < LOSS_TOL
    worst, hatch = 0.0, []
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        g, rg = p.grad.float().cpu(), ref_grads[n].float()
        e = _rel(g, rg)
        cos = float(torch.nn.functional.cosine_similarity(g.flatten(), rg.flatten(), dim=0))
        floor = 1.5 * recipe_err[n] if recipe_err is not None else 0.0
        if e >= GRAD_TOL:
            assert e < floor, (n, e, cos, floor)  # only allowed where the reference's own bf16 mode is as noisy
            assert "attention.query" in n or "attention.key" in n, (n, e)  # ... which is the ~1e-6 q/k gradients only
            hatch.append(f"{n}={e:.2e} (reference bf16 recipe {recipe_err[n]:.2e})")
        else:
            assert cos > 0.998, (n, e, cos)
            worst = max(worst, e)
    report.append(f"worst grad rel-L2 {worst:.3e}")
    report.append(f"{len(hatch)} parameter(s) above {GRAD_TOL:g} within 1.5x of the reference's own bf16 noise: {hatch}")
    assert len(hatch) <= max_hatch, hatch


def test_micro_class_conditional_vs_reference(golden):
    g = golden("micro_transformer.pt")
    m = MaskGitTransformer(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, loss = m(g["batch"]["input_ids"].to(DEV), labels=g["batch"]["labels"].to(DEV),
                         label_smoothing=g["label_smoothing"])
    assert logits.dtype == torch.bfloat16 and logits.shape == g["logits"].shape and loss.dtype == torch.float32
    loss.backward()
    rep = []
    cal = _bf16_recipe_grad_errors(g, input_ids=g["batch"]["input_ids"], labels=g["batch"]["labels"],
                                   label_smoothing=g["label_smoothing"])
    _check(m, logits, loss, g["logits"], g["loss"], g["grads"], rep, recipe_err=cal, max_hatch=2)
    print("micro:", "; ".join(rep))


def test_micro_text_conditional_vs_reference(golden):
    """cross-attention + RMSNorm + no normformer + codebook-sized output (the cc12m-style wiring)."""
    g = golden("micro_t2i_transformer.pt")
    m = MaskGitTransformer(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, loss = m(g["input_ids"].to(DEV), encoder_hidden_states=g["encoder_hidden_states"].to(DEV),
                         labels=g["labels"].to(DEV), cond_embeds=None, loss_weight=None, micro_conds=None)
    loss.backward()
    rep = []
    _check(m, logits, loss, g["logits"], g["loss"], g["grads"], rep)
    print("micro t2i:", "; ".join(rep))


def test_micro_text_conditional_projected_encoder_states_vs_reference(golden):
    """project_encoder_hidden_states=True: encoder_proj + norm in front of every cross-attention; seeded construction
    reproduces the reference's initial weights (construction order), gradients reach encoder_proj through all layers."""
    g = golden("micro_t2i_proj_transformer.pt")
    torch.manual_seed(g["seed"])
    m = MaskGitTransformer(**g["config"])
    for k, v in g["state_dict"].items():
        assert torch.equal(m.state_dict()[k], v), k
    m.to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, loss = m(g["input_ids"].to(DEV), encoder_hidden_states=g["encoder_hidden_states"].to(DEV),
                         labels=g["labels"].to(DEV))
    loss.backward()
    rep = []
    cal = _bf16_recipe_grad_errors(g, input_ids=g["input_ids"], labels=g["labels"],
                                   encoder_hidden_states=g["encoder_hidden_states"])
    _check(m, logits, loss, g["logits"], g["loss"], g["grads"], rep, recipe_err=cal, max_hatch=4)
    assert float(m.encoder_proj.weight.grad.abs().max()) > 0
    print("micro t2i proj:", "; ".join(rep))


def test_tiny_config1_vs_reference(golden):
    """BASELINE config 1 (L2, H128, S257, V2025, B2): seeded init == reference init, then fwd+bwd parity."""
    g = golden("tiny_transformer.pt")
    torch.manual_seed(g["seed"])
    m = MaskGitTransformer(**g["config"]).to(DEV).train()
    b = g["batch"]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, loss = m(b["input_ids"].to(DEV), labels=b["labels"].to(DEV))
    loss.backward()
    assert _rel(logits[:, ::16, ::25], g["logits_slice"]) < LOGIT_TOL
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < LOSS_TOL
    for n, p in m.named_parameters():
        gn = float(p.grad.float().norm())
        assert abs(gn - float(g["grad_norms"][n])) <= GRAD_TOL * float(g["grad_norms"][n]) + 1e-8, n


def test_head_dim_48_vs_reference_and_oracle(golden):
    """configs/imagenet.yaml's head dimension (48): seeded init == reference init; logits slice, loss and gradient norms
    against the fixture of the unmodified reference; every gradient against the oracle's fp32 autograd."""
    g = golden("hd48_transformer.pt")
    torch.manual_seed(g["seed"])
    m = MaskGitTransformer(**g["config"])
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    b = g["batch"]
    ref_logits, ref_loss, ref_grads = T.forward_backward(sd, g["config"], b["input_ids"], b["labels"],
                                                         label_smoothing=g["label_smoothing"])
    m.to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, loss = m(b["input_ids"].to(DEV), labels=b["labels"].to(DEV), label_smoothing=g["label_smoothing"])
    loss.backward()
    assert _rel(logits[:, ::8, ::16], g["logits_slice"]) < LOGIT_TOL
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < LOSS_TOL
    assert _rel(logits, ref_logits) < LOGIT_TOL
    worst, worst_qk = 0.0, 0.0
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        gr, rg = p.grad.float().cpu(), ref_grads[n].float()
        e = _rel(gr, rg)
        if "attention.query" in n or "attention.key" in n:
            # ~1e-6 gradients at random init (near-uniform attention): the reference's own bf16 recipe sits at ~0.2 there
            worst_qk = max(worst_qk, e)
            assert e < 0.3, (n, e)
        else:
            cos = float(torch.nn.functional.cosine_similarity(gr.flatten(), rg.flatten(), dim=0))
            assert e < GRAD_TOL and cos > 0.998, (n, e, cos)
            worst = max(worst, e)
            assert abs(float(gr.norm()) - float(g["grad_norms"][n])) <= GRAD_TOL * float(g["grad_norms"][n]) + 1e-8, n
    print(f"head_dim 48 (768 / 16 heads), L=2, B=2: logits rel-L2 {_rel(logits, ref_logits):.3e}; worst grad rel-L2 {worst:.3e} "
          f"(query / key weights {worst_qk:.3e})")


def test_base_shape_vs_oracle():
    """Base-256 architecture (8x512, S257, V2025) at a small batch against the fp32 oracle on the same seeded
    weights and the training masking recipe -- the size-independent check of the benchmark configuration."""
    cfg = dict(vocab_size=2025, max_position_embeddings=257, hidden_size=512, num_hidden_layers=8,
               num_attention_heads=8, intermediate_size=2048, codebook_size=1024, num_vq_tokens=256, num_classes=1000,
               hidden_dropout=0.0, attention_dropout=0.0)
    torch.manual_seed(0)
    m = MaskGitTransformer(**cfg)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    B = 4
    tokens = torch.randint(0, 1024, (B, 256), generator=g)
    cls = torch.randint(0, 1000, (B,), generator=g)
    inp, lab = T.mask_tokens(tokens, cls, torch.rand(B, generator=g), torch.rand(B, 256, generator=g), 1024, 2024)
    ref_logits, ref_loss, ref_grads = T.forward_backward(sd, cfg, inp, lab)
    same_logits, same_loss, _ = _same_mode_reference(sd, cfg, inp, lab)
    m.to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, loss = m(inp.to(DEV), labels=lab.to(DEV))
    loss.backward()
    rep = []
    r_same, r_ref_own = _rel(logits, same_logits), _rel(same_logits, ref_logits)
    rep.append(f"logits vs reference recipe in bf16 autocast on this GPU {r_same:.3e} (that recipe vs fp32: {r_ref_own:.3e})")
    assert r_same < LOGIT_TOL
    assert abs(float(loss) - float(same_loss)) / float(same_loss) < LOSS_TOL
    # against fp32: no further from it than the reference's own bf16 mode (+10 %)
    _check(m, logits, loss, ref_logits, ref_loss, ref_grads, rep, logit_tol=max(LOGIT_TOL, 1.1 * r_ref_own))
    print("base-256 B=4:", "; ".join(rep))
    # linearity-style property at the loss level: two identical half-batches give the same loss as one
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        _, l2 = m(torch.cat([inp, inp]).to(DEV), labels=torch.cat([lab, lab]).to(DEV))
    assert abs(float(l2) - float(loss)) < 1e-4 * float(loss)


def test_cc12m_width_text_conditional_vs_oracle():
    """BASELINE config 4 at its own widths (configs/cc12m_uvit_clip.yaml:29-54: H 1024, 16 heads, I 4096, vocab 8256 with
    codebook-sized output 8192, cross-attention to 77 x 768 CLIP states, RMSNorm, no normformer, eps 1e-6), two layers,
    batch 2: logits, loss and every gradient against the fp32 oracle, and logits against the reference recipe in bf16
    autocast on this GPU."""
    cfg = dict(vocab_size=8256, max_position_embeddings=256, hidden_size=1024, num_hidden_layers=2,
               num_attention_heads=16, intermediate_size=4096, codebook_size=8192, num_vq_tokens=256,
               add_cross_attention=True, encoder_hidden_size=768, norm_type="rmsnorm", layer_norm_eps=1e-6,
               use_normformer=False, use_codebook_size_for_output=True, hidden_dropout=0.0, attention_dropout=0.0)
    torch.manual_seed(0)
    m = MaskGitTransformer(**cfg)
    assert m.output_size == 8192 and m.config.mask_token_id == 8255
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    B = 2
    tokens = torch.randint(0, 8192, (B, 256), generator=g)
    mask = torch.rand(B, 256, generator=g) < 0.6
    inp = torch.where(mask, 8255, tokens)
    lab = torch.where(mask, tokens, -100)
    ehs = torch.randn(B, 77, 768, generator=g)
    ref_logits, ref_loss, ref_grads = T.forward_backward(sd, cfg, inp, lab, encoder_hidden_states=ehs)
    same_logits, same_loss, same_grads = _same_mode_reference(sd, cfg, inp, lab, encoder_hidden_states=ehs)
    cal = {k: _rel(same_grads[k], ref_grads[k]) for k in ref_grads}  # the reference recipe's own bf16 gradient noise
    m.to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, loss = m(inp.to(DEV), encoder_hidden_states=ehs.to(DEV), labels=lab.to(DEV))
    assert logits.shape == (B, 256, 8192)
    loss.backward()
    rep = []
    r_same, r_ref_own = _rel(logits, same_logits), _rel(same_logits, ref_logits)
    rep.append(f"logits vs reference recipe in bf16 autocast on this GPU {r_same:.3e} (that recipe vs fp32: {r_ref_own:.3e})")
    assert r_same < LOGIT_TOL
    _check(m, logits, loss, ref_logits, ref_loss, ref_grads, rep, recipe_err=cal, logit_tol=max(LOGIT_TOL, 1.1 * r_ref_own),
           max_hatch=8)
    print("cc12m widths L=2 B=2:", "; ".join(rep))


def test_train_step_gradients_are_bit_reproducible():
    """The reference's bf16 step is run-to-run bit-identical.  Two identical forward+backward passes of the
    base-256 architecture must give bit-identical logits, loss and gradients for EVERY parameter: weight gradients come
    from fixed-order reductions (deterministic split-K GEMM, ordered column sums, sorted embedding backward), no atomics."""
    cfg = dict(vocab_size=2025, max_position_embeddings=257, hidden_size=512, num_hidden_layers=3,
               num_attention_heads=8, intermediate_size=2048, codebook_size=1024, num_vq_tokens=256, num_classes=1000,
               hidden_dropout=0.0, attention_dropout=0.0)
    torch.manual_seed(0)
    m = MaskGitTransformer(**cfg).to(DEV).train()
    g = torch.Generator().manual_seed(5)
    B = 24
    tokens = torch.randint(0, 1024, (B, 256), generator=g)
    cls = torch.randint(0, 1000, (B,), generator=g)
    inp, lab = T.mask_tokens(tokens, cls, torch.rand(B, generator=g), torch.rand(B, 256, generator=g), 1024, 2024)
    inp, lab = inp.to(DEV), lab.to(DEV)
    runs = []
    for _ in range(3):
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, loss = m(inp, labels=lab)
        loss.backward()
        torch.cuda.synchronize()
        runs.append((logits.clone(), loss.clone(), {n: p.grad.clone() for n, p in m.named_parameters()}))
    for other in runs[1:]:
        assert torch.equal(runs[0][0], other[0]) and torch.equal(runs[0][1], other[1])
        for n, gr in runs[0][2].items():
            assert torch.equal(gr, other[2][n]), n


def test_programmatic_dependent_launch_is_bit_identical(golden):
    """include/muse_b200.h muse_set_pdl: with programmatic dependent launch every kernel may be scheduled while its
    predecessor drains, but blocks (griddepcontrol.wait) before its first global access.  Memory effects must therefore be
    exactly those of plain stream order: logits, loss and EVERY gradient of a train step bit-identical with the switch on
    and off -- launched one by one and replayed from a captured CUDA graph -- and the same generate2 ids."""
    from open_muse_b200 import ops
    from open_muse_b200.graphs import GraphedStep

    cfg = dict(vocab_size=2025, max_position_embeddings=257, hidden_size=512, num_hidden_layers=3,
               num_attention_heads=8, intermediate_size=2048, codebook_size=1024, num_vq_tokens=256, num_classes=1000,
               hidden_dropout=0.0, attention_dropout=0.0)
    g = torch.Generator().manual_seed(11)
    B = 24
    tokens = torch.randint(0, 1024, (B, 256), generator=g)
    cls = torch.randint(0, 1000, (B,), generator=g)
    inp, lab = T.mask_tokens(tokens, cls, torch.rand(B, generator=g), torch.rand(B, 256, generator=g), 1024, 2024)
    inp, lab = inp.to(DEV), lab.to(DEV)
    gm = golden("micro_transformer.pt")
    was = ops.get_pdl()
    results = {}
    try:
        for pdl in (False, True):
            ops.set_pdl(pdl)
            assert ops.get_pdl() == pdl
            torch.manual_seed(0)
            m = MaskGitTransformer(**cfg).to(DEV).train()

            def fwd_bwd(i, l):
                m.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    logits, loss = m(i, labels=l)
                loss.backward()
                return logits, loss

            # graph first: the first backward of this model then runs on GraphedStep's side stream (as in bench.py); an eager
            # backward on the legacy default stream BEFORE the capture would tie the AccumulateGrad nodes to that stream and
            # the capture would fail with cudaErrorStreamCaptureImplicit, whatever the launch mode
            graphed = GraphedStep(fwd_bwd, (inp, lab))
            for _ in range(2):
                logits, loss = graphed(inp, lab)
            torch.cuda.synchronize()
            replay = (logits.clone(), loss.clone(), {n: p.grad.clone() for n, p in m.named_parameters()})
            logits, loss = fwd_bwd(inp, lab)
            torch.cuda.synchronize()
            eager = (logits.clone(), loss.clone(), {n: p.grad.clone() for n, p in m.named_parameters()})
            assert torch.equal(eager[0], replay[0]) and torch.equal(eager[1], replay[1])
            for n in eager[2]:
                assert torch.equal(eager[2][n], replay[2][n]), (pdl, n)
            # the captured decode loop (fresh model instance: its graph cache is keyed without the launch mode)
            mg = MaskGitTransformer(**gm["config"])
            mg.load_state_dict(gm["state_dict"])
            mg.to(DEV).eval()
            ids = []
            for use_graph in (False, True):
                gen = torch.Generator(device=DEV).manual_seed(7)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    ids.append(mg.generate2(class_ids=torch.tensor([1, 5, 0, 3], device=DEV), timesteps=4, generator=gen,
                                            use_cuda_graph=use_graph))
            assert torch.equal(ids[0], ids[1])
            results[pdl] = (eager, ids[0])
    finally:
        ops.set_pdl(was)
    off, on = results[False], results[True]
    assert torch.equal(off[0][0], on[0][0]) and torch.equal(off[0][1], on[0][1])
    for n in off[0][2]:
        assert torch.equal(off[0][2][n], on[0][2][n]), n
    assert torch.equal(off[1], on[1])


def test_soft_target_loss_on_returned_logits(golden):
    """train_maskgit_imagenet.py:101-117: with soft targets the script computes its own loss from the returned logits,
    so the gradient reaches the model through the logits output (not through the fused CE).  Parity vs the oracle's fp32
    autograd on the same custom loss."""
    g = golden("micro_transformer.pt")
    cfg, b = g["config"], g["batch"]
    gen = torch.Generator().manual_seed(21)
    soft = torch.softmax(torch.randn(b["input_ids"].shape[0], b["input_ids"].shape[1] - 1, 64, generator=gen), -1)

    def soft_ce(logits, targets, soft_targets):  # restated from the training script
        logits, targets = logits[:, 1:], targets[:, 1:]
        logp = torch.log_softmax(logits[..., : soft_targets.shape[-1]].float(), dim=-1)
        pad = targets.eq(-100)
        loss = torch.sum(-soft_targets * logp, dim=-1).masked_fill(pad, 0.0)
        return loss.sum() / (pad.numel() - pad.long().sum())

    q = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    ref_loss = soft_ce(T.forward(q, cfg, b["input_ids"]), b["labels"], soft)
    ref_loss.backward()
    m = MaskGitTransformer(**cfg)
    m.load_state_dict(g["state_dict"])
    m.to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(b["input_ids"].to(DEV))
    loss = soft_ce(logits, b["labels"].to(DEV), soft.to(DEV))
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) / float(ref_loss) < LOSS_TOL
    worst = 0.0
    for n, p in m.named_parameters():
        e = _rel(p.grad, q[n].grad)
        if "attention.query" in n or "attention.key" in n:  # ~1e-6 gradients at random init (see _bf16_recipe_grad_errors)
            continue
        worst = max(worst, e)
    assert worst < GRAD_TOL, worst


def test_forward_is_deterministic_and_eval_matches_train(golden):
    g = golden("micro_transformer.pt")
    m = MaskGitTransformer(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.to(DEV).eval()
    ids = g["batch"]["input_ids"].to(DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a = m(ids)
        b = m(ids)
    assert torch.equal(a, b)
    out = m(ids)  # outside autocast the reference returns fp32 logits
    assert out.dtype == torch.float32


def test_save_load_roundtrip_and_optimizer_step(tmp_path, golden):
    g = golden("micro_transformer.pt")
    m = MaskGitTransformer(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.save_pretrained(tmp_path)
    assert sorted(os.listdir(tmp_path)) == ["config.json", "pytorch_model.bin"]
    m2 = MaskGitTransformer.from_pretrained(tmp_path).to(DEV)
    assert not m2.training
    m2.train()
    opt = torch.optim.AdamW(m2.parameters(), lr=1e-3)
    ids, lab = g["batch"]["input_ids"].to(DEV), g["batch"]["labels"].to(DEV)
    losses = []
    for _ in range(8):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, loss = m2(ids, labels=lab)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(float(loss))
    assert losses[-1] < losses[0] - 0.05, losses  # packed bf16 weights follow the optimizer updates


def test_generate2_runs_and_is_seed_deterministic(golden):
    g = golden("micro_transformer.pt")
    m = MaskGitTransformer(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.to(DEV).eval()
    outs = []
    for _ in range(2):
        cls = torch.tensor([1, 5, 0, 3], device=DEV)
        gen = torch.Generator(device=DEV).manual_seed(7)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ids = m.generate2(class_ids=cls, timesteps=4, generator=gen)
        assert ids.shape == (4, 16) and int(ids.max()) < 64 and int(ids.min()) >= 0
        assert torch.equal(cls.cpu(), torch.tensor([1, 5, 0, 3]) + 64)  # in-place shift, quirk Q3
        outs.append(ids)
    assert torch.equal(outs[0], outs[1])


def test_ema_copy_to_refreshes_the_packed_operands(golden):
    """training/train_muse.py:857-908 validates with EMA weights: ema.store / ema.copy_to / forward / ema.restore.  The
    forward after copy_to must use the EMA weights in every GEMM (the packed bf16 operand cache follows the parameters'
    version counters), and the forward after restore must reproduce the original output."""
    from open_muse_b200 import EMAModel

    g = golden("micro_transformer.pt")
    m = MaskGitTransformer(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.to(DEV).eval()
    ids = g["batch"]["input_ids"].to(DEV)
    ema = EMAModel(m.parameters(), decay=0.5)
    with torch.no_grad():
        base = m(ids).clone()
        gen = torch.Generator(device=DEV).manual_seed(3)
        for sp in ema.shadow_params:  # make the EMA weights differ from the live ones
            sp.add_(torch.randn(sp.shape, device=sp.device, generator=gen) * 0.02)
        ema.store(m.parameters())
        ema.copy_to(m.parameters())
        with_ema = m(ids).clone()
        fresh = MaskGitTransformer(**g["config"]).to(DEV).eval()
        fresh.load_state_dict({k: p.detach().clone() for k, p in m.state_dict().items()})
        expect = fresh(ids)
        ema.restore(m.parameters())
        back = m(ids)
    assert not torch.equal(with_ema, base)
    assert torch.equal(with_ema, expect)
    assert torch.equal(back, base)
