"""GPU: FusedAdamW (AdamW + EMA + bf16 operand packing in one pass) against torch.optim.AdamW, the reference-surface
EMAModel.step and the stand-alone pack kernel, on the micro transformer fixture; CUDA-graph replay of the fused step."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from open_muse_b200 import EMAModel, FusedAdamW, MaskGitTransformer  # noqa: E402

DEV = "cuda"


def _setup(golden):
    g = golden("micro_transformer.pt")
    models = []
    for _ in range(2):
        m = MaskGitTransformer(**g["config"])
        m.load_state_dict(g["state_dict"])
        models.append(m.to(DEV).train())
    return g, models


def _fwd_bwd(m, g):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, loss = m(g["batch"]["input_ids"].to(DEV), labels=g["batch"]["labels"].to(DEV))
    loss.backward()
    # detached: a live loss keeps the autograd graph -- and its AccumulateGrad nodes, bound to the stream they were created
    # on -- alive, which breaks a later CUDA-graph capture of the same model on another stream
    return loss.detach()


@pytest.mark.parametrize("warmup", [False, True])
def test_fused_adamw_ema_matches_torch_adamw_and_emamodel(golden, warmup):
    g, (a, b) = _setup(golden)
    kw = dict(lr=2e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    ref_opt = torch.optim.AdamW(a.parameters(), **kw)
    ref_ema = EMAModel(a.parameters(), decay=0.99, update_after_step=1, update_every=2, use_ema_warmup=warmup, inv_gamma=2.0, power=0.75)
    ema = EMAModel(b.parameters(), decay=0.99, update_after_step=1, update_every=2, use_ema_warmup=warmup, inv_gamma=2.0, power=0.75)
    opt = FusedAdamW(b.parameters(), ema=ema, model=b, **kw)
    for step in range(7):
        la, lb = _fwd_bwd(a, g), _fwd_bwd(b, g)
        assert torch.equal(la, lb), step  # same weights -> same (deterministic) loss before every update
        ref_opt.step(); ref_ema.step(a.parameters()); ref_opt.zero_grad(set_to_none=True)
        # force identical gradients (the two models are bit-identical so far, so these already are)
        opt.step(); opt.zero_grad(set_to_none=True)
        for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
            torch.testing.assert_close(pb, pa, rtol=2e-6, atol=1e-8, msg=lambda m_: f"step {step} {n}: {m_}")
            with torch.no_grad():
                pb.copy_(pa)  # keep the two trajectories locked so that later steps compare like with like
        for i, (sa, sb) in enumerate(zip(ref_ema.shadow_params, ema.shadow_params)):
            torch.testing.assert_close(sb, sa, rtol=2e-6, atol=1e-8, msg=lambda m_: f"step {step} shadow {i}: {m_}")
            sb.copy_(sa)
        assert ema.optimization_step == ref_ema.optimization_step
        assert (ema.cur_decay_value is None) == (ref_ema.cur_decay_value is None)
        assert ema.cur_decay_value is None or abs(ema.cur_decay_value - ref_ema.cur_decay_value) < 1e-12


def test_fused_step_fills_the_packed_operands_and_replays_in_a_graph(golden):
    g, (a, b) = _setup(golden)
    opt = FusedAdamW(b.parameters(), lr=1e-3, model=b)
    _fwd_bwd(b, g)
    opt.step()
    fused_flat = b._packed.flat.clone()
    b._packed.key = None            # force the stand-alone pack kernel over the same (updated) weights
    b._packed.refresh()
    assert torch.equal(fused_flat, b._packed.flat)
    opt.zero_grad(set_to_none=True)
    # losses keep decreasing through the fused step (the packed operands follow the updates without a re-pack)
    losses = []
    for _ in range(6):
        l = _fwd_bwd(b, g)
        opt.step(); opt.zero_grad(set_to_none=True)
        losses.append(float(l))
    assert losses[-1] < losses[0]
    sd = opt.state_dict()
    assert int(next(iter(sd["state"].values()))["step"]) == 7
    # graph capture: two replays advance the device-side step counter and the weights
    from open_muse_b200.graphs import GraphedStep

    ids, lab = g["batch"]["input_ids"].to(DEV), g["batch"]["labels"].to(DEV)

    def step(i, l):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, loss = b(i, labels=l)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    gs = GraphedStep(step, (ids, lab), warmup=2)
    before = int(opt._plans[0]["step"])
    w0 = b.transformer_layers[0].ffn.wo.weight.detach().clone()
    l1 = float(gs(ids, lab)); l2 = float(gs(ids, lab))
    assert int(opt._plans[0]["step"]) == before + 2 and l2 < l1
    assert not torch.equal(w0, b.transformer_layers[0].ffn.wo.weight)
