"""FusedAdamW's host logic WITHOUT a GPU: the pointer table it hands to the C ABI (per-group rows of parameter / gradient /
moment / EMA-shadow / packed-operand addresses), the device-resident step counter, the EMA schedule arguments and its host
mirror, against torch.optim.AdamW + EMAModel.step on the same trajectory -- with ``muse_adamw_ema_step`` replaced by a torch
restatement of csrc/optim.cu that dereferences the same table (tests/cpu_math_ops.py).  GPU twin: tests/test_optim_gpu.py."""
import pytest
import torch

from open_muse_b200 import EMAModel, FusedAdamW, MaskGitTransformer
from tests import cpu_math_ops


def _models(g, n=2):
    out = []
    for _ in range(n):
        m = MaskGitTransformer(**g["config"])
        m.load_state_dict(g["state_dict"])
        out.append(m.train())
    return out


def _fwd_bwd(m, g):
    _, loss = m(g["batch"]["input_ids"], labels=g["batch"]["labels"])
    loss.backward()
    return loss.detach()


@pytest.mark.parametrize("warmup", [False, True])
def test_fused_adamw_ema_host_logic_matches_torch_adamw_and_emamodel(golden, monkeypatch, warmup):
    g = golden("micro_transformer.pt")
    cpu_math_ops.install(monkeypatch, exact=False)
    cpu_math_ops.install_optimizer(monkeypatch)
    a, b = _models(g)
    # the script's name-based groups (train_muse.py:427-437): weight decay on the matrices only
    no_decay = ["bias", "layer_norm.weight", "mlm_ln.weight", "embeddings.weight"]

    def groups(m):
        return [{"params": [p for n, p in m.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 0.05},
                {"params": [p for n, p in m.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]

    kw = dict(lr=2e-3, betas=(0.9, 0.95), eps=1e-8)
    ref_opt = torch.optim.AdamW(groups(a), **kw)
    flat = lambda m: [p for gr in groups(m) for p in gr["params"]]
    ekw = dict(decay=0.99, update_after_step=1, update_every=2, use_ema_warmup=warmup, inv_gamma=2.0, power=0.75)
    ref_ema, ema = EMAModel(flat(a), **ekw), EMAModel(flat(b), **ekw)
    opt = FusedAdamW(groups(b), ema=ema, model=b, **kw)
    for step in range(6):
        la, lb = _fwd_bwd(a, g), _fwd_bwd(b, g)
        assert torch.equal(la, lb), step
        ref_opt.step(); ref_ema.step(flat(a)); ref_opt.zero_grad(set_to_none=True)
        opt.step(); opt.zero_grad(set_to_none=True)
        for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
            torch.testing.assert_close(pb, pa, rtol=2e-6, atol=1e-8, msg=lambda m_: f"step {step} {n}: {m_}")
            with torch.no_grad():
                pb.copy_(pa)
        for i, (sa, sb) in enumerate(zip(ref_ema.shadow_params, ema.shadow_params)):
            torch.testing.assert_close(sb, sa, rtol=2e-6, atol=1e-8, msg=lambda m_: f"step {step} shadow {i}: {m_}")
            sb.copy_(sa)
        assert ema.optimization_step == ref_ema.optimization_step
        assert ema.cur_decay_value is None or abs(ema.cur_decay_value - ref_ema.cur_decay_value) < 1e-12
    sd = opt.state_dict()
    assert int(next(iter(sd["state"].values()))["step"]) == 6 and len(sd["param_groups"]) == 2


def test_fused_step_writes_the_packed_operands(golden, monkeypatch):
    """FusedAdamW(model=...) looks the bf16 destination of every Linear weight up in the model's packed-operand cache and the
    fused pass writes the UPDATED weight there: identical to a stand-alone re-pack, and the next forward needs none."""
    g = golden("micro_transformer.pt")
    cpu_math_ops.install(monkeypatch, exact=False)
    cpu_math_ops.install_optimizer(monkeypatch)
    (b,) = _models(g, 1)
    opt = FusedAdamW(b.parameters(), lr=1e-3, model=b)
    l0 = _fwd_bwd(b, g)
    opt.step()
    fused = b._packed.flat.clone()
    b._packed.key = None
    b._packed.refresh()
    assert torch.equal(fused, b._packed.flat) and bool(fused.abs().sum() > 0)
    opt.zero_grad(set_to_none=True)
    losses = [float(l0)]
    for _ in range(5):
        losses.append(float(_fwd_bwd(b, g)))
        opt.step(); opt.zero_grad(set_to_none=True)
    assert losses[-1] < losses[0]
