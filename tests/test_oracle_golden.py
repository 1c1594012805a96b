"""CPU: the oracle (oracle/*.py, oracle/vq_oracle.c) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  This is what "pins" the oracle; the -m gpu tests then compare CUDA to it."""
import numpy as np
import torch

from oracle import transformer_oracle as T
from oracle import vq_oracle as VQ
from oracle import vqgan_oracle as G


def _close(a, b, rtol, atol):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_micro_transformer_forward_backward(golden):
    g = golden("micro_transformer.pt")
    logits, loss, grads = T.forward_backward(g["state_dict"], g["config"], g["batch"]["input_ids"], g["batch"]["labels"],
                                             label_smoothing=g["label_smoothing"])
    _close(logits, g["logits"], 1e-5, 1e-6)
    _close(loss, g["loss"], 1e-6, 0)
    assert set(grads) == set(g["grads"])
    for k, v in g["grads"].items():
        _close(grads[k], v, 1e-4, 1e-7)


def test_micro_t2i_transformer_forward_backward(golden):
    g = golden("micro_t2i_transformer.pt")
    logits, loss, grads = T.forward_backward(g["state_dict"], g["config"], g["input_ids"], g["labels"],
                                             encoder_hidden_states=g["encoder_hidden_states"])
    assert logits.shape[-1] == 64  # use_codebook_size_for_output
    _close(logits, g["logits"], 1e-5, 1e-6)
    _close(loss, g["loss"], 1e-6, 0)
    for k, v in g["grads"].items():
        _close(grads[k], v, 1e-4, 1e-7)


def test_micro_t2i_projected_encoder_states_forward_backward(golden):
    g = golden("micro_t2i_proj_transformer.pt")
    logits, loss, grads = T.forward_backward(g["state_dict"], g["config"], g["input_ids"], g["labels"],
                                             encoder_hidden_states=g["encoder_hidden_states"])
    _close(logits, g["logits"], 1e-5, 1e-6)
    _close(loss, g["loss"], 1e-6, 0)
    assert "encoder_proj.weight" in grads and "encoder_proj_layer_norm.weight" in grads
    for k, v in g["grads"].items():
        _close(grads[k], v, 1e-4, 1e-7)


def test_micro_conv_in_out_forward_backward(golden):
    """use_conv_in_out=True (ConvEmbed / ConvMlmLayer): oracle == unmodified reference, logits / loss / every gradient."""
    g = golden("micro_conv_transformer.pt")
    logits, loss, grads = T.forward_backward(g["state_dict"], g["config"], g["input_ids"], g["labels"],
                                             encoder_hidden_states=g["encoder_hidden_states"],
                                             label_smoothing=g["label_smoothing"])
    _close(logits, g["logits"], 1e-5, 1e-6)
    _close(loss, g["loss"], 1e-6, 0)
    assert set(grads) == set(g["grads"])
    for k, v in g["grads"].items():
        _close(grads[k], v, 1e-4, 1e-7)


def test_micro_uvit_v2_forward_loss_and_generate2(golden):
    """MaskGiTUViT_v2 restatement (oracle/transformer_v2_oracle.py) against the unmodified reference: logits, both loss
    forms, and the CFG generate2 id trace through the same torch generator."""
    from oracle import transformer_v2_oracle as V2

    g = golden("micro_uvit_v2.pt")
    p, cfg = g["state_dict"], g["config"]
    with torch.no_grad():
        logits, loss = V2.forward(p, cfg, g["input_ids"], g["encoder_hidden_states"], g["cond_embeds"], g["micro_conds"],
                                  labels=g["labels"], label_smoothing=0.1)
        _close(logits, g["logits"], 1e-4, 2e-5)
        _close(loss, g["loss"], 1e-5, 0)
        _, loss_w = V2.forward(p, cfg, g["input_ids"], g["encoder_hidden_states"], g["cond_embeds"], g["micro_conds"],
                               labels=g["labels"], loss_weight=g["loss_weight"])
        _close(loss_w, g["loss_weighted"], 1e-5, 0)
        ids = V2.generate2(p, cfg, g["encoder_hidden_states"], g["cond_embeds"], g["micro_conds"][:1], g["empty_embeds"],
                           g["empty_cond_embeds"], timesteps=4, temperature=(2.0, 0.0), guidance_scale=3.0,
                           generator=torch.Generator().manual_seed(g["gen_seed"]), seq_len=16)
    assert torch.equal(ids, g["gen_ids"])


def test_micro_uvit_v2_oracle_gradients_match_the_reference(golden):
    """the oracle's autograd gradients (what the hand-written U-ViT backward is compared with) against the gradient signatures
    of the UNMODIFIED MaskGiTUViT_v2: per-parameter norm and leading elements, plain and loss_weight losses, with and
    without force_down_up_sample"""
    from oracle import transformer_v2_oracle as V2

    sig = golden("micro_uvit_v2_grads.pt")
    for key, ref in sig.items():
        name, form = key.split("/")
        g = golden(name + ".pt")
        kw = dict(label_smoothing=0.1) if form == "plain" else dict(loss_weight=g["loss_weight"])
        q = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
        _, loss = V2.forward(q, g["config"], g["input_ids"], g["encoder_hidden_states"], g["cond_embeds"], g["micro_conds"],
                             labels=g["labels"], **kw)
        loss.backward()
        _close(loss.detach(), ref["loss"], 1e-5, 0)
        assert set(ref["norms"]) == set(q)
        for k in q:
            _close(q[k].grad.norm(), ref["norms"][k], 2e-4, 1e-9)
            _close(q[k].grad.flatten()[:8], ref["heads"][k], 2e-3, 1e-8)


def test_micro_uvit_v2_force_down_up_sample(golden):
    """force_down_up_sample=True (Norm2D + k2s2 conv / Norm2D + ConvTranspose2d, modeling_transformer_v2.py:505-583) against
    the unmodified reference: the two resampling outputs, logits, loss and the generate2 id trace."""
    from oracle import transformer_v2_oracle as V2

    g = golden("micro_uvit_v2_downup.pt")
    p, cfg = g["state_dict"], g["config"]
    with torch.no_grad():
        st = {}
        logits, loss = V2.forward(p, cfg, g["input_ids"], g["encoder_hidden_states"], g["cond_embeds"], g["micro_conds"],
                                  labels=g["labels"], label_smoothing=0.1, stages=st)
        _close(st["downsample"], g["stages"]["downsample"], 1e-4, 1e-5)
        _close(st["upsample"], g["stages"]["upsample"], 1e-4, 2e-5)
        _close(logits, g["logits"], 1e-4, 2e-5)
        _close(loss, g["loss"], 1e-5, 0)
        ids = V2.generate2(p, cfg, g["encoder_hidden_states"], g["cond_embeds"], g["micro_conds"], g["empty_embeds"],
                           g["empty_cond_embeds"], timesteps=4, temperature=(2.0, 0.0), guidance_scale=3.0,
                           generator=torch.Generator().manual_seed(g["gen_seed"]), seq_len=64)
    assert torch.equal(ids, g["gen_ids"])


def test_micro_taming_vqgan(golden):
    """taming VQGANModel restatement (strided Downsample with (0,1,0,1) padding, AttnBlocks, input-side shortcuts)."""
    from oracle import taming_vqgan_oracle as TG

    g = golden("micro_taming_vqgan.pt")
    with torch.no_grad():
        z, zq, ids = TG.encode(g["state_dict"], g["config"], g["image"])
        _close(z, g["z"], 1e-5, 1e-6)
        assert torch.equal(ids, g["ids"])
        _close(zq, g["z_q"], 0, 0)
        _close(TG.decoder(g["state_dict"], g["config"], g["z_q"]), g["recon"], 1e-5, 1e-5)


def test_masking_recipe(golden):
    b = golden("micro_transformer.pt")["batch"]
    inp, lab = T.mask_tokens(b["tokens"], b["class_ids"], b["timesteps"], b["rand"], 64, 71)
    assert torch.equal(inp, b["input_ids"]) and torch.equal(lab, b["labels"])
    b = golden("tiny_transformer.pt")["batch"]
    inp, lab = T.mask_tokens(b["tokens"], b["class_ids"], b["timesteps"], b["rand"], 1024, 2024)
    assert torch.equal(inp, b["input_ids"]) and torch.equal(lab, b["labels"])


def test_micro_generate2_matches_reference_stream(golden):
    g = golden("micro_generate2.pt")
    p = golden("micro_transformer.pt")
    for steps in (4, 7):
        gen = torch.Generator().manual_seed(g["seed"])
        with torch.no_grad():
            ids = T.generate2(p["state_dict"], p["config"], g["class_ids"].clone(), steps, g["temperature"], gen)
        assert torch.equal(ids, g["ids"][f"micro_generate2_steps{steps}"])


def test_micro_t2i_generate2_classifier_free_guidance(golden):
    """generate2 on the text-conditional micro model: guidance with zero / explicit negative states, given start tokens."""
    w = golden("micro_t2i_transformer.pt")
    g = golden("micro_t2i_generate2.pt")
    with torch.no_grad():
        ids = T.generate2(w["state_dict"], w["config"], None, 4, 1.0, torch.Generator().manual_seed(5),
                          encoder_hidden_states=g["encoder_hidden_states"], guidance_scale=3.0)
        assert torch.equal(ids, g["cfg_ids"])
        ids = T.generate2(w["state_dict"], w["config"], None, 3, 0.7, torch.Generator().manual_seed(6),
                          encoder_hidden_states=g["encoder_hidden_states"], negative_embeds=g["negative_embeds"],
                          guidance_scale=1.5, input_ids=g["start"].clone())
        assert torch.equal(ids, g["neg_ids"]) and torch.equal(ids[:, :3], g["start"][:, :3])


def test_sample_step_equals_generator_path(golden):
    """The pre-drawn-noise formulation the CUDA kernel implements == the torch.multinomial formulation."""
    p = golden("micro_transformer.pt")
    trace = []
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        T.generate2(p["state_dict"], p["config"], torch.tensor([2, 4]), 3, 1.0, gen, trace=trace)
    gen = torch.Generator().manual_seed(11)
    ids = torch.full((2, 16), 71, dtype=torch.long)
    for st in trace:
        probs = st["probs"]
        # ATen multinomial(n=1): q ~ Exp(1) drawn with the generator over probs.shape, result argmax(p / q)
        q = torch.empty_like(probs.reshape(-1, 64)).exponential_(1, generator=gen).view_as(probs)
        u = torch.zeros(2, 16).uniform_(0, 1, generator=gen)
        sampled, nxt = T.sample_step(probs, ids, 71, q, u, st["mask_len"], st["temperature"])
        assert torch.equal(sampled, st["sampled"])
        assert torch.equal(nxt == 71, st["masking"])
        ids = nxt


def test_tiny_config1_seeded(golden):
    """BASELINE config 1: our facade's seeded init == the reference's, and oracle(loss, logits, grads) == reference."""
    from open_muse_b200.modeling_transformer import MaskGitTransformer

    g = golden("tiny_transformer.pt")
    torch.manual_seed(g["seed"])
    m = MaskGitTransformer(**g["config"])
    sd = m.state_dict()
    for k, n in g["param_norms"].items():
        _close(sd[k].norm(), n, 1e-6, 0)
    logits, loss, grads = T.forward_backward(sd, g["config"], g["batch"]["input_ids"], g["batch"]["labels"])
    _close(loss, g["loss"], 1e-6, 0)
    _close(logits[:, ::16, ::25], g["logits_slice"], 1e-4, 1e-6)
    for k, n in g["grad_norms"].items():
        _close(grads[k].norm(), n, 1e-4, 1e-8)
        _close(grads[k].flatten()[:8], g["grad_heads"][k], 1e-3, 1e-8)


def test_head_dim_48_seeded(golden):
    """configs/imagenet.yaml at its own widths (hidden 768, 16 heads -> head_dim 48, intermediate 3072, vocabulary 2048,
    264 positions, its norm switches) with 2 of its 24 layers: our facade's seeded init == the reference's, oracle(loss, logits, grads) == reference."""
    from open_muse_b200.modeling_transformer import MaskGitTransformer

    g = golden("hd48_transformer.pt")
    assert g["config"]["hidden_size"] // g["config"]["num_attention_heads"] == 48
    torch.manual_seed(g["seed"])
    m = MaskGitTransformer(**g["config"])
    sd = m.state_dict()
    assert set(sd) == set(g["param_norms"])
    for k, n in g["param_norms"].items():
        _close(sd[k].norm(), n, 1e-6, 0)
    logits, loss, grads = T.forward_backward(sd, g["config"], g["batch"]["input_ids"], g["batch"]["labels"],
                                             label_smoothing=g["label_smoothing"])
    _close(loss, g["loss"], 1e-6, 0)
    _close(logits[:, ::8, ::16], g["logits_slice"], 1e-4, 1e-6)
    for k, n in g["grad_norms"].items():
        _close(grads[k].norm(), n, 1e-4, 1e-8)
        _close(grads[k].flatten()[:8], g["grad_heads"][k], 1e-3, 1e-8)


def test_vq_oracle_c_matches_reference_ids(golden):
    g = golden("vq_quantizer.pt")
    z = VQ.nchw_to_rows(g["z"].numpy())
    ids, dmin = VQ.argmin(z, g["codebook"].numpy())
    assert np.array_equal(ids.reshape(2, -1), g["ids"].numpy())
    np.testing.assert_allclose(dmin, g["dmin"].numpy(), rtol=1e-5, atol=1e-4)
    # numpy restatement of the reference formula agrees as well
    d = VQ.distances_numpy(z, g["codebook"].numpy())
    assert np.array_equal(d.argmin(axis=1).reshape(2, -1), g["ids"].numpy())
    np.testing.assert_array_equal(VQ.codebook_entry_nchw(g["ids"].numpy(), g["codebook"].numpy()), g["entry"].numpy())


def test_vq_oracle_soft_code_matches_reference(golden):
    g = golden("vq_soft_code.pt")
    z, cb = VQ.nchw_to_rows(g["z"].numpy()), g["codebook"].numpy()
    soft, code = VQ.soft_code(z, cb, g["temp"])
    np.testing.assert_allclose(soft.reshape(2, 64, -1), g["soft"].numpy(), rtol=2e-4, atol=1e-9)
    assert np.array_equal(code.reshape(2, -1), g["code"].numpy())
    # stochastic=True: torch.multinomial(soft, 1) == argmax soft / q with q the Exp(1) draws of the same RNG stream
    torch.manual_seed(g["seed_s"])
    q = torch.empty(z.shape[0], cb.shape[0]).exponential_().numpy()
    soft_s, code_s = VQ.soft_code(z, cb, g["temp_s"], q)
    np.testing.assert_allclose(soft_s.reshape(2, 64, -1), g["soft_s"].numpy(), rtol=2e-4, atol=1e-9)
    assert np.array_equal(code_s.reshape(2, -1), g["code_s"].numpy())


def test_vq_oracle_tie_break_and_edges():
    cb = np.zeros((5, 16), dtype=np.float32)
    cb[1] = 1.0
    cb[3] = 1.0  # duplicate of row 1: the lower index must win
    z = np.ones((3, 16), dtype=np.float32)
    ids, _ = VQ.argmin(z, cb)
    assert ids.tolist() == [1, 1, 1]
    ids, _ = VQ.argmin(np.zeros((0, 16), dtype=np.float32), cb)
    assert ids.shape == (0,)


def test_micro_vqgan(golden):
    g = golden("micro_vqgan.pt")
    p, cfg = g["state_dict"], g["config"]
    with torch.no_grad():
        z = G.encoder(p, cfg, g["image"])
        _close(z, g["z"], 1e-5, 1e-6)
        zq, ids = G.quantize(p, g["z"])
        assert torch.equal(ids, g["ids"])
        _close(zq, g["z_q"], 0, 0)
        rec = G.decode_code(p, cfg, g["ids"])
        _close(rec, g["recon"], 1e-5, 1e-6)
    ids_c, _ = VQ.argmin(VQ.nchw_to_rows(g["z"].numpy()), p["quantize.embedding.weight"].numpy())
    ok = g["margin"].numpy() > 1e-4  # margin screen: fp32 re-association cannot flip these
    assert np.array_equal(ids_c[ok], g["ids"].numpy().reshape(-1)[ok])


F16_CFG = dict(resolution=256, num_channels=3, hidden_channels=128, channel_mult=(1, 1, 2, 2, 4), num_res_blocks=2,
               z_channels=256, num_embeddings=1024, quantized_embed_dim=256)


def _f16_state_dict(g):
    """The f16-256 fixture stores no weights: seeded construction reproduces the reference's default init (construction
    order == RNG order), which the stored per-tensor signatures verify."""
    from open_muse_b200.modeling_maskgit_vqgan import MaskGitVQGAN

    torch.manual_seed(g["seed"])
    m = MaskGitVQGAN()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert set(sd) == set(g["init_signature"])
    for k, (s, n) in g["init_signature"].items():
        assert abs(float(sd[k].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(float(sd[k].double().norm()) - n) <= 1e-9 * max(1.0, n), k
    sd["quantize.embedding.weight"] = g["codebook"].clone()
    return sd


def test_f16_256_vqgan_oracle_at_full_architecture(golden):
    """BASELINE config 3 at its own size: the oracle's encoder / quantiser / decoder against the outputs of the unmodified
    reference MaskGitVQGAN f16-256 (class defaults) on the committed two-image batch."""
    g = golden("f16_256_vqgan.pt")
    p = _f16_state_dict(g)
    img = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(g["image_seed"]))
    with torch.no_grad():
        z = G.encoder(p, F16_CFG, img)
        _close(z, g["z"], 1e-4, 1e-5)
        zq, ids = G.quantize(p, g["z"])
        assert torch.equal(ids, g["ids"])
        _close(zq, g["z_q"], 0, 0)
        rec = G.decode_code(p, F16_CFG, g["ids"])
        _close(rec, g["recon"], 1e-3, 1e-4)  # 32 fp32 convolutions: oneDNN re-association differs with the thread count
    ids_c, dmin = VQ.argmin(VQ.nchw_to_rows(g["z"].numpy()), p["quantize.embedding.weight"].numpy())
    ok = (g["margin"] > 1e-4 * g["dmin"].abs()).numpy()  # margin screen: fp32 re-association cannot flip these
    assert ok.mean() > 0.98
    assert np.array_equal(ids_c[ok], g["ids"].numpy().reshape(-1)[ok])


def test_public_signatures_match_reference():
    """Parameter names, order and defaults of the boundary's public calls against the reference's
    (tests/golden/signatures.json, written by make_golden.py from the imported reference)."""
    import inspect
    import json
    import os

    import open_muse_b200 as ours
    from open_muse_b200.modeling_taming_vqgan import VQGANModel
    from open_muse_b200.modeling_transformer_v2 import MaskGiTUViT_v2

    table = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "signatures.json")))
    objs = {"MaskGitTransformer": ours.MaskGitTransformer, "MaskGiTUViT_v2": MaskGiTUViT_v2, "MaskGitVQGAN": ours.MaskGitVQGAN,
            "VQGANModel": VQGANModel, "PipelineMuse": ours.PipelineMuse, "PipelineMuseInpainting": ours.PipelineMuseInpainting,
            "EMAModel": ours.EMAModel}
    private_ok = lambda n: n.startswith("_")  # our private test hooks (e.g. _raw_bf16) are keyword-only extras
    for key, ref in table.items():
        cls, meth = key.split(".")
        params = inspect.signature(getattr(objs[cls], meth)).parameters
        ours_named = {n: p for n, p in params.items() if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL) and not private_ok(n)}
        ref_named = [r for r in ref if "VAR_" not in r[1]]
        has_kwargs = any(p.kind == p.VAR_KEYWORD for p in params.values())
        for name, kind, default in ref_named:
            if name not in ours_named:
                assert has_kwargs, f"{key}: reference parameter {name!r} is neither named nor swallowed by **kwargs"
                continue
            d = ours_named[name].default
            if default == "<required>":
                assert d is inspect.Parameter.empty, (key, name)
            elif isinstance(default, str) and default.startswith("<callable"):
                assert callable(d), (key, name)
            else:
                d = list(d) if isinstance(d, tuple) else d
                assert d == default, f"{key}: default of {name!r} is {d!r}, reference has {default!r}"
        # positional order of the shared leading parameters
        ref_order = [r[0] for r in ref_named if r[0] in ours_named]
        ours_order = [n for n in ours_named if n in set(ref_order)]
        assert ours_order == ref_order, (key, ours_order, ref_order)
