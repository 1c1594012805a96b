"""ORACLE (test infrastructure, never shipped or measured as the product): a plain fp32, CPU-runnable,
functional restatement of the reference's MaskGitTransformer forward / loss / generate2.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It is pinned against the fixtures in tests/golden/ that were produced by running the
unmodified reference itself (tests/golden/make_golden.py; the reference holds no golden vectors of
its own).

Each function cites the reference lines it restates (paths relative to huggingface/open-muse @ 64e1afe).
Parameters are taken from a flat ``state_dict`` (reference names), gradients come from torch autograd
over these plain ops -- the checker may use autograd; the product implements backward by hand in CUDA.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _norm(x, w, eps, norm_type):
    """muse/modeling_transformer.py:124-137 (LayerNorm, weight only) and :79-100 (RMSNorm)."""
    if norm_type == "layernorm":
        return F.layer_norm(x, (x.shape[-1],), w, None, eps)
    var = x.float().pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(var + eps) * w


def _attention(x, ctx, p, prefix, nh):
    """Attention.forward + .attention, muse/modeling_transformer.py:190-241 (no mask, no dropout)."""
    B, Sq, H = x.shape
    Skv = ctx.shape[1]
    hd = H // nh
    q = (x @ p[prefix + "query.weight"].t()).view(B, Sq, nh, hd).transpose(1, 2)
    k = (ctx @ p[prefix + "key.weight"].t()).view(B, Skv, nh, hd).transpose(1, 2)
    v = (ctx @ p[prefix + "value.weight"].t()).view(B, Skv, nh, hd).transpose(1, 2)
    scores = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    probs = scores.softmax(dim=-1)
    out = (probs @ v).transpose(1, 2).reshape(B, Sq, H)
    return out @ p[prefix + "out.weight"].t()


def _layer(x, enc, p, pre, cfg):
    """TransformerLayer.forward (:875-904) + FeedForward.forward (:785-799)."""
    eps, nt, nf, nh = cfg["layer_norm_eps"], cfg["norm_type"], cfg["use_normformer"], cfg["num_attention_heads"]
    h = _norm(x, p[pre + "attn_layer_norm.weight"], eps, nt)
    a = _attention(h, h, p, pre + "attention.", nh)
    if nf:
        a = _norm(a, p[pre + "post_attn_layer_norm.weight"], eps, nt)
    x = x + a
    if enc is not None:
        h = _norm(x, p[pre + "crossattn_layer_norm.weight"], eps, nt)
        a = _attention(h, enc, p, pre + "crossattention.", nh)
        if nf:
            a = _norm(a, p[pre + "post_crossattn_layer_norm.weight"], eps, nt)
        x = x + a
    h = _norm(x, p[pre + "ffn.pre_mlp_layer_norm.weight"], eps, "layernorm")  # always LayerNorm (:767)
    g = F.gelu(h @ p[pre + "ffn.wi_0.weight"].t()) * (h @ p[pre + "ffn.wi_1.weight"].t())
    if nf:
        g = _norm(g, p[pre + "ffn.mid_mlp_layer_norm.weight"], eps, nt)
    return x + g @ p[pre + "ffn.wo.weight"].t()


DEFAULTS = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                max_position_embeddings=256, add_cross_attention=False, norm_type="layernorm", layer_norm_eps=1e-5,
                use_normformer=True, use_encoder_layernorm=True, use_mlm_layer=True, use_mlm_layernorm=True,
                codebook_size=1024, num_vq_tokens=256, use_codebook_size_for_output=False, use_conv_in_out=False,
                patch_size=1)


def full_config(cfg: dict) -> dict:
    c = dict(DEFAULTS)
    c.update(cfg)
    c["mask_token_id"] = c["vocab_size"] - 1
    c["output_size"] = c["codebook_size"] if c["use_codebook_size_for_output"] else c["vocab_size"]
    return c


def forward(p: Dict[str, torch.Tensor], cfg: dict, input_ids, encoder_hidden_states=None, labels=None,
            label_smoothing: float = 0.0):
    """MaskGitTransformer.forward, muse/modeling_transformer.py:1224-1281.  Returns logits or (logits, loss)."""
    c = full_config(cfg)
    S = input_ids.shape[-1]
    conv = c["use_conv_in_out"]
    if conv:  # ConvEmbed.forward :1023-1041 (NCHW, PixelUnshuffle, 1x1 Conv2d, positions added on the patch grid)
        B, n, ps = input_ids.shape[0], math.isqrt(S), c["patch_size"]
        e = _norm(F.embedding(input_ids.view(B, n, n), p["embed.embeddings.weight"]), p["embed.layer_norm.weight"],
                  c["layer_norm_eps"], c["norm_type"]).permute(0, 3, 1, 2)
        if ps > 1:
            e = F.pixel_unshuffle(e, ps)
        e = F.conv2d(e, p["embed.conv.weight"])
        x = e.permute(0, 2, 3, 1).reshape(B, -1, e.shape[1])
        x = x + p["embed.position_embeddings.weight"][: x.shape[1]][None]
    else:  # Embed.forward :942-957
        x = F.embedding(input_ids, p["embed.word_embeddings.weight"]) + p["embed.position_embeddings.weight"][:S][None]
    enc = encoder_hidden_states if c["add_cross_attention"] else None
    if enc is not None and c.get("project_encoder_hidden_states", False):  # encoder_proj + norm, :1239-1241
        enc = _norm(enc @ p["encoder_proj.weight"].t(), p["encoder_proj_layer_norm.weight"], c["layer_norm_eps"],
                    c["norm_type"])
    for i in range(c["num_hidden_layers"]):
        x = _layer(x, enc, p, f"transformer_layers.{i}.", c)
    if c["use_encoder_layernorm"]:
        x = _norm(x, p["encoder_layer_norm.weight"], c["layer_norm_eps"], c["norm_type"])
    if c["use_mlm_layer"] and conv:  # ConvMlmLayer.forward :1070-1080
        m = n // ps
        h = F.conv2d(x.view(B, m, m, -1).permute(0, 3, 1, 2), p["mlm_layer.conv1.weight"])
        if ps > 1:
            h = F.pixel_shuffle(h, ps)
        h = _norm(h.permute(0, 2, 3, 1), p["mlm_layer.layer_norm.norm.weight"], c["layer_norm_eps"], c["norm_type"])
        logits = F.conv2d(h.permute(0, 3, 1, 2), p["mlm_layer.conv2.weight"]).permute(0, 2, 3, 1).reshape(B, n * n, -1)
    elif c["use_mlm_layer"]:  # MlmLayer.forward :979-985
        x = F.gelu(x @ p["mlm_layer.mlm_dense.weight"].t())
        if c["use_mlm_layernorm"]:
            x = _norm(x, p["mlm_layer.mlm_ln.weight"], c["layer_norm_eps"], c["norm_type"])
        logits = x @ p["mlm_layer.to_logits.weight"].t()
    else:
        logits = x @ p["to_logits.weight"].t()
    if labels is None:
        return logits
    loss = F.cross_entropy(logits.view(-1, c["output_size"]), labels.view(-1), ignore_index=-100,
                           label_smoothing=label_smoothing)  # :1276-1280
    return logits, loss


def forward_backward(p, cfg, input_ids, labels, encoder_hidden_states=None, label_smoothing=0.0):
    """Loss + gradients w.r.t. every parameter (autograd over the restated ops)."""
    q = {k: v.detach().clone().float().requires_grad_(True) for k, v in p.items()}
    logits, loss = forward(q, cfg, input_ids, encoder_hidden_states, labels, label_smoothing)
    loss.backward()
    return logits.detach(), loss.detach(), {k: v.grad for k, v in q.items()}


def cosine_schedule(t):
    """muse/sampling.py:38-39"""
    return torch.cos(t * math.pi * 0.5)


def generate2(p, cfg, class_ids, timesteps, temperature=1.0, generator=None, trace=None, encoder_hidden_states=None,
              negative_embeds=None, guidance_scale=0.0, input_ids=None):
    """MaskGitTransformer.generate2, muse/modeling_transformer.py:1363-1456 with mask_by_random_topk / gumbel_noise / log
    from muse/sampling.py:9-35: class-conditional (class token prepended, :1404-1407,1420-1423) or text-conditional with
    classifier-free guidance (:1395-1414: doubled batch, zeros or ``negative_embeds`` as the unconditional states,
    ``uncond + g * (cond - uncond)``).  Consumes ``generator`` exactly like the reference: one torch.multinomial and one
    uniform_ per step."""
    c = full_config(cfg)
    mask_id, L, K = c["mask_token_id"], c["num_vq_tokens"], c["codebook_size"]
    cls = None if class_ids is None else class_ids + K
    B = cls.shape[0] if cls is not None else encoder_hidden_states.shape[0]
    if input_ids is None:
        input_ids = torch.full((B, L), mask_id, dtype=torch.long)
    use_cfg = encoder_hidden_states is not None and guidance_scale > 0
    if use_cfg:
        uncond = torch.zeros_like(encoder_hidden_states) if negative_embeds is None else negative_embeds
        cond_states = torch.cat([encoder_hidden_states, uncond])
    sampled = input_ids
    for step in range(timesteps):
        model_in = input_ids if cls is None else torch.cat([cls[:, None], input_ids], dim=1)
        if use_cfg:
            cl, ul = forward(p, cfg, torch.cat([model_in] * 2), encoder_hidden_states=cond_states).chunk(2)
            logits = ul[..., :K] + guidance_scale * (cl[..., :K] - ul[..., :K])
        else:
            logits = forward(p, cfg, model_in, encoder_hidden_states=encoder_hidden_states)[..., :K]
        if cls is not None:
            logits = logits[:, 1:]
        probs = logits.softmax(dim=-1)
        sampled = torch.multinomial(probs.reshape(-1, K), 1, generator=generator)[:, 0].view(B, L)
        unknown = input_ids == mask_id
        sampled = torch.where(unknown, sampled, input_ids)
        ratio = 1.0 * (step + 1) / timesteps
        mask_ratio = cosine_schedule(torch.tensor(ratio))
        sel = probs.gather(-1, sampled[..., None]).squeeze(-1)
        sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
        mask_len = (L * mask_ratio).floor().unsqueeze(0)
        mask_len = torch.max(torch.tensor([1]), torch.min(unknown.sum(dim=-1, keepdim=True) - 1, mask_len))
        temperature = temperature * (1.0 - ratio)
        u = torch.zeros_like(sel).uniform_(0, 1, generator=generator)
        gumbel = -torch.log((-torch.log(u.clamp(min=1e-20))).clamp(min=1e-20))
        conf = torch.log(sel.clamp(min=1e-20)) + temperature * gumbel
        cut = conf.sort(dim=-1).values.gather(1, mask_len.long())
        masking = conf < cut
        if trace is not None:
            trace.append(dict(logits=logits, probs=probs, sampled=sampled.clone(), unknown=unknown, u=u, conf=conf,
                              mask_len=mask_len.clone(), temperature=temperature, masking=masking))
        input_ids = torch.where(masking, mask_id, sampled)
    return sampled


def sample_step(probs, input_ids, mask_id, q_exp, u, mask_len, temperature):
    """One generate2 step with PRE-DRAWN noise (what the fused CUDA kernel implements):
    ``q_exp`` ~ Exp(1) [B,L,K] (torch.multinomial(p,1) == argmax(p / q), ATen's multinomial recipe),
    ``u`` ~ U(0,1) [B,L].  Returns (sampled_ids, next_input_ids)."""
    sampled = (probs / q_exp).argmax(dim=-1)
    unknown = input_ids == mask_id
    sampled = torch.where(unknown, sampled, input_ids)
    sel = probs.gather(-1, sampled[..., None]).squeeze(-1)
    sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
    gumbel = -torch.log((-torch.log(u.clamp(min=1e-20))).clamp(min=1e-20))
    conf = torch.log(sel.clamp(min=1e-20)) + temperature * gumbel
    cut = conf.sort(dim=-1).values.gather(1, mask_len.long())
    masking = conf < cut
    return sampled, torch.where(masking, mask_id, sampled)


def mask_tokens(tokens, class_ids, timesteps, rand, codebook_size, mask_id, min_masking_rate=0.0):
    """Training-time masking recipe, training/train_maskgit_imagenet.py:375-393, with the two random draws
    (``timesteps`` = rand(B), ``rand`` = rand(B,S)) passed in."""
    B, S = tokens.shape
    mask_prob = cosine_schedule(timesteps).clip(min_masking_rate)
    n_mask = (S * mask_prob).round().clamp(min=1)
    perm = rand.argsort(dim=-1)
    mask = perm < n_mask.unsqueeze(-1)  # quirk Q7: compares the argsort *indices*
    input_ids = torch.where(mask, mask_id, tokens)
    labels = torch.where(mask, tokens, -100)
    input_ids = torch.cat([(class_ids + codebook_size).unsqueeze(-1), input_ids], dim=-1)
    labels = torch.cat([torch.full((B, 1), -100, dtype=labels.dtype), labels], dim=-1)
    return input_ids, labels


def generate2_scalars(cfg, timesteps, temperature=1.0, seq_len=None):
    """The host-side scalars of the reference loop (:1443-1451, sampling.py:38-39), per step: (mask_len before the
    per-row clamp, temperature AFTER the compounding update of quirk Q4 -- the value the step's gumbel term uses)."""
    c = full_config(cfg)
    L = c["num_vq_tokens"] if seq_len is None else seq_len
    out = []
    for step in range(timesteps):
        ratio = 1.0 * (step + 1) / timesteps
        mask_len = (L * cosine_schedule(torch.tensor(ratio))).floor()
        temperature = temperature * (1.0 - ratio)
        out.append((mask_len, temperature))
    return out


def generate2_step_teacher_forced(p, cfg, model_input_ids, input_ids, step, timesteps, scalars, q_exp, u, logits=None,
                                  encoder_hidden_states=None):
    """ONE iteration of the reference loop (:1397-1454) from a GIVEN state, with pre-drawn noise: forward (or the given
    ``logits`` [B, L, K]) -> softmax -> categorical draw as argmax(p / q_exp) (ATen's multinomial(n=1) recipe, proven
    equal to the generator path in tests/test_oracle_golden.py) -> keep known tokens -> confidence -> cut -> re-mask.
    Returns a dict with the sampled ids, the next input ids and the quantities a margin screen needs."""
    c = full_config(cfg)
    mask_id, K = c["mask_token_id"], c["codebook_size"]
    L = input_ids.shape[1]
    if logits is None:
        full = forward(p, cfg, model_input_ids, encoder_hidden_states=encoder_hidden_states)[..., :K]
        logits = full[:, model_input_ids.shape[1] - L:]
    probs = logits.float().softmax(dim=-1)
    mask_len, temperature = scalars[step]
    unknown = input_ids == mask_id
    ml = torch.max(torch.ones(1, dtype=torch.float32, device=probs.device),
                   torch.min(unknown.sum(dim=-1, keepdim=True) - 1, mask_len.to(probs.device)))
    sampled, nxt = sample_step(probs, input_ids, mask_id, q_exp, u, ml, temperature)
    score = (probs / q_exp).topk(2, dim=-1).values
    sel = probs.gather(-1, sampled[..., None]).squeeze(-1)
    sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
    gumbel = -torch.log((-torch.log(u.clamp(min=1e-20))).clamp(min=1e-20))
    conf = torch.log(sel.clamp(min=1e-20)) + temperature * gumbel
    cut = conf.sort(dim=-1).values.gather(1, ml.long())
    return dict(sampled=sampled, next_ids=nxt, probs=probs, top2=score, conf=conf, cut=cut, mask_len=ml, unknown=unknown)
