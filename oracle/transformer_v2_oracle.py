"""ORACLE (test infrastructure only -- never imported by the product path): plain fp32 functional restatement of
``MaskGiTUViT_v2`` (muse/modeling_transformer_v2.py) on a state_dict, pinned against outputs of the unmodified reference
(tests/golden/micro_uvit_v2.pt, tests/test_oracle_golden.py).

Layout note: the reference shuffles between NCHW and token-major; here every activation is token-major [B, h*w, C]
(h = w = sqrt(S)), which is the same data.  ``stages`` (optional dict) collects the activations at the block boundaries so a
mismatch in the CUDA path can be localised."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

DEFAULTS = dict(hidden_size=1024, use_bias=False, hidden_dropout=0.0, cond_embed_dim=768, micro_cond_encode_dim=256,
                micro_cond_embed_dim=1280, encoder_hidden_size=768, vocab_size=8256, mask_token_id=8255,
                codebook_size=8192, in_channels=768, block_out_channels=(768,), num_res_blocks=3,
                force_down_up_sample=False, block_num_heads=12, num_hidden_layers=22, num_attention_heads=16,
                attention_dropout=0.0, intermediate_size=2816, use_fused_mlp=False, norm_type="rmsnorm",
                layer_norm_eps=1e-6, ln_elementwise_affine=True, use_fused_residual_norm=False, add_cond_embeds=True,
                add_micro_cond_embeds=True)  # MaskGiTUViT_v2Config, :79-123


def full_config(cfg: dict) -> dict:
    c = dict(DEFAULTS)
    c.update({k: v for k, v in cfg.items() if k in DEFAULTS})  # config_from_legacy_kwargs drops unknown keys (:126-147)
    if isinstance(c["block_num_heads"], (tuple, list)):
        c["block_num_heads"] = c["block_num_heads"][0]
    return c


def sinusoidal_encode(features, dim, max_positions=10000):
    """:59-76: [cos(f w_k), sin(f w_k)], w_k = max_positions^(-k / (dim/2))."""
    half = dim // 2
    w = torch.exp(torch.arange(half, dtype=torch.float32, device=features.device) * (-math.log(max_positions) / half))
    e = features[:, None].float() * w[None, :]
    return torch.cat([e.cos(), e.sin()], dim=1)


def _norm(x, w, c, kind=None, residual=None):
    """Norm (:647-738): returns (normed, prenorm_residual); the residual argument is added first."""
    if residual is not None:
        x = x + residual
    pre = x
    if (kind or c["norm_type"]) == "layernorm":
        y = F.layer_norm(x, (x.shape[-1],), w, None, c["layer_norm_eps"])
    else:
        y = x * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + c["layer_norm_eps"])
        if w is not None:
            y = y * w
    return y, pre


def _attention(x, ctx, p, pre, nh):
    """Attention.forward (:853-916): q from x, k/v from ctx, softmax(q k^T / sqrt(hd)) v, out projection."""
    B, S, H = x.shape
    hd = H // nh
    q = (x @ p[pre + "query.weight"].t()).view(B, S, nh, hd).transpose(1, 2)
    k = (ctx @ p[pre + "key.weight"].t()).view(B, -1, nh, hd).transpose(1, 2)
    v = (ctx @ p[pre + "value.weight"].t()).view(B, -1, nh, hd).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1) @ v
    return a.transpose(1, 2).reshape(B, S, H) @ p[pre + "out.weight"].t()


def _adaln(x, cond, p, pre):
    """AdaLNModulation (:1025-1037): x * (1 + scale) + shift, (scale, shift) = Linear(SiLU(cond)) broadcast over tokens."""
    scale, shift = (F.silu(cond) @ p[pre + "mapper.weight"].t()).chunk(2, dim=1)
    return x * (1 + scale[:, None]) + shift[:, None]


def _res_block(x, cond, p, pre, c, hw):
    """ResBlock (:586-618) on token-major x [B, h*w, C]."""
    B, S, C = x.shape
    img = x.view(B, hw, hw, C).permute(0, 3, 1, 2)
    d = F.conv2d(img, p[pre + "depthwise.weight"], None, padding=1, groups=C).permute(0, 2, 3, 1)  # [B,h,w,C]
    d, _ = _norm(d, p.get(pre + "norm.norm.weight"), c)
    g = F.gelu(d @ p[pre + "channelwise.0.weight"].t())
    gx = torch.norm(g, p=2, dim=(1, 2), keepdim=True)  # GlobalResponseNorm (:741-751)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    g = p[pre + "channelwise.2.gamma"] * (g * nx) + p[pre + "channelwise.2.beta"] + g
    y = (g @ p[pre + "channelwise.4.weight"].t()).reshape(B, S, C) + x
    return _adaln(y, cond, p, pre + "adaLN_modulation.")


def _attention_block(x, enc, p, pre, c):
    """AttentionBlock2D (:795-831): two cross-attentions to the text states with the prenorm-residual carry."""
    if pre + "kv_mapper.weight" in p:
        enc = F.silu(enc) @ p[pre + "kv_mapper.weight"].t()
    nh = c["block_num_heads"]
    h, r = _norm(x, p.get(pre + "attn_layer_norm.weight"), c)
    h = _attention(h, enc, p, pre + "attention.", nh)
    h, r = _norm(h, p.get(pre + "crossattn_layer_norm.weight"), c, residual=r)
    h = _attention(h, enc, p, pre + "crossattention.", nh)
    return h + r


def _layer(x, r, enc, cond, p, pre, c):
    """TransformerLayer (:757-792) + GLUFeedForward (:926-951), carrying (hidden, residual)."""
    nh = c["num_attention_heads"]
    h, r = _norm(x, p.get(pre + "attn_layer_norm.weight"), c, residual=r)
    h = _adaln(h, cond, p, pre + "self_attn_adaLN_modulation.")
    h = _attention(h, h, p, pre + "attention.", nh)
    h, r = _norm(h, p.get(pre + "crossattn_layer_norm.weight"), c, residual=r)
    h = _adaln(h, cond, p, pre + "cross_attn_adaLN_modulation.")
    h = _attention(h, enc, p, pre + "crossattention.", nh)
    h, r = _norm(h, p.get(pre + "ffn.pre_mlp_layer_norm.weight"), c, kind="layernorm", residual=r)  # always LayerNorm (:929)
    h = _adaln(h, cond, p, pre + "ffn.adaLN_modulation.")
    h = (F.gelu(h @ p[pre + "ffn.wi_0.weight"].t()) * (h @ p[pre + "ffn.wi_1.weight"].t())) @ p[pre + "ffn.wo.weight"].t()
    return h, r


def forward(p: Dict[str, torch.Tensor], cfg: dict, input_ids, encoder_hidden_states, cond_embeds, micro_conds,
            labels=None, label_smoothing: float = 0.0, loss_weight=None, stages: Optional[dict] = None):
    """MaskGiTUViT_v2.forward (:242-319).  Returns logits [B, S, codebook_size] or (logits, loss)."""
    c = full_config(cfg)
    if c["use_bias"] or c["use_fused_mlp"]:
        raise NotImplementedError("oracle restates the bias-free GLU wiring only")
    B, S = input_ids.shape
    hw = int(S ** 0.5)
    C = c["block_out_channels"][0]
    st = stages if stages is not None else {}
    enc, _ = _norm(encoder_hidden_states @ p["encoder_proj.weight"].t(), p.get("encoder_proj_layer_norm.weight"), c)
    mc = sinusoidal_encode(micro_conds.flatten(), c["micro_cond_encode_dim"]).reshape(B, -1)
    cond = torch.cat([cond_embeds, mc], dim=1)
    cond = F.silu(cond @ p["cond_embed.0.weight"].t()) @ p["cond_embed.2.weight"].t()
    st["enc"], st["cond"] = enc, cond
    # ConvEmbed (:485-500): embedding -> Norm -> 1x1 conv
    e, _ = _norm(F.embedding(input_ids, p["embed.embeddings.weight"]), p.get("embed.layer_norm.weight"), c)
    x = e @ p["embed.conv.weight"][:, :, 0, 0].t()
    st["embed"] = x
    if c["force_down_up_sample"]:  # DownsampleBlock.downsample (:509-513): Norm2D -> Conv2d(k=2, s=2), tokens hw^2 -> (hw/2)^2
        x, _ = _norm(x, p.get("down_blocks.0.downsample.0.norm.weight"), c)
        x = F.conv2d(x.view(B, hw, hw, C).permute(0, 3, 1, 2), p["down_blocks.0.downsample.1.weight"], stride=2)
        st["downsample"] = x
        hw //= 2
        x = x.permute(0, 2, 3, 1).reshape(B, hw * hw, C)
    for i in range(c["num_res_blocks"]):
        x = _res_block(x, cond, p, f"down_blocks.0.res_blocks.{i}.", c, hw)
        x = _attention_block(x, enc, p, f"down_blocks.0.attention_blocks.{i}.", c)
    st["down"] = x
    x, _ = _norm(x, p.get("project_to_hidden_norm.weight"), c)
    x = x @ p["project_to_hidden.weight"].t()
    st["hidden0"] = x
    r = None
    for i in range(c["num_hidden_layers"]):
        x, r = _layer(x, r, enc, cond, p, f"transformer_layers.{i}.", c)
        st[f"layer{i}"] = x + r
    x = x + r
    x, _ = _norm(x, p.get("project_from_hidden_norm.weight"), c)
    x = x @ p["project_from_hidden.weight"].t()
    st["from_hidden"] = x
    for i in range(c["num_res_blocks"]):
        x = _res_block(x, cond, p, f"up_blocks.0.res_blocks.{i}.", c, hw)
        x = _attention_block(x, enc, p, f"up_blocks.0.attention_blocks.{i}.", c)
    st["up"] = x
    if c["force_down_up_sample"]:  # UpsampleBlock.upsample (:555-559): Norm2D -> ConvTranspose2d(k=2, s=2)
        x, _ = _norm(x, p.get("up_blocks.0.upsample.0.norm.weight"), c)
        x = F.conv_transpose2d(x.view(B, hw, hw, C).permute(0, 3, 1, 2), p["up_blocks.0.upsample.1.weight"], stride=2)
        st["upsample"] = x
        hw *= 2
        x = x.permute(0, 2, 3, 1).reshape(B, hw * hw, C)
    # ConvMlmLayer (:1002-1022): 1x1 conv -> Norm2D -> 1x1 conv
    y = x @ p["mlm_layer.conv1.weight"][:, :, 0, 0].t()
    y, _ = _norm(y, p.get("mlm_layer.layer_norm.norm.weight"), c)
    logits = y @ p["mlm_layer.conv2.weight"][:, :, 0, 0].t()
    if labels is None:
        return logits
    if loss_weight is not None:  # (:305-317)
        loss = F.cross_entropy(logits.view(-1, c["codebook_size"]), labels.view(-1), ignore_index=-100,
                               label_smoothing=label_smoothing, reduction="none")
        lw = loss_weight.view(-1)
        loss = ((loss * lw).sum(dim=-1) / lw.sum(dim=-1)).mean()
    else:
        loss = F.cross_entropy(logits.view(-1, c["codebook_size"]), labels.view(-1), ignore_index=-100,
                               label_smoothing=label_smoothing)
    return logits, loss


def cosine_schedule(t):
    return torch.cos(t * math.pi * 0.5)


def generate2(p, cfg, encoder_hidden_states, cond_embeds, micro_conds, empty_embeds, empty_cond_embeds, timesteps,
              temperature=1.0, guidance_scale=0.0, generator=None, seq_len=16):
    """MaskGiTUViT_v2.generate2 (:330-479) with the constant guidance schedule: consumes `generator` exactly like the
    reference (torch.multinomial for the categorical draw, one uniform per token for the gumbel noise)."""
    c = full_config(cfg)
    B = encoder_hidden_states.shape[0]
    mask_id, K = c["vocab_size"] - 1, c["codebook_size"]
    temps = torch.linspace(temperature[0], temperature[1], timesteps) if isinstance(temperature, tuple) else \
        torch.linspace(temperature, 0.01, timesteps)
    input_ids = torch.full((B, seq_len), mask_id, dtype=torch.long)
    if micro_conds.shape[0] == 1:
        micro_conds = micro_conds.repeat(B, 1)
    if guidance_scale > 0:
        unc_e = empty_embeds.expand(B, -1, -1) if empty_embeds.shape[0] == 1 else empty_embeds
        unc_c = empty_cond_embeds.expand(B, -1) if empty_cond_embeds.shape[0] == 1 else empty_cond_embeds
        encoder_hidden_states = torch.cat([encoder_hidden_states, unc_e])
        cond_embeds = torch.cat([cond_embeds, unc_c])
        micro_conds = torch.cat([micro_conds, micro_conds])
    sampled_ids = None
    for step in range(timesteps):
        model_in = torch.cat([input_ids] * 2) if guidance_scale > 0 else input_ids
        out = forward(p, cfg, model_in, encoder_hidden_states, cond_embeds, micro_conds)
        if guidance_scale > 0:
            cl, ul = out.chunk(2)
            logits = ul[..., :K] + guidance_scale * (cl[..., :K] - ul[..., :K])
        else:
            logits = out[..., :K]
        probs = logits.softmax(dim=-1)
        sampled_ids = torch.multinomial(probs.reshape(-1, K), 1, generator=generator)[:, 0].view(B, seq_len)
        unknown = input_ids == mask_id
        sampled_ids = torch.where(unknown, sampled_ids, input_ids)
        ratio = 1.0 * (step + 1) / timesteps
        mask_len = (seq_len * cosine_schedule(torch.tensor(ratio))).floor().unsqueeze(0)
        mask_len = torch.max(torch.tensor([1]), torch.min(unknown.sum(dim=-1, keepdim=True) - 1, mask_len))
        sel = torch.gather(probs, -1, sampled_ids[..., None]).squeeze(-1)
        sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
        u = torch.zeros_like(sel).uniform_(0, 1, generator=generator)
        gumbel = -torch.log((-torch.log(u.clamp(1e-20))).clamp(1e-20))  # sampling.py:9-15
        conf = torch.log(sel.clamp(1e-20)) + temps[step] * gumbel
        cut = torch.gather(torch.sort(conf, dim=-1).values, -1, mask_len.long())
        input_ids = torch.where(conf < cut, mask_id, sampled_ids)
    return sampled_ids
