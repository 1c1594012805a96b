"""ORACLE (test infrastructure only): fp32 functional restatement of the taming ``VQGANModel`` encoder / decoder
(muse/modeling_taming_vqgan.py), driven by a reference-named ``state_dict``.  Pinned against
tests/golden/micro_taming_vqgan.pt (outputs of the unmodified reference)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .vqgan_oracle import quantize  # same VectorQuantizer arithmetic (:404-510)


def _conv(x, p, pre, stride=1, padding=1):
    return F.conv2d(x, p[pre + "weight"], p.get(pre + "bias"), stride=stride, padding=padding)


def _gn(x, p, pre):
    return F.group_norm(x, 32, p[pre + "weight"], p[pre + "bias"], 1e-6)


def resnet_block(x, p, pre):
    """ResnetBlock.forward (:117-134): shortcut on the block INPUT (unlike MaskGitVQGAN's quirk)."""
    h = _conv(F.silu(_gn(x, p, pre + "norm1.")), p, pre + "conv1.")
    h = _conv(F.silu(_gn(h, p, pre + "norm2.")), p, pre + "conv2.")
    if pre + "conv_shortcut.weight" in p:
        x = _conv(x, p, pre + "conv_shortcut.")
    elif pre + "nin_shortcut.weight" in p:
        x = _conv(x, p, pre + "nin_shortcut.", padding=0)
    return h + x


def attn_block(x, p, pre):
    """AttnBlock.forward (:148-174): one head of width C over the h*w positions."""
    b, c, hh, ww = x.shape
    n = _gn(x, p, pre + "norm.")
    q = _conv(n, p, pre + "q.", padding=0).reshape(b, c, hh * ww).permute(0, 2, 1)
    k = _conv(n, p, pre + "k.", padding=0).reshape(b, c, hh * ww)
    v = _conv(n, p, pre + "v.", padding=0).reshape(b, c, hh * ww)
    w = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    o = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return _conv(o, p, pre + "proj_out.", padding=0) + x


def _level(h, p, pre, n_blocks):
    n_attn = len({k.split(".attn.")[1].split(".")[0] for k in p if k.startswith(pre + "attn.")})
    for i in range(n_blocks):
        h = resnet_block(h, p, f"{pre}block.{i}.")
        if n_attn > 1:  # `if len(self.attn) > 1` (:211, :250)
            h = attn_block(h, p, f"{pre}attn.{i}.")
    return h


def _mid(h, p, pre):
    h = resnet_block(h, p, pre + "block_1.")
    if pre + "attn_1.norm.weight" in p:
        h = attn_block(h, p, pre + "attn_1.")
    return resnet_block(h, p, pre + "block_2.")


def encoder(p: Dict[str, torch.Tensor], cfg: dict, pixels):
    """Encoder.forward (:326-340) + quant_conv (:548)."""
    n_res = len(cfg["channel_mult"])
    h = _conv(pixels, p, "encoder.conv_in.")
    for lvl in range(n_res):
        h = _level(h, p, f"encoder.down.{lvl}.", cfg["num_res_blocks"])
        if lvl != n_res - 1:  # Downsample.forward (:55-62)
            h = _conv(F.pad(h, (0, 1, 0, 1)), p, f"encoder.down.{lvl}.downsample.conv.", stride=2, padding=0)
    h = _mid(h, p, "encoder.mid.")
    h = _conv(F.silu(_gn(h, p, "encoder.norm_out.")), p, "encoder.conv_out.")
    return _conv(h, p, "quant_conv.", padding=0)


def decoder(p, cfg, z_q):
    """post_quant_conv (:558) + Decoder.forward (:385-401)."""
    n_res = len(cfg["channel_mult"])
    h = _conv(_conv(z_q, p, "post_quant_conv.", padding=0), p, "decoder.conv_in.")
    h = _mid(h, p, "decoder.mid.")
    for lvl in reversed(range(n_res)):
        h = _level(h, p, f"decoder.up.{lvl}.", cfg["num_res_blocks"] + 1)
        if lvl != 0:  # Upsample.forward (:40-44)
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), p, f"decoder.up.{lvl}.upsample.conv.")
    return _conv(F.silu(_gn(h, p, "decoder.norm_out.")), p, "decoder.conv_out.")


def encode(p, cfg, pixels):
    z = encoder(p, cfg, pixels)
    z_q, ids = quantize(p, z)
    return z, z_q, ids
