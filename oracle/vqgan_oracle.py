"""ORACLE (test infrastructure only): fp32 functional restatement of MaskGitVQGAN's encoder, quantiser and
decoder (muse/modeling_maskgit_vqgan.py), driven by a reference-named ``state_dict``.  Pinned against
tests/golden/micro_vqgan.pt (outputs of the unmodified reference)."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def conv_same(x, w, b=None):
    """Conv2dSame.forward (:33-45): stride 1 -> symmetric (k-1)/2 padding for odd k."""
    k = w.shape[-1]
    total = max(k - 1, 0)
    if total > 0:
        x = F.pad(x, [total // 2, total - total // 2, total // 2, total - total // 2])
    return F.conv2d(x, w, b)


def gn_silu(x, w, b):
    """nn.GroupNorm(32, C, eps=1e-6) followed by F.silu (:61-79)."""
    return F.silu(F.group_norm(x, 32, w, b, 1e-6))


def resnet_block(x, p, pre):
    """ResnetBlock.forward (:71-85). NB quirk Q8: with a channel change the shortcut is applied to the
    *post-conv2* activation and the block input is dropped."""
    h = conv_same(gn_silu(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"]), p[pre + "conv1.weight"])
    h = conv_same(gn_silu(h, p[pre + "norm2.weight"], p[pre + "norm2.bias"]), p[pre + "conv2.weight"])
    if pre + "nin_shortcut.weight" in p:
        return h + conv_same(h, p[pre + "nin_shortcut.weight"])
    return h + x


def encoder(p: Dict[str, torch.Tensor], cfg: dict, pixels):
    """Encoder.forward (:175-189)."""
    n_res, n_blocks = len(cfg["channel_mult"]), cfg["num_res_blocks"]
    h = conv_same(pixels, p["encoder.conv_in.weight"])
    for lvl in range(n_res):
        for b in range(n_blocks):
            h = resnet_block(h, p, f"encoder.down.{lvl}.block.{b}.")
        if lvl != n_res - 1:
            h = F.avg_pool2d(h, kernel_size=2, stride=2)  # DownsamplingBlock.forward :107-114
    for b in range(n_blocks):
        h = resnet_block(h, p, f"encoder.mid.{b}.")
    h = gn_silu(h, p["encoder.norm_out.weight"], p["encoder.norm_out.bias"])
    return conv_same(h, p["encoder.conv_out.weight"], p["encoder.conv_out.bias"])


def decoder(p, cfg, z):
    """Decoder.forward (:223-240) + UpsamplingBlock.forward (:141-149)."""
    n_res, n_blocks = len(cfg["channel_mult"]), cfg["num_res_blocks"]
    h = conv_same(z, p["decoder.conv_in.weight"], p["decoder.conv_in.bias"])
    for b in range(n_blocks):
        h = resnet_block(h, p, f"decoder.mid.{b}.")
    for lvl in reversed(range(n_res)):
        for b in range(n_blocks):
            h = resnet_block(h, p, f"decoder.up.{lvl}.block.{b}.")
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = conv_same(h, p[f"decoder.up.{lvl}.upsample_conv.weight"], p[f"decoder.up.{lvl}.upsample_conv.bias"])
    h = gn_silu(h, p["decoder.norm_out.weight"], p["decoder.norm_out.bias"])
    return conv_same(h, p["decoder.conv_out.weight"], p["decoder.conv_out.bias"])


def quantize(p, z):
    """VectorQuantizer.forward (:267-301) without the loss: returns (z_q NCHW, ids [B, T])."""
    e = p["quantize.embedding.weight"]
    b, c, hh, ww = z.shape
    flat = z.permute(0, 2, 3, 1).reshape(-1, c)
    d = torch.addmm(flat.pow(2.0).sum(1, keepdim=True) + e.t().pow(2.0).sum(0, keepdim=True), flat, e.t(), alpha=-2.0)
    ids = d.argmin(dim=1)
    z_q = e[ids].view(b, hh, ww, c).permute(0, 3, 1, 2).contiguous()
    return z_q, ids.view(b, -1)


def codebook_entry(p, ids):
    """get_codebook_entry (:318-324)."""
    b, t = ids.shape
    s = int(math.sqrt(t))
    return p["quantize.embedding.weight"][ids].reshape(b, s, s, -1).permute(0, 3, 1, 2)


def encode(p, cfg, pixels):
    return quantize(p, encoder(p, cfg, pixels))


def decode_code(p, cfg, ids):
    return decoder(p, cfg, codebook_entry(p, ids))
