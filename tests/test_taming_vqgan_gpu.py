"""GPU parity of the taming VQGANModel path: strided Downsample convolution, single-head AttnBlock attention, plain
GroupNorm, the micro model vs the reference fixture and the f16 architecture on both convolution routes."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from open_muse_b200 import ops  # noqa: E402
from open_muse_b200.modeling_taming_vqgan import VQGANModel  # noqa: E402
from oracle import vq_oracle as VQ  # noqa: E402

DEV = "cuda"


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 16, 16, 128, 128), (1, 8, 16, 64, 256), (2, 16, 16, 32, 48), (1, 4, 128, 16, 32)])
def test_downsample_conv_vs_fp64(B, H, W, cin, cout, monkeypatch):
    """pad (0,1,0,1) + 3x3 stride 2 (taming Downsample): space-to-depth tensor-core route and fp32 SIMT route."""
    g = torch.Generator().manual_seed(H * W + cin)
    x = torch.randn(B, cin, 2 * H, 2 * W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    b = torch.randn(cout, generator=g)
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x.double(), (0, 1, 0, 1)), w.double(), b.double(), stride=2)
    for route in ("tc", "simt"):
        with ops.conv_route(None if route == "tc" else "simt"):
            y = ops.to_nchw(ops.conv2d_down(ops.to_nhwc(x.to(DEV)), w.to(DEV), b.to(DEV)))
        assert y.shape == ref.shape and _rel(y, ref) < 2e-5, (route, _rel(y, ref))


@pytest.mark.parametrize("B,hw,C", [(3, 16, 512), (2, 16, 64), (2, 4, 32)])
def test_single_head_attention_and_plain_groupnorm(B, hw, C, monkeypatch):
    g = torch.Generator().manual_seed(C)
    HW = hw * hw
    q, k, v = (torch.randn(B, HW, C, generator=g) for _ in range(3))
    w = torch.softmax(torch.bmm(q.double(), k.double().transpose(1, 2)) * C ** -0.5, dim=2)
    ref = torch.bmm(w, v.double())
    for route in ("tc", "simt"):
        with ops.conv_route(None if route == "tc" else "simt"):
            o = ops.attention_single_head(q.view(-1, C).to(DEV), k.view(-1, C).to(DEV), v.view(-1, C).to(DEV), B, hw, hw)
        assert _rel(o.view(B, HW, C), ref) < 3e-5, (route, _rel(o.view(B, HW, C), ref))
    x = torch.randn(B, C, hw, hw, generator=g)
    ga, be = torch.randn(C, generator=g), torch.randn(C, generator=g)
    n = ops.groupnorm_silu(ops.to_nhwc(x.to(DEV)), ga.to(DEV), be.to(DEV), 32, 1e-6, silu=0)
    assert _rel(ops.to_nchw(n), torch.nn.functional.group_norm(x.double(), 32, ga.double(), be.double(), 1e-6)) < 1e-5


def test_micro_taming_vqgan_vs_reference(golden):
    g = golden("micro_taming_vqgan.pt")
    m = VQGANModel(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.to(DEV).eval()
    img = g["image"].to(DEV)
    z = ops.to_nchw(m._encode_nhwc(img))
    assert _rel(z, g["z"]) < 1e-4, _rel(z, g["z"])
    z_q, ids = m.encode(img)
    safe = (g["margin"] > 1e-3).view(2, -1)
    assert int(safe.sum()) > 400 and torch.equal(ids.cpu()[safe], g["ids"][safe])
    ids_o, _ = VQ.argmin(VQ.nchw_to_rows(z.cpu().numpy()), g["state_dict"]["quantize.embedding.weight"].numpy())
    assert np.array_equal(ids.cpu().numpy().reshape(-1), ids_o)  # the search itself is bit-exact on our own encoder output
    rec = m.decode_code(g["ids"].to(DEV))
    assert _rel(rec, g["recon"]) < 1e-4, _rel(rec, g["recon"])
    assert torch.equal(m.decode(g["z_q"].to(DEV)), rec)
    out = m(img)
    assert len(out) == 3 and out[0].shape == g["recon"].shape
    assert torch.equal(m.get_code(img), ids)
    # display bytes on the device == the reference's host recipe (pipeline_muse.py:245-252) on the decoded tensor
    x = rec.permute(0, 2, 3, 1).float().cpu().numpy()
    want = (255 * ((np.clip(2.0 * x - 1.0, -1.0, 1.0) + 1.0) / 2.0)).astype(np.uint8)
    assert np.array_equal(m.decode_code_uint8(g["ids"].to(DEV)).cpu().numpy(), want)


def test_taming_f16_tensor_core_route_matches_fp32_simt_route(monkeypatch):
    """Default architecture (256 px -> 16x16, attention at 16x16 and in the mid blocks, 512-wide single head): both
    convolution routes agree to fp32-level tolerance on the encoder output and the decoded pixels."""
    torch.manual_seed(3)
    m = VQGANModel(num_embeddings=8192).to(DEV).eval()
    img = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(4)).to(DEV)
    outs = {}
    for route in ("simt", "tc"):
        with ops.conv_route(None if route == "tc" else "simt"):
            outs[route] = m._encode_nhwc(img)
    z_s, z_t = outs["simt"], outs["tc"]
    assert z_t.shape == (2, 16, 16, 256) and _rel(z_t, z_s) < 1e-4, _rel(z_t, z_s)
    ids = torch.randint(0, 8192, (2, 256), generator=torch.Generator().manual_seed(5)).to(DEV)
    rec_t = m.decode_code(ids)
    with ops.conv_route("simt"):
        rec_s = m.decode_code(ids)
    assert rec_t.shape == (2, 3, 256, 256) and _rel(rec_t, rec_s) < 2e-4, _rel(rec_t, rec_s)
