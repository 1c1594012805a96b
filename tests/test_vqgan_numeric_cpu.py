"""Host logic of the two tokenizers (MaskGitVQGAN, taming VQGANModel) checked NUMERICALLY without a GPU (method:
tests/test_v1_numeric_cpu.py): the product's module wiring -- block order, GroupNorm prologues, the ResnetBlock shortcut
quirk Q8, down / up sampling, mid attention, quantiser entry points, display-byte recipe -- runs on the CPU with torch
restatements of the convolution / VQ kernel contracts (tests/cpu_math_ops.py) and must reproduce what the UNMODIFIED
reference computed on the same weights and images (tests/golden/micro_vqgan.pt, micro_taming_vqgan.pt)."""
import numpy as np
import pytest
import torch

from open_muse_b200 import MaskGitVQGAN, VQGANModel, ops
from tests import cpu_math_ops


def _rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


@pytest.mark.parametrize("cls,name", [(MaskGitVQGAN, "micro_vqgan.pt"), (VQGANModel, "micro_taming_vqgan.pt")])
def test_tokenizer_wiring_reproduces_the_reference(golden, monkeypatch, cls, name):
    g = golden(name)
    cpu_math_ops.install(monkeypatch, exact=True)
    m = cls(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.eval()
    img = g["image"]
    z = ops.to_nchw(m._encode_nhwc(img))
    assert _rel(z, g["z"]) < 2e-5
    z_q, ids = m.encode(img)
    safe = (g["margin"] > 1e-4).view(ids.shape)  # fp32 re-association cannot flip these arg-mins
    assert int(safe.sum()) > 400 and torch.equal(ids[safe], g["ids"][safe])
    assert torch.equal(m.get_code(img), ids)
    assert torch.equal(m.quantize.get_codebook_entry(g["ids"]).reshape(g["z_q"].shape), g["z_q"])  # lookup: exact values
    rec = m.decode_code(g["ids"])
    assert _rel(rec, g["recon"]) < 2e-5
    assert torch.equal(m.decode(g["z_q"]), rec)
    out = m(img)
    assert len(out) == 3 and out[0].shape == g["recon"].shape and torch.equal(out[2], ids)
    rec4, zq4, ids4, loss = m(img, return_loss=True)
    assert loss.dim() == 0 and torch.equal(ids4, ids)
    soft, code = m.get_soft_code(img, temp=2.0)
    assert soft.shape == (2, ids.shape[1], g["config"]["num_embeddings"]) and torch.equal(code[safe], g["ids"][safe])
    assert torch.allclose(soft.sum(-1), torch.ones(2, ids.shape[1]), atol=1e-5)
    # display bytes == the reference's host recipe (pipeline_muse.py:245-252) on the decoded tensor
    x = rec.permute(0, 2, 3, 1).float().numpy()
    want = (255 * ((np.clip(2.0 * x - 1.0, -1.0, 1.0) + 1.0) / 2.0)).astype(np.uint8)
    assert np.array_equal(m.decode_code_uint8(g["ids"]).numpy(), want)
