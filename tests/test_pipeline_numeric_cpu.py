"""PipelineMuse end to end WITHOUT a GPU (method: tests/test_v1_numeric_cpu.py): precomputed text states -> MaskGiTUViT_v2
generate2 with classifier-free guidance -> taming VQGAN decode -> display bytes, every kernel replaced by its torch
restatement, against the PIL images the UNMODIFIED reference pipeline produced from the same weights, inputs and generator
seed (tests/golden/micro_pipeline.pt, written by make_golden.py::make_pipeline)."""
import numpy as np
import torch

from open_muse_b200 import MaskGiTUViT_v2, PipelineMuse, VQGANModel
from tests import cpu_math_ops


def test_pipeline_reproduces_the_reference_images(golden, monkeypatch):
    g, gu, gv = golden("micro_pipeline.pt"), golden("micro_uvit_v2.pt"), golden("micro_taming_vqgan.pt")
    cpu_math_ops.install(monkeypatch, exact=True)
    monkeypatch.setenv("MUSE_B200_CUDA_GRAPH", "0")  # the decode-step CUDA graph needs a device; the loop is the same code
    monkeypatch.setattr(MaskGiTUViT_v2, "device", property(lambda self: torch.device("cpu")), raising=False)
    tr = MaskGiTUViT_v2(**gu["config"])
    tr.load_state_dict(gu["state_dict"])
    vae = VQGANModel(**gv["config"])
    vae.load_state_dict(gv["state_dict"])
    pipe = PipelineMuse(vae=vae.eval(), transformer=tr.eval())
    images = pipe(text=["a", "b"], negative_text=None, generator=torch.Generator().manual_seed(g["seed"]), use_tqdm=False,
                  **g["inputs"], **g["call"])
    assert len(images) == 2 and images[0].size == (8, 8) and images[0].mode == "RGB"
    got = np.stack([np.asarray(im) for im in images]).astype(np.int16)
    want = g["images"].numpy().astype(np.int16)
    # identical token ids (asserted below) -> decoded pixels agree to fp32 rounding; a byte may move by one where 255 * t sits
    # on an integer boundary of the truncation
    assert np.abs(got - want).max() <= 1 and (got != want).mean() < 0.02
    tokens = pipe.transformer.generate2(
        encoder_hidden_states=g["inputs"]["prompt_embeds"], cond_embeds=g["inputs"]["pooled_embeds"],
        negative_embeds=g["inputs"]["negative_prompt_embeds"], negative_cond_embeds=g["inputs"]["negative_pooled_embeds"],
        empty_embeds=None, empty_cond_embeds=None, micro_conds=torch.tensor([[256, 256, 0, 0, 6.0]]), timesteps=4,
        guidance_scale=3.0, temperature=(2, 0), generator=torch.Generator().manual_seed(g["seed"]), seq_len=16,
        use_cuda_graph=False)
    assert torch.equal(tokens, g["tokens"])
    dec = pipe(text=["a", "b"], negative_text=None, generator=torch.Generator().manual_seed(g["seed"]), output_type="pt",
               **g["inputs"], **g["call"])
    assert float((dec - g["decoded"]).norm() / g["decoded"].norm()) < 2e-5
