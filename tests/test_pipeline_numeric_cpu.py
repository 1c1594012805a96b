"""PipelineMuse end to end WITHOUT a GPU (method: tests/test_v1_numeric_cpu.py): precomputed text states -> MaskGiTUViT_v2
generate2 with classifier-free guidance -> taming VQGAN decode -> display bytes, every kernel replaced by its torch
restatement, against the PIL images the UNMODIFIED reference pipeline produced from the same weights, inputs and generator
seed (tests/golden/micro_pipeline.pt, written by make_golden.py::make_pipeline)."""
import os

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


def _v1_models(golden, monkeypatch):
    from open_muse_b200 import MaskGitTransformer, MaskGitVQGAN

    gt, gv = golden("micro_transformer.pt"), golden("micro_vqgan.pt")
    cpu_math_ops.install(monkeypatch, exact=True)
    monkeypatch.setattr(MaskGitTransformer, "device", property(lambda self: torch.device("cpu")), raising=False)
    monkeypatch.setenv("MUSE_B200_GENERATE_GRAPH", "0")
    tr = MaskGitTransformer(**gt["config"])
    tr.load_state_dict(gt["state_dict"])
    vae = MaskGitVQGAN(**gv["config"])
    vae.load_state_dict(gv["state_dict"])
    return gt, gv, tr.eval(), vae.eval()


def _bytes(decoded):
    x = decoded.permute(0, 2, 3, 1).float().numpy()
    return (255 * ((np.clip(2.0 * x - 1.0, -1.0, 1.0) + 1.0) / 2.0)).astype(np.uint8).astype(np.int16)


def test_class_conditional_pipeline_matches_the_oracle_composition(golden, monkeypatch):
    """PipelineMuse(class_ids=...) with a v1 MaskGitTransformer + MaskGitVQGAN (a superset of upstream, whose pipeline
    reads U-ViT-only config keys, quirk Q15): class ids -> generate2 -> decode_code -> display bytes, against the oracle's
    generate2 (pinned to the reference's id traces) followed by the oracle decoder."""
    from oracle import transformer_oracle as T
    from oracle import vqgan_oracle as G

    gt, gv, tr, vae = _v1_models(golden, monkeypatch)
    pipe = PipelineMuse(vae=vae, transformer=tr, is_class_conditioned=True)
    images = pipe(class_ids=[1, 5], timesteps=4, temperature=1.0, num_images_per_prompt=2,
                  generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        ids = T.generate2(gt["state_dict"], gt["config"], torch.tensor([1, 1, 5, 5]), 4, 1.0, torch.Generator().manual_seed(7))
        want = _bytes(G.decode_code(gv["state_dict"], gv["config"], ids))
    got = np.stack([np.asarray(im) for im in images]).astype(np.int16)
    assert got.shape == want.shape == (4, 8, 8, 3)
    assert np.abs(got - want).max() <= 1 and (got != want).mean() < 0.02


def test_inpainting_pipeline_matches_the_oracle_composition(golden, monkeypatch):
    """PipelineMuseInpainting (pipeline_muse.py:372-512): tokenise, overwrite the masked positions with the mask id,
    generate2 from those start tokens (known tokens kept), decode -- against encoder / quantiser / generate2 / decoder of the
    oracles chained the same way."""
    from open_muse_b200 import PipelineMuseInpainting
    from oracle import transformer_oracle as T
    from oracle import vqgan_oracle as G

    gt, gv, tr, vae = _v1_models(golden, monkeypatch)
    pipe = PipelineMuseInpainting(vae=vae, transformer=tr, is_class_conditioned=True)
    image = torch.rand(1, 3, 8, 8, generator=torch.Generator().manual_seed(5))
    mask = torch.zeros(16, dtype=torch.bool)
    mask[4:12] = True
    out = pipe(image, mask, class_ids=[2], timesteps=3, num_images_per_prompt=2, temperature=1.0,
               generator=torch.Generator().manual_seed(9), output_type="pt")
    with torch.no_grad():
        _, ids0 = G.quantize(gv["state_dict"], G.encoder(gv["state_dict"], gv["config"], image))
        start = ids0.reshape(1, 16).clone()
        start[:, mask] = gt["config"]["vocab_size"] - 1
        ids = T.generate2(gt["state_dict"], gt["config"], torch.tensor([2, 2]), 3, 1.0, torch.Generator().manual_seed(9),
                          input_ids=start.repeat(2, 1))
        want = G.decode_code(gv["state_dict"], gv["config"], ids)
    assert torch.equal(ids[:, ~mask], ids0.reshape(1, 16)[:, ~mask].expand(2, -1))  # known tokens are never resampled
    assert out.shape == want.shape == (2, 3, 8, 8)
    assert float((out - want).norm() / want.norm()) < 2e-5


def test_pipeline_runs_the_text_encoder_like_the_reference(golden, monkeypatch, tmp_path):
    """``pipe(text=[...])`` with a text encoder and tokenizer ATTACHED (the reference's main entry, pipeline_muse.py:113-197):
    tokenisation, penultimate-layer states + projected pooled embedding, the encoded negative prompt (default "") or the
    encoded empty prompt (``negative_text=None``), ``clip_skip``, micro-conditioning -- every tensor the pipeline hands to
    generate2 equals what the unmodified reference pipeline handed over with the same (tiny, seeded, real ``transformers``)
    CLIP encoder; token ids equal, images within one LSB."""
    from transformers import CLIPTextModelWithProjection, CLIPTokenizer

    from tests.train_script_harness import make_tiny_clip

    g, gu, gv = golden("micro_pipeline_text.pt"), golden("micro_uvit_v2.pt"), golden("micro_taming_vqgan.pt")
    real_bf16 = torch.bfloat16
    clip_dir = make_tiny_clip(str(tmp_path / "clip"), projection_dim=gu["config"]["cond_embed_dim"], weight_std=0.3)
    clip = CLIPTextModelWithProjection.from_pretrained(clip_dir).eval()
    tok = CLIPTokenizer.from_pretrained(clip_dir)
    assert abs(float(sum(v.double().abs().sum() for v in clip.state_dict().values())) - g["clip_signature"]) < 1e-6 * g["clip_signature"]
    cpu_math_ops.install(monkeypatch, exact=True)
    monkeypatch.setenv("MUSE_B200_CUDA_GRAPH", "0")
    monkeypatch.setattr(MaskGiTUViT_v2, "device", property(lambda self: torch.device("cpu")), raising=False)
    tr = MaskGiTUViT_v2(**gu["config"])
    tr.load_state_dict(gu["state_dict"])
    vae = VQGANModel(**gv["config"])
    vae.load_state_dict(gv["state_dict"])
    pipe = PipelineMuse(vae=vae.eval(), transformer=tr.eval(), text_encoder=clip, tokenizer=tok)
    fed, gen2 = [], tr.generate2

    def spy(**kw):
        fed.append({k: v for k, v in kw.items() if torch.is_tensor(v)})
        return gen2(**kw)

    monkeypatch.setattr(tr, "generate2", spy, raising=False)
    for name, kw in (("negative_default", {}), ("negative_none", dict(negative_text=None)), ("clip_skip", dict(clip_skip=2))):
        ref = g["runs"][name]
        images = pipe(text=g["text"], generator=torch.Generator().manual_seed(g["seed"]), use_tqdm=False, **g["call"], **kw)
        got = fed[-1]
        assert set(ref["fed"]) <= set(got), (name, set(ref["fed"]) - set(got))
        for k, v in ref["fed"].items():
            assert got[k].shape == v.shape, (name, k, got[k].shape, v.shape)
            torch.testing.assert_close(got[k].float(), v.float(), rtol=1e-5, atol=1e-6, msg=lambda m: f"{name} {k}: {m}")
        a = np.stack([np.asarray(im) for im in images]).astype(np.int16)
        b = ref["images"].numpy().astype(np.int16)
        assert np.abs(a - b).max() <= 1 and (a != b).mean() < 0.02, name
    assert not torch.equal(g["runs"]["clip_skip"]["fed"]["encoder_hidden_states"],
                           g["runs"]["negative_default"]["fed"]["encoder_hidden_states"])
    # <dir>/{text_encoder, vae, transformer} round trip: the text encoder comes back as the projection class, the tokenizer
    # is loaded with it, and the reloaded pipeline computes the same images
    pipe.save_pretrained(str(tmp_path / "pipe"))
    assert sorted(os.listdir(tmp_path / "pipe")) == ["text_encoder", "transformer", "vae"]
    pipe2 = PipelineMuse.from_pretrained(str(tmp_path / "pipe"))
    assert type(pipe2.text_encoder).__name__ == "CLIPTextModelWithProjection" and pipe2.tokenizer is not None
    assert isinstance(pipe2.transformer, MaskGiTUViT_v2) and isinstance(pipe2.vae, VQGANModel)
    pipe2.text_encoder.eval()
    images2 = pipe2(text=g["text"], generator=torch.Generator().manual_seed(g["seed"]), use_tqdm=False, **g["call"], clip_skip=2)
    assert all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(images, images2))
    assert torch.bfloat16 is not real_bf16  # (exact mode was active for the product code above)


def test_inpainting_pipeline_with_text_encoder_matches_the_reference(golden, monkeypatch, tmp_path):
    """PipelineMuseInpainting with MaskGiTUViT_v2, text + attached encoder (pipeline_muse.py:372-512): PIL -> resize / centre
    crop / ToTensor -> tokens -> masked start ids -> generate2 (start tokens of their own length, default seq_len for the
    mask schedule) -> decode.  Every tensor handed to generate2 (start ids included), the generated ids and the images against
    the unmodified reference pipeline (fixture micro_inpainting_text.pt)."""
    from PIL import Image
    from transformers import CLIPTextModelWithProjection, CLIPTokenizer

    from open_muse_b200 import PipelineMuseInpainting
    from tests.train_script_harness import make_tiny_clip

    g, gu, gv = golden("micro_inpainting_text.pt"), golden("micro_uvit_v2.pt"), golden("micro_taming_vqgan.pt")
    clip_dir = make_tiny_clip(str(tmp_path / "clip"), projection_dim=gu["config"]["cond_embed_dim"], weight_std=0.3)
    clip = CLIPTextModelWithProjection.from_pretrained(clip_dir).eval()
    tok = CLIPTokenizer.from_pretrained(clip_dir)
    cpu_math_ops.install(monkeypatch, exact=True)
    monkeypatch.setenv("MUSE_B200_CUDA_GRAPH", "0")
    monkeypatch.setattr(MaskGiTUViT_v2, "device", property(lambda self: torch.device("cpu")), raising=False)
    tr = MaskGiTUViT_v2(**gu["config"])
    tr.load_state_dict(gu["state_dict"])
    vae = VQGANModel(**gv["config"])
    vae.load_state_dict(gv["state_dict"])
    pipe = PipelineMuseInpainting(vae=vae.eval(), transformer=tr.eval(), text_encoder=clip, tokenizer=tok)
    fed, gen2 = [], tr.generate2

    def spy(**kw):
        fed.append({k: v.clone() for k, v in kw.items() if torch.is_tensor(v)})
        return gen2(**kw)

    monkeypatch.setattr(tr, "generate2", spy, raising=False)
    image = Image.fromarray(g["pixels"].numpy())
    for name, kw in (("negative_text", dict(negative_text="dog")), ("no_negative", {})):
        ref = g["runs"][name]
        images = pipe(image, g["mask"], text="a cat", timesteps=3, guidance_scale=2.0, temperature=1.0,
                      num_images_per_prompt=1, image_size=8, generator=torch.Generator().manual_seed(g["seed"]), **kw)
        got = fed[-1]
        assert set(ref["fed"]) <= set(got), (name, set(ref["fed"]) - set(got))
        assert torch.equal(got["input_ids"], ref["fed"]["input_ids"]), name  # tokenised image with the masked positions
        for k, v in ref["fed"].items():
            torch.testing.assert_close(got[k].float(), v.float(), rtol=1e-5, atol=1e-6, msg=lambda m: f"{name} {k}: {m}")
        a = np.stack([np.asarray(im) for im in images]).astype(np.int16)
        b = ref["images"].numpy().astype(np.int16)
        assert a.shape == b.shape and np.abs(a - b).max() <= 1 and (a != b).mean() < 0.02, name
