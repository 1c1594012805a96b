"""Runs the UNMODIFIED reference training script (training/train_maskgit_imagenet.py) against the drop-in ``muse`` package
(open_muse_b200/compat) for a couple of optimizer steps -- the "scripts run unchanged" half of the boundary (north_star,
SURVEY.md 8c).  Third-party packages the image lacks come from tests/shims; the script itself is executed with runpy from
where it lies: /root/reference/training in the build container, oracle/_ref/training (the git-ignored snapshot taken by
__graft_entry__.build()) on the GPU box.  Test infrastructure only."""
import os
import runpy
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "tests", "shims")
COMPAT = os.path.join(ROOT, "open_muse_b200", "compat")

MICRO_VQ = dict(resolution=32, num_channels=3, hidden_channels=32, channel_mult=(1, 2), num_res_blocks=1,
                z_channels=16, num_embeddings=64, quantized_embed_dim=16)


def find_script(name="train_maskgit_imagenet.py"):
    for base in (os.environ.get("MUSE_REFERENCE", "/root/reference"), os.path.join(ROOT, "oracle", "_ref")):
        p = os.path.join(base, "training", name)
        if os.path.exists(p):
            return p
    return None


def make_config(tmp, steps, batch, mixed_precision, soft_targets=False, save_every=1000, transformer=None):
    from open_muse_b200 import MaskGitVQGAN

    torch.manual_seed(3)
    vq_dir = os.path.join(tmp, "vq")
    MaskGitVQGAN(**MICRO_VQ).save_pretrained(vq_dir)
    out = os.path.join(tmp, "run")
    cfg = {
        "wandb": {"entity": None},
        "experiment": {"project": "muse", "name": "shim-run", "output_dir": out, "max_train_examples": batch * 64,
                       "max_eval_examples": batch * 2, "save_every": save_every, "eval_every": 1000,
                       "generate_every": 1000,  # quirk Q11: generate_images crashes upstream; keep it beyond max_train_steps
                       "log_every": 1, "log_grad_norm_every": 1, "resume_from_checkpoint": False, "resume_lr_scheduler": True},
        "model": {"vq_model": {"type": "maskgit_vqgan", "pretrained": vq_dir},
                  "transformer": {"vocab_size": 64 + 10 + 1, "max_position_embeddings": 257, "hidden_size": 64,
                                  "num_hidden_layers": 2, "num_attention_heads": 1, "intermediate_size": 128,
                                  "codebook_size": 64, "num_vq_tokens": 256, "num_classes": 10, "initializer_range": 0.02,
                                  "norm_type": "layernorm", "layer_norm_eps": 1e-6, "use_normformer": True,
                                  "use_encoder_layernorm": True, "use_mlm_layer": True, "use_mlm_layernorm": True,
                                  "use_bias": False, "hidden_dropout": 0.0, "attention_dropout": 0.0},
                  "gradient_checkpointing": True, "enable_xformers_memory_efficient_attention": True},
        "dataset": {"params": {"train_shards_path_or_url": "synthetic", "eval_shards_path_or_url": "synthetic",
                               "batch_size": batch, "shuffle_buffer_size": 10, "num_workers": 0, "resolution": 32,
                               "pin_memory": False, "persistent_workers": False},
                    "preprocessing": {"resolution": 32, "center_crop": True, "random_flip": False}},
        "optimizer": {"name": "adamw", "params": {"learning_rate": 1.0e-3, "scale_lr": False, "beta1": 0.9, "beta2": 0.999,
                                                  "weight_decay": 0.01, "epsilon": 1.0e-8}},
        "lr_scheduler": {"scheduler": "constant_with_warmup", "params": {"learning_rate": 1.0e-3, "warmup_steps": 1}},
        "training": {"gradient_accumulation_steps": 1, "batch_size": batch, "mixed_precision": mixed_precision,
                     "enable_tf32": True, "use_ema": False, "seed": 42, "max_train_steps": steps, "overfit_one_batch": False,
                     "min_masking_rate": 0.0, "label_smoothing": 0.1, "max_grad_norm": 1.0,
                     "use_soft_code_target": soft_targets, "use_stochastic_code": False, "soft_code_temp": 1.0},
    }
    if transformer is not None:  # e.g. the model section of the reference's own configs/imagenet.yaml
        cfg["model"]["transformer"] = dict(transformer)
    path = os.path.join(tmp, "config.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path, out


def make_tiny_clip(path, projection_dim=768, weight_std=None):
    """A 2-layer, 32-wide CLIP text encoder with a 768-wide projection (training/train_muse.py:336 forces projection_dim=768)
    and a 54-entry byte-pair vocabulary, saved where ``CLIPTextModelWithProjection / CLIPTokenizer.from_pretrained`` find them:
    the text encoder is third-party and out of scope, the script just needs one to call."""
    import json

    from transformers import CLIPTextConfig, CLIPTextModelWithProjection, CLIPTokenizer

    os.makedirs(path, exist_ok=True)
    chars = [chr(c) for c in range(ord("a"), ord("z") + 1)]
    vocab = {t: i for i, t in enumerate(chars + [c + "</w>" for c in chars] + ["<|startoftext|>", "<|endoftext|>"])}
    with open(os.path.join(path, "vocab.json"), "w") as f:
        json.dump(vocab, f)
    with open(os.path.join(path, "merges.txt"), "w") as f:
        f.write("#version: 0.2\n")
    CLIPTokenizer(os.path.join(path, "vocab.json"), os.path.join(path, "merges.txt"), model_max_length=8).save_pretrained(path)
    cfg = CLIPTextConfig(vocab_size=len(vocab), hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                         num_attention_heads=2, max_position_embeddings=8, projection_dim=projection_dim,
                         bos_token_id=vocab["<|startoftext|>"], eos_token_id=vocab["<|endoftext|>"],
                         pad_token_id=vocab["<|endoftext|>"])
    torch.manual_seed(5)
    clip = CLIPTextModelWithProjection(cfg)
    if weight_std is not None:  # large random matrices: prompts, layers and the empty prompt give clearly different outputs
        with torch.no_grad():
            for p in clip.parameters():
                if p.dim() > 1:
                    p.normal_(0.0, weight_std)
    clip.save_pretrained(path)
    return path


UVIT_MICRO = dict(hidden_size=128, num_attention_heads=2, in_channels=64, block_out_channels=[64], block_num_heads=1,
                  num_res_blocks=1, num_hidden_layers=2, intermediate_size=128, vocab_size=72, codebook_size=64,
                  encoder_hidden_size=32, cond_embed_dim=768, micro_cond_encode_dim=8, micro_cond_embed_dim=40,
                  norm_type="rmsnorm", add_cond_embeds=True, add_micro_cond_embeds=True, use_empty_embeds_for_uncond=True,
                  hidden_dropout=0.0, attention_dropout=0.0)


T2I_MICRO = dict(vocab_size=72, hidden_size=64, num_hidden_layers=2, num_attention_heads=1, intermediate_size=128,
                 hidden_dropout=0.0, attention_dropout=0.0, max_position_embeddings=256, codebook_size=64, num_vq_tokens=256,
                 add_cross_attention=True, encoder_hidden_size=32, norm_type="rmsnorm", use_normformer=False,
                 layer_norm_eps=1e-6, use_codebook_size_for_output=True)


def make_muse_config(tmp, steps, batch, mixed_precision, save_every=1000, extra_experiment=None, use_ema=False,
                     f16_tokenizer=False, v1_soft_targets=False):
    """config for the UNMODIFIED training/train_muse.py on the one wiring that runs end to end at this commit (quirk Q12):
    ``architecture: uvit`` (MaskGiTUViT = MaskGiTUViT_v2) with pooled + micro conditioning, CLIP text encoder with
    projection, classifier-free-guidance dropout through the encoded empty prompt, taming-style VQGAN tokenizer."""
    from open_muse_b200 import VQGANModel

    torch.manual_seed(3)
    vq_dir = os.path.join(tmp, "vq")
    res = 64 if f16_tokenizer else 32  # f16_tokenizer: five levels (x16 reduction, what the script's inpainting-mask helper
    mult = (1, 1, 1, 1, 1) if f16_tokenizer else (1, 2)  # hard-codes), 64 px -> 4 x 4 = 16 tokens; else x2, 32 px -> 256
    VQGANModel(resolution=res, num_channels=3, hidden_channels=32, channel_mult=mult, num_res_blocks=1,
               attn_resolutions=(16,), z_channels=16, num_embeddings=64, quantized_embed_dim=16).save_pretrained(vq_dir)
    clip_dir = make_tiny_clip(os.path.join(tmp, "clip"))
    out = os.path.join(tmp, "run")
    exp = {"project": "muse", "name": "shim-run", "output_dir": out, "max_train_examples": batch * 64,
           "max_eval_examples": batch * 2, "save_every": save_every, "eval_every": 1000, "generate_every": 1000,
           "log_every": 1, "log_grad_norm_every": 1, "resume_from_checkpoint": False, "resume_lr_scheduler": True}
    exp.update(extra_experiment or {})
    cfg = {
        "wandb": {"entity": None},
        "experiment": exp,
        "model": {"architecture": "uvit", "vq_model": {"type": "vqgan", "pretrained": vq_dir},
                  "text_encoder": {"type": "clip", "pretrained": clip_dir},
                  "transformer": dict(UVIT_MICRO, num_vq_tokens=16 if f16_tokenizer else 256),
                  "gradient_checkpointing": True, "enable_xformers_memory_efficient_attention": True},
        "dataset": {"type": "text2image",
                    "params": {"train_shards_path_or_url": "synthetic", "eval_shards_path_or_url": "synthetic",
                               "batch_size": batch, "shuffle_buffer_size": 10, "num_workers": 0, "resolution": res,
                               "pin_memory": False, "persistent_workers": False, "validation_prompts_file": None},
                    "preprocessing": {"resolution": res, "center_crop": True, "random_flip": False, "max_seq_length": 8}},
        "optimizer": {"name": "adamw", "params": {"learning_rate": 1.0e-3, "scale_lr": False, "beta1": 0.9, "beta2": 0.999,
                                                  "weight_decay": 0.01, "epsilon": 1.0e-8}},
        "lr_scheduler": {"scheduler": "constant_with_warmup", "params": {"learning_rate": 1.0e-3, "warmup_steps": 1}},
        "training": {"gradient_accumulation_steps": 1, "batch_size": batch, "mixed_precision": mixed_precision,
                     "enable_tf32": True, "use_ema": use_ema, "ema_decay": 0.99, "ema_update_after_step": 0,
                     "ema_update_every": 1, "seed": 42, "max_train_steps": steps, "overfit_one_batch": False,
                     "cond_dropout_prob": 0.1, "guidance_scale": 2.0, "generation_timesteps": 3, "min_masking_rate": 0.0, "label_smoothing": 0.1, "max_grad_norm": 1.0,
                     "use_soft_code_target": False, "use_stochastic_code": False, "soft_code_temp": 1.0},
    }
    if v1_soft_targets:
        # the OTHER wiring of train_muse.py that runs upstream (quirk Q12): architecture "transformer" = the text-conditional
        # MaskGitTransformer (BASELINE config 4's model class), plain CLIPTextModel, no CFG dropout, soft code targets
        cfg["model"]["architecture"] = "transformer"
        cfg["model"]["transformer"] = dict(T2I_MICRO)
        cfg["training"].update(cond_dropout_prob=0.0, use_soft_code_target=True, soft_code_temp=2.0)
    path = os.path.join(tmp, "config.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path, out


def run_script(script, config_path, extra_cli=()):
    """Execute the script as __main__ with the shims and the drop-in package importable; returns the shim Accelerator the
    script created (its .logged list holds every accelerator.log call)."""
    saved_path, saved_argv = list(sys.path), list(sys.argv)
    shimmed = ("accelerate", "omegaconf", "wandb", "data", "optimizer", "plotly")
    saved_mods = {k: sys.modules.get(k) for k in ("muse",) + shimmed}
    for k in list(sys.modules):
        if k == "muse" or k.startswith("muse.") or k in shimmed or k.startswith("accelerate.") or k.startswith("plotly."):
            del sys.modules[k]
    sys.path[:0] = [SHIMS, COMPAT, os.path.dirname(script)]  # shims shadow the script directory's own data.py
    sys.argv = [script, f"config={config_path}", *extra_cli]
    try:
        import accelerate

        runpy.run_path(script, run_name="__main__")
        return accelerate.Accelerator.last
    finally:
        sys.path[:], sys.argv[:] = saved_path, saved_argv
        for k in list(sys.modules):
            if k == "muse" or k.startswith("muse.") or k in shimmed or k.startswith("accelerate.") or k.startswith("plotly."):
                del sys.modules[k]
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v
