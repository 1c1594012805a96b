"""Checkpoints cross the boundary in BOTH directions (no GPU): a model saved by this package's ``save_pretrained`` loads into
the UNMODIFIED reference class with ``from_pretrained`` (every key consumed, none missing), and a model saved by the reference
loads here -- ``config.json`` keys, ``pytorch_model.bin`` names and shapes are the compatibility contract.  The
reference runs in a subprocess (its package is also called ``muse``); the forward of the loaded weights is then compared
numerically: the reference's own CPU forward against this package's host code on the kernels' torch restatements."""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref_snapshot
from tests import cpu_math_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_snapshot.available(), reason="oracle/_ref snapshot of the reference not available")

V1 = dict(vocab_size=72, hidden_size=64, num_hidden_layers=2, num_attention_heads=1, intermediate_size=128, hidden_dropout=0.0,
          attention_dropout=0.0, max_position_embeddings=17, codebook_size=64, num_vq_tokens=16, num_classes=7)
V2 = dict(hidden_size=128, num_attention_heads=2, in_channels=64, block_out_channels=[64], block_num_heads=1, num_res_blocks=1,
          num_hidden_layers=2, intermediate_size=128, vocab_size=72, codebook_size=64, encoder_hidden_size=32, cond_embed_dim=16,
          micro_cond_encode_dim=8, micro_cond_embed_dim=40, norm_type="rmsnorm", force_down_up_sample=True)
VQ = dict(resolution=32, num_channels=3, hidden_channels=32, channel_mult=[1, 2], num_res_blocks=1, z_channels=16,
          num_embeddings=64, quantized_embed_dim=16)
TVQ = dict(VQ, num_res_blocks=2, attn_resolutions=[16])

REF_SIDE = r'''
import json, sys, torch
sys.path.insert(0, sys.argv[1])
from oracle.ref_snapshot import import_reference
muse = import_reference()
from muse.modeling_transformer_v2 import MaskGiTUViT_v2
classes = {"MaskGitTransformer": muse.MaskGitTransformer, "MaskGiTUViT_v2": MaskGiTUViT_v2, "MaskGitVQGAN": muse.MaskGitVQGAN,
           "VQGANModel": muse.VQGANModel}
for kind, base in json.load(open(sys.argv[2])):  # one job per model class: <base>/{ours, theirs, inputs.pt, cfg.json}
    cls = classes[kind]
    dir_a, dir_b = base + "/ours", base + "/theirs"
    inputs = torch.load(base + "/inputs.pt")
    # (1) what this package saved loads into the unmodified reference class
    m = cls.from_pretrained(dir_a, low_cpu_mem_usage=False)
    saved = torch.load(dir_a + "/pytorch_model.bin")
    sd = m.state_dict()
    assert list(sd) == list(saved), (kind, set(sd) ^ set(saved))
    assert all(torch.equal(sd[k], saved[k]) for k in sd), kind
    # (2) a differently seeded reference model, saved by the reference, with its own forward output
    torch.manual_seed(1234)
    r = cls(**json.load(open(base + "/cfg.json")))
    with torch.no_grad():
        for p in r.parameters():  # every tensor non-trivial (norm weights are ones, adaLN mappers and GRN parameters zeros at init)
            p.add_((0.1 if p.dim() == 1 else 0.02) * torch.randn_like(p))
    r.eval()
    r.save_pretrained(dir_b)
    with torch.no_grad():
        if kind == "MaskGitTransformer":
            out = r(inputs["input_ids"])
        elif kind == "MaskGiTUViT_v2":
            out = r(inputs["input_ids"], inputs["enc"], inputs["cond"], inputs["micro"])
        else:
            out = r.decode_code(r.get_code(inputs["image"]))
    torch.save(out, dir_b + "/reference_output.pt")
print("OK")
'''


def _inputs(kind):
    g = torch.Generator().manual_seed(7)
    if kind == "MaskGitTransformer":
        return dict(input_ids=torch.randint(0, 72, (2, 17), generator=g))
    if kind == "MaskGiTUViT_v2":
        return dict(input_ids=torch.randint(0, 72, (2, 64), generator=g), enc=torch.randn(2, 5, 32, generator=g),
                    cond=torch.randn(2, 16, generator=g), micro=torch.tensor([[256.0, 256.0, 0.0, 0.0, 6.0]] * 2))
    return dict(image=torch.rand(2, 3, 32, 32, generator=g))


KINDS = [("MaskGitTransformer", V1), ("MaskGiTUViT_v2", V2), ("MaskGitVQGAN", VQ), ("VQGANModel", TVQ)]


@pytest.fixture(scope="module")
def exchanged(tmp_path_factory):
    """this package saves one model per class, ONE reference subprocess loads them all and saves its own"""
    import open_muse_b200 as ours

    root = tmp_path_factory.mktemp("interchange")
    jobs = []
    for kind, cfg in KINDS:
        base = str(root / kind)
        os.makedirs(base)
        torch.manual_seed(99)
        getattr(ours, kind)(**cfg).save_pretrained(os.path.join(base, "ours"))
        torch.save(_inputs(kind), os.path.join(base, "inputs.pt"))
        with open(os.path.join(base, "cfg.json"), "w") as f:
            json.dump(cfg, f)
        jobs.append((kind, base))
    with open(root / "jobs.json", "w") as f:
        json.dump(jobs, f)
    r = subprocess.run([sys.executable, "-c", REF_SIDE, ROOT, str(root / "jobs.json")], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr[-3000:]
    return dict(jobs)


@pytest.mark.parametrize("kind,cfg", KINDS)
def test_checkpoints_are_interchangeable_with_the_unmodified_reference(exchanged, monkeypatch, kind, cfg):
    import open_muse_b200 as ours

    cls = getattr(ours, kind)
    dir_b = os.path.join(exchanged[kind], "theirs")
    inp = _inputs(kind)
    # (3) what the reference saved loads here: same keys / shapes / values, same config
    theirs = torch.load(os.path.join(dir_b, "pytorch_model.bin"))
    cpu_math_ops.install(monkeypatch, exact=True)
    monkeypatch.setenv("MUSE_B200_CUDA_GRAPH", "0")
    loaded = cls.from_pretrained(dir_b).eval()
    sd = loaded.state_dict()
    assert list(sd) == list(theirs) and all(torch.equal(sd[k], theirs[k]) for k in sd)
    ref_cfg = json.load(open(os.path.join(dir_b, "config.json")))
    for k, v in cfg.items():
        assert ref_cfg[k] == v and getattr(loaded.config, k) == (tuple(v) if isinstance(v, list) else v) or \
            list(getattr(loaded.config, k)) == list(v), k
    # (4) and computes what the reference computed with them
    want = torch.load(os.path.join(dir_b, "reference_output.pt"))
    with torch.no_grad():
        if kind == "MaskGitTransformer":
            got = loaded(inp["input_ids"])
        elif kind == "MaskGiTUViT_v2":
            got = loaded(inp["input_ids"], inp["enc"], inp["cond"], inp["micro"])
        else:
            got = loaded.decode_code(loaded.get_code(inp["image"]))
    assert got.shape == want.shape
    assert float((got.float() - want).norm() / want.norm()) < (5e-5 if "VQ" not in kind else 1e-3)


SURFACE_SIDE = r'''
import sys, torch
sys.path.insert(0, sys.argv[1])
from oracle.ref_snapshot import import_reference
ref = import_reference()
from muse.modeling_transformer_v2 import MaskGiTUViT_v2 as RefV2
import open_muse_b200 as ours
base = set(dir(torch.nn.Module))
V1 = dict(vocab_size=72, hidden_size=64, num_hidden_layers=1, num_attention_heads=1, intermediate_size=128, add_cross_attention=True,
          encoder_hidden_size=32, project_encoder_hidden_states=True)
V2 = dict(hidden_size=128, num_attention_heads=2, in_channels=64, block_out_channels=(64,), block_num_heads=1, num_res_blocks=1,
          num_hidden_layers=1, intermediate_size=128, vocab_size=72, codebook_size=64, encoder_hidden_size=32, cond_embed_dim=16,
          micro_cond_encode_dim=8, micro_cond_embed_dim=40)
VQ = dict(resolution=32, hidden_channels=32, channel_mult=(1, 2), num_res_blocks=1, z_channels=16, num_embeddings=64,
          quantized_embed_dim=16)
bad = []
for name, r, o, kw in (("MaskGitTransformer", ref.MaskGitTransformer, ours.MaskGitTransformer, V1), ("MaskGiTUViT_v2", RefV2, ours.MaskGiTUViT_v2, V2),
                       ("MaskGitVQGAN", ref.MaskGitVQGAN, ours.MaskGitVQGAN, VQ), ("VQGANModel", ref.VQGANModel, ours.VQGANModel, VQ),
                       ("EMAModel", ref.EMAModel, ours.EMAModel, None), ("PipelineMuse", ref.PipelineMuse, ours.PipelineMuse, None),
                       ("PipelineMuseInpainting", ref.PipelineMuseInpainting, ours.PipelineMuseInpainting, None)):
    missing = ({a for a in dir(r) if not a.startswith("_")} - base) - {a for a in dir(o) if not a.startswith("_")}
    if missing:
        bad.append((name, "class attributes", sorted(missing)))
    if kw is None:
        continue
    with torch.device("meta"):
        a, b = r(**kw), o(**kw)
    missing = {k for k in vars(a) if not k.startswith("_")} - {k for k in vars(b) if not k.startswith("_")}
    if missing:
        bad.append((name, "instance attributes", sorted(missing)))
    if list(a.config.keys()) != list(b.config.keys()):
        bad.append((name, "config keys", sorted(set(a.config.keys()) ^ set(b.config.keys()))))
    if [n for n, _ in a.named_modules()] != [n for n, _ in b.named_modules()]:
        bad.append((name, "module tree", "differs"))
print("BAD", bad) if bad else print("OK")
'''


def test_public_surface_of_every_class_covers_the_reference():
    """every public class attribute / method, every public instance attribute, the config keys in order and the module tree
    (names, order) of the reference classes exist on the drop-in classes (reference imported in a subprocess)"""
    r = subprocess.run([sys.executable, "-c", SURFACE_SIDE, ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stdout[-2000:], r.stderr[-2000:])
