"""CPU, world_size 2 (gloo): the N>1 recipe of the hot path.  The path shards by batch with ONE exchange per step --
the DDP gradient all-reduce (mean).  The reference's loss is a per-rank mean over that rank's masked tokens and DDP
averages the per-rank gradients equally; bench.py uses torch DDP the same way.  These tests pin that
recipe with the oracle as the per-rank model, and the launch contract of `bench.py --impl reference` under torchrun."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(vocab_size=72, hidden_size=64, num_hidden_layers=2, num_attention_heads=1, intermediate_size=128,
           max_position_embeddings=17, codebook_size=64, num_vq_tokens=16, num_classes=7, layer_norm_eps=1e-6)


class OracleModule(torch.nn.Module):
    def __init__(self, state_dict):
        super().__init__()
        self.names = list(state_dict)
        self.params = torch.nn.ParameterList([torch.nn.Parameter(v.clone()) for v in state_dict.values()])

    def forward(self, input_ids, labels):
        from oracle import transformer_oracle as T

        return T.forward(dict(zip(self.names, self.params)), CFG, input_ids, labels=labels)[1]


def _worker(rank, world, port, sd, batches, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = torch.nn.parallel.DistributedDataParallel(OracleModule(sd))
    inp, lab = batches[rank]
    loss = model(inp, lab)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in zip(model.module.names, model.module.params)}
    if rank == 0:
        torch.save(dict(grads=grads, loss=loss.detach()), out)
    dist.destroy_process_group()


def test_ddp_mean_of_per_rank_means(tmp_path, golden):
    from oracle import transformer_oracle as T

    g = golden("micro_transformer.pt")
    sd = g["state_dict"]
    gen = torch.Generator().manual_seed(0)
    batches = []
    for r in range(2):
        tokens = torch.randint(0, 64, (2 + r, 16), generator=gen)
        cls = torch.randint(0, 7, (2 + r,), generator=gen)
        batches.append(T.mask_tokens(tokens, cls, torch.rand(2 + r, generator=gen), torch.rand(2 + r, 16, generator=gen), 64, 71))
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, sd, batches, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    per_rank = [T.forward_backward(sd, CFG, b[0], b[1])[2] for b in batches]
    for k in sd:
        expect = 0.5 * (per_rank[0][k] + per_rank[1][k])  # equal-weight mean of per-rank (mean-loss) gradients
        torch.testing.assert_close(got["grads"][k], expect, rtol=1e-5, atol=1e-7)
    # and this is NOT the gradient of the global masked-token mean when ranks hold different numbers of masked tokens
    n0, n1 = [(b[1] != -100).sum().item() for b in batches]
    assert n0 != n1


def _product_worker(rank, world, port, sd, batches, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from open_muse_b200 import MaskGitTransformer
    from tests import cpu_math_ops

    real_bf16 = torch.bfloat16
    cpu_math_ops.install(cpu_math_ops.PlainSetter, exact=True)
    m = MaskGitTransformer(**dict(CFG, hidden_dropout=0.0, attention_dropout=0.0))
    m.load_state_dict(sd)
    m.train()
    model = torch.nn.parallel.DistributedDataParallel(m)
    inp, lab = batches[rank]
    _, loss = model(inp, labels=lab)
    loss.backward()
    torch.bfloat16 = real_bf16  # the alias of exact mode confuses torch.save's dtype tables
    if rank == 0:
        torch.save(dict(grads={n: p.grad.clone() for n, p in m.named_parameters()}, loss=loss.detach()), out)
    dist.destroy_process_group()


def test_product_model_under_ddp_matches_mean_of_oracle_rank_gradients(tmp_path, golden):
    """The PRODUCT's MaskGitTransformer (its per-layer autograd Functions hand their parameter gradients back layer by layer, so
    DDP's bucket hooks fire during the backward) under torch DDP, gloo world 2, kernels replaced by their exact fp32 torch
    restatements: the all-reduced gradients equal the equal-weight mean of the oracle's per-rank gradients."""
    from oracle import transformer_oracle as T

    g = golden("micro_transformer.pt")
    sd = g["state_dict"]
    gen = torch.Generator().manual_seed(1)
    batches = []
    for r in range(2):
        tokens = torch.randint(0, 64, (2 + r, 16), generator=gen)
        cls = torch.randint(0, 7, (2 + r,), generator=gen)
        batches.append(T.mask_tokens(tokens, cls, torch.rand(2 + r, generator=gen), torch.rand(2 + r, 16, generator=gen), 64, 71))
    out = str(tmp_path / "r0.pt")
    mp.spawn(_product_worker, args=(2, 31500 + os.getpid() % 2000, sd, batches, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    per_rank = [T.forward_backward(sd, CFG, b[0], b[1])[2] for b in batches]
    for k in sd:
        expect = 0.5 * (per_rank[0][k] + per_rank[1][k])
        torch.testing.assert_close(got["grads"][k], expect, rtol=2e-4, atol=1e-7)


def _uvit_worker(rank, world, port, g, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from open_muse_b200 import MaskGiTUViT_v2
    from tests import cpu_math_ops

    real_bf16 = torch.bfloat16
    cpu_math_ops.install(cpu_math_ops.PlainSetter, exact=True)
    m = MaskGiTUViT_v2(**g["config"])
    m.load_state_dict(g["state_dict"])
    m.train()
    model = torch.nn.parallel.DistributedDataParallel(m)
    sl = slice(0, 2) if rank == 0 else slice(2, 3)  # uneven per-rank batches
    _, loss = model(g["input_ids"][sl], g["encoder_hidden_states"][sl], g["cond_embeds"][sl], g["micro_conds"][sl],
                    labels=g["labels"][sl], label_smoothing=0.1)
    loss.backward()
    torch.bfloat16 = real_bf16
    if rank == 0:
        torch.save({n: p.grad.clone() for n, p in m.named_parameters()}, out)
    dist.destroy_process_group()


def test_uvit_product_model_under_ddp_matches_mean_of_oracle_rank_gradients(tmp_path, golden):
    """MaskGiTUViT_v2's per-block autograd Functions (text states and the conditioning vector enter every block as shared
    inputs whose gradients autograd sums; each block hands its parameter gradients back as soon as it has run) under torch
    DDP, gloo world 2, uneven per-rank batches: all-reduced gradients == equal-weight mean of the oracle's per-rank ones."""
    from oracle import transformer_v2_oracle as V2

    g = golden("micro_uvit_v2.pt")
    out = str(tmp_path / "r0.pt")
    mp.spawn(_uvit_worker, args=(2, 33500 + os.getpid() % 2000, g, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    per_rank = []
    for sl in (slice(0, 2), slice(2, 3)):
        q = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
        _, loss = V2.forward(q, g["config"], g["input_ids"][sl], g["encoder_hidden_states"][sl], g["cond_embeds"][sl],
                             g["micro_conds"][sl], labels=g["labels"][sl], label_smoothing=0.1)
        loss.backward()
        per_rank.append({k: v.grad for k, v in q.items()})
    for k in got:
        expect = 0.5 * (per_rank[0][k] + per_rank[1][k])
        assert float((got[k] - expect).norm() / expect.norm().clamp_min(1e-20)) < 1e-4, k


def test_reference_training_script_under_two_rank_ddp(tmp_path):
    """The UNMODIFIED training/train_maskgit_imagenet.py on two ranks (gloo): the accelerate stand-in wraps the drop-in model
    in torch DDP, the script's own multi-process plumbing runs -- loss gather for logging, main-process guards around
    evaluation / checkpoint writing, wait_for_everyone, state-dict retrieval through the DDP wrapper -- and the ranks end with
    identical weights (the all-reduced gradients of the product's per-layer Functions)."""
    from tests.train_script_harness import find_script, make_config

    script = find_script()
    if script is None:
        import pytest

        pytest.skip("reference training script not available")
    cfg, out = make_config(str(tmp_path), steps=3, batch=2, mixed_precision="no", save_every=3)
    res = str(tmp_path / "res.json")
    env = dict(os.environ, OMP_NUM_THREADS="2", ACCELERATE_USE_CPU="1", WANDB_MODE="disabled")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29900 + os.getpid() % 90), os.path.join(ROOT, "tests", "ddp_script_launcher.py"), script, cfg, res]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    ranks = [json.load(open(f"{res}.rank{i}")) for i in range(2)]
    assert [x["world"] for x in ranks] == [2, 2] and ranks[0]["is_main"] and not ranks[1]["is_main"]
    steps0 = [s for v, s in ranks[0]["logged"] if "step_loss" in v]
    assert steps0 == [1, 2, 3] and any("eval_loss" in v for v, _ in ranks[0]["logged"])
    # the gathered loss is the same number on both ranks, the checkpoint was written once (by the main process)
    l0 = [v["step_loss"] for v, s in ranks[0]["logged"] if "step_loss" in v]
    l1 = [v["step_loss"] for v, s in ranks[1]["logged"] if "step_loss" in v]
    assert l0 == l1
    assert os.path.exists(os.path.join(out, "checkpoint-3", "unwrapped_model", "pytorch_model.bin"))


def test_reference_train_muse_script_under_two_rank_ddp(tmp_path):
    """The UNMODIFIED training/train_muse.py on two ranks (gloo): MaskGiTUViT_v2's per-block Functions under the DDP the
    accelerate stand-in builds, text encoder + tokenizer per rank, EMA, main-process checkpointing."""
    from tests.train_script_harness import find_script, make_muse_config

    script = find_script("train_muse.py")
    if script is None:
        import pytest

        pytest.skip("reference training script not available")
    cfg, out = make_muse_config(str(tmp_path), steps=2, batch=2, mixed_precision="no", save_every=2, use_ema=True)
    res = str(tmp_path / "res.json")
    env = dict(os.environ, OMP_NUM_THREADS="2", ACCELERATE_USE_CPU="1", WANDB_MODE="disabled")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(30100 + os.getpid() % 90), os.path.join(ROOT, "tests", "ddp_script_launcher.py"), script, cfg, res]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    ranks = [json.load(open(f"{res}.rank{i}")) for i in range(2)]
    assert [x["world"] for x in ranks] == [2, 2]
    l0 = [v["step_loss"] for v, s in ranks[0]["logged"] if "step_loss" in v]
    l1 = [v["step_loss"] for v, s in ranks[1]["logged"] if "step_loss" in v]
    assert len(l0) == 2 and l0 == l1
    assert os.path.isdir(os.path.join(out, "checkpoint-2", "ema_model"))


def test_reference_arm_prints_only_on_rank0():
    env = dict(os.environ, OMP_NUM_THREADS="2", MUSE_B200_CPU_SAMPLE_BATCH="4")  # the launch contract, not the number
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "bench.py"), "--impl",
           "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    # "reference": the unmodified reference modules from the oracle/_ref snapshot; "port": the oracle restatement
    from oracle import ref_snapshot

    assert d["impl"] == "reference" and d["n_gpus"] == 2
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_snapshot.available() else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "images/s"
