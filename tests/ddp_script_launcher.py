"""Launched by torch.distributed.run (one process per rank) from tests/test_multi_rank_cpu.py: installs the numeric kernel
restatements in this process and executes the UNMODIFIED reference training script against the drop-in package under the
accelerate stand-in, which builds a gloo process group and wraps the model in torch DDP when WORLD_SIZE > 1."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import transformers  # noqa: E402,F401  (before the accelerate stand-in is on sys.path: its own accelerate probing needs a real version)

from tests import cpu_math_ops  # noqa: E402
from tests.train_script_harness import run_script  # noqa: E402

if __name__ == "__main__":
    script, cfg, out_json = sys.argv[1:4]
    torch.set_num_threads(2)
    os.environ.setdefault("ACCELERATE_USE_CPU", "1")
    os.environ.setdefault("WANDB_MODE", "disabled")
    cpu_math_ops.install(cpu_math_ops.PlainSetter, exact=False)
    acc = run_script(script, cfg)
    rank = int(os.environ.get("RANK", "0"))
    with open(f"{out_json}.rank{rank}", "w") as f:
        json.dump(dict(rank=rank, world=acc.num_processes, is_main=acc.is_main_process,
                       logged=[(v, s) for v, s in acc.logged if "step_loss" in v or "eval_loss" in v]), f)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
