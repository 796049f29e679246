"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box,
gloo in the CPU tests).

The fake-quant path shards over the batch with NO data-path collective.  The
only exchanges are the calibration statistics:
  * min/max estimators : one MAX all-reduce of the packed [-min, max] vector per quantiser,
  * MSE estimator      : MAX of absmax (grid definition) + MEAN of the per-batch MSE table,
  * validation metrics : one SUM of [correct@1, correct@5, loss_sum, count].
They are a few bytes each, hence latency-bound; NCCL is the right tool (no compute to fuse with).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as td

_group_enabled = False
counters = {"all_reduce": 0}   # collectives issued through this module (bench.py reports them per calibration pass)


def init_from_env(backend: str = None) -> bool:
    """Initialise the default process group from torchrun's environment.  Returns True when
    world_size > 1 and calibration collectives are enabled."""
    global _group_enabled
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return False
    if not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        td.init_process_group(backend=backend)
    _group_enabled = True
    return True


def enable(flag: bool = True):
    """Turn the calibration collectives on/off (requires an initialised process group)."""
    global _group_enabled
    _group_enabled = bool(flag) and td.is_available() and td.is_initialized() and td.get_world_size() > 1


def active() -> bool:
    return _group_enabled


def world_size() -> int:
    return td.get_world_size() if (td.is_available() and td.is_initialized()) else 1


def rank() -> int:
    return td.get_rank() if (td.is_available() and td.is_initialized()) else 0


def all_reduce_max(t: torch.Tensor):
    counters["all_reduce"] += 1
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return t


def all_reduce_sum(t: torch.Tensor):
    counters["all_reduce"] += 1
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def all_reduce_mean(t: torch.Tensor):
    """Equal-sized shards: mean over ranks of per-shard means == mean over the global batch."""
    counters["all_reduce"] += 1
    td.all_reduce(t, op=td.ReduceOp.SUM)
    t.div_(td.get_world_size())
    return t


def barrier():
    if td.is_available() and td.is_initialized():
        td.barrier()


def shard_batch(x: torch.Tensor) -> torch.Tensor:
    """This rank's contiguous slice of a global batch (dim 0)."""
    w, r = world_size(), rank()
    if w == 1:
        return x
    if x.shape[0] % w != 0:
        raise ValueError(f"global batch {x.shape[0]} is not divisible by world size {w}")
    per = x.shape[0] // w
    return x[r * per:(r + 1) * per].contiguous()
