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


class PeerExchange:
    """Exchange buffers for the calibration statistics over NVLink peer memory (include/fp8fq.h:
    fp8fq_estimate_prepare_p2p_f32): one small symmetric-memory buffer per rank, every rank holding a device array of all
    ranks' buffer addresses.  ``next()`` hands out the (pointers, rank, world, epoch) tuple of the next exchange; every
    rank must issue the same sequence of exchanges (the calibration forward is SPMD, so it does)."""

    def __init__(self, device):
        import torch.distributed._symmetric_memory as symm_mem

        from ._lib import lib

        self.rank, self.world = td.get_rank(), td.get_world_size()
        words = int(lib().fp8fq_dp_exchange_words(self.world))
        if words <= 0:
            raise RuntimeError(f"peer exchange supports at most 16 ranks, got {self.world}")
        self.buf = symm_mem.empty(words, dtype=torch.int64, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, td.group.WORLD)     # collective: exchanges the memory handles
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        td.barrier()                    # nobody writes into a buffer that its owner has not zeroed yet
        self.epoch = 0

    def next(self):
        self.epoch += 1
        return self.ptrs, self.rank, self.world, self.epoch

    def unused(self):
        """The exchange handed out last did not take place (the same decision on every rank): give its epoch back."""
        self.epoch -= 1


_peer_exchange = None
_peer_exchange_tried = False


def peer_exchange(device=None):
    """The process-wide PeerExchange when the calibration collectives are active on NCCL / CUDA and symmetric memory
    could be set up on EVERY rank (the decision is all-reduced, so that no rank waits in a kernel for a peer that took
    the NCCL route); None otherwise -- the estimators then use the all-reduce path.  FP8FQ_DP_P2P=0 forces NCCL."""
    global _peer_exchange, _peer_exchange_tried
    if not active():
        return None
    if not _peer_exchange_tried:
        _peer_exchange_tried = True
        ok, px = 0, None
        if (os.environ.get("FP8FQ_DP_P2P", "1") != "0" and torch.cuda.is_available() and td.get_backend() == "nccl"
                and device is not None and device.type == "cuda" and td.get_world_size() <= 16):
            try:
                px = PeerExchange(device)
                ok = 1
            except Exception as exc:   # no P2P between the GPUs, symmetric memory unavailable in this build, ...
                import warnings

                warnings.warn(f"peer-memory exchange unavailable ({exc!r}); calibration statistics go through NCCL")
        if torch.cuda.is_available() and td.get_backend() == "nccl" and device is not None and device.type == "cuda":
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            td.all_reduce(flag, op=td.ReduceOp.MIN)
            ok = int(flag.item())
        _peer_exchange = px if ok else None
    return _peer_exchange


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
