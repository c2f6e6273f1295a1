"""Process-group plumbing and document sharding (replaces improved_diffusion/dist_util.py:21-50).

One process per GPU, torchrun-style environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Documents are independent,
so the hot path has NO collective: rank r processes documents r, r+world, r+2*world, ...  The only communication is a
gather of per-document timings/metrics at the end (NCCL on GPUs, Gloo on CPU) and the reference's trailing barrier
(val_TDiff.py:115).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def _free_port() -> int:
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def setup_dist(backend: str | None = None):
    """dist_util.py:21-41 without the MPI bootstrap.  Returns (rank, world, device).  Like the reference, a process group is ALWAYS
    initialised (also for a single process), because val_TDiff.run() ends with an unconditional dist.barrier() (val_TDiff.py:115)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if torch.cuda.is_available():
        dev = torch.device("cuda", local % max(torch.cuda.device_count(), 1))     # dist_util.py:44-50 dev()
        torch.cuda.set_device(dev)
    else:
        dev = torch.device("cpu")
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if world == 1:
            os.environ.setdefault("MASTER_PORT", str(_free_port()))
        else:
            os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if dev.type == "cuda" else "gloo")
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, dev


def load_state_dict(path, **kwargs):
    """dist_util.py:53-63 without the MPI broadcast: every rank reads the file itself."""
    return torch.load(path, **kwargs)


def dev() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def shard_documents(n_docs: int, rank: int, world: int) -> list[int]:
    """Document ids owned by `rank` (round robin, so ragged tails spread evenly)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_docs, world))


def batches(ids: list[int], batch: int) -> list[list[int]]:
    return [ids[i:i + batch] for i in range(0, len(ids), batch)]


def gather_metrics(values: dict, device: torch.device | None = None) -> list[dict]:
    """All ranks contribute a small dict of {doc_id: seconds}; every rank receives the list of all dicts.
    Timing/metrics only — never on the data path."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [values]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, values)
    return out


def max_over_ranks(x: float, device: torch.device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def barrier():
    if dist.is_initialized():
        dist.barrier()
