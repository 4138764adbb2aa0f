"""Data-parallel plumbing for the CHAP iteration: one process per GPU, `torch.distributed` (NCCL over NVLink on the
B200 box, gloo in the CPU tests).  The reference is single-GPU; the hot path shards over samples, the
only exchange is ONE all-reduce of the flat gradient arena per iteration."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def make_grad_hook(world_size):
    """grad_hook for FlatSGD: sum the flat gradient over ranks (the 1/world factor is folded into the SGD kernel's
    grad_scale).  None for a single process."""
    if world_size <= 1:
        return None

    def hook(flat_grad):
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return hook


def broadcast_parameters(flat_params, src=0):
    """All replicas start from rank `src`'s weights (one broadcast of the flat arena)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_params, src)


def shard_round_robin(n_items, rank, world_size):
    """Indices of the cases / windows owned by `rank` (sliding-window inference is sharded by case)."""
    return list(range(rank, n_items, world_size))


def gather_rows(rows, world_size):
    """Gather per-case metric rows (numpy [k, m]) from all ranks onto every rank, in case order."""
    if world_size <= 1 or not dist.is_initialized():
        return rows
    out = [None] * world_size
    dist.all_gather_object(out, rows)
    return out


class DevicePrefetcher:
    """Host -> device input pipeline for the training loop (the reference's DataLoader uses pin_memory=True and then a blocking
    `.cuda()` per batch, code/train_ours_2D.py:274,304-305): batch i + 1 is copied from pinned host memory on a side stream while
    iteration i computes.  Iterating yields tuples of device tensors; every yielded tensor is safe to use on the current stream."""

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.next = None
        self._load()

    def _load(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self.next = None
            return
        with torch.cuda.stream(self.stream):
            self.next = tuple(t.to(self.device, non_blocking=True) for t in batch)

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)                       # the copy of THIS batch has finished before it is used
        batch = self.next
        for t in batch:
            t.record_stream(cur)                           # allocated on the side stream, consumed on the current one
        self._load()                                       # start the next copy; it overlaps with the caller's iteration
        return batch
