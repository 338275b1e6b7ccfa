"""Data-parallel plumbing for the frame-pair hot path.

Frame pairs are independent units, BatchNorm statistics and the loss normaliser are per GPU
(/root/reference/exps/train_utils/double_trainer.py:171 ``broadcast_buffers=False``, no SyncBN), so
forward+loss needs no data-path collective: each rank takes ``global_batch / world`` pairs
(/root/reference/cfgs/s_s50_onex_dfp_tal_flip.py:93-94).  The only exchanges are the timing
reduction of the benchmark (max over ranks) and, for training, the gradient all-reduce that
``DistributedDataParallel`` adds around the model.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()


def shard_pairs(global_batch: int, world: int, rank: int):
    """[start, end) of the frame pairs rank ``rank`` owns; sizes differ by at most one."""
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_grads(params, bucket_bytes: int = 25 << 20):
    """What DistributedDataParallel adds around the model (/root/reference/exps/train_utils/double_trainer.py:171): the MEAN
    of every parameter gradient over the ranks.  Gradients are packed into flat fp32 buckets of ~``bucket_bytes`` in reverse
    parameter order (the order in which the backward walk finishes them), one all-reduce per bucket (NCCL over
    NVLink / NVSwitch on GPUs, gloo in the CPU tests), then unpacked in place.  BatchNorm buffers are not exchanged
    (``broadcast_buffers=False``).  No-op in a single process."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    todo = [p for p in reversed(list(params)) if p.grad is not None]
    n_buckets, i = 0, 0
    while i < len(todo):
        bucket, size = [], 0
        while i < len(todo) and (not bucket or size + todo[i].grad.numel() * 4 <= bucket_bytes):
            bucket.append(todo[i])
            size += todo[i].grad.numel() * 4
            i += 1
        flat = torch.cat([p.grad.detach().float().reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        n_buckets += 1
    return n_buckets
