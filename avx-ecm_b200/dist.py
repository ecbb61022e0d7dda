"""Multi-GPU plumbing: one process per GPU, curves partitioned by sigma, no data-path collective.

Curves are independent, so rank r of `world` runs the contiguous sigma slice
[sigma0 + first, sigma0 + first + count) on its own context; the only communication is the final
host gather of (sigma, save line, factors) to rank 0, merged in sigma order so that the output
equals what the reference writes with threads=1 (ecm.c:1319-1388 order: batch, thread, lane).
torch.distributed is used for that gather and for the timing reductions only (NCCL on GPUs, gloo
in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced partition of `total` curves: -> (first, count) of this rank."""
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_max(x):
    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(x):
    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_results(local):
    """local: dict(sigmas=[...], save_lines=[...], factors=[(sigma, stage, factor)...]).
    Returns the merged dict on rank 0 (sigma order), None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()):
        return merge([local])
    world = dist.get_world_size()
    out = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(local, out, dst=0)
    return merge(out) if dist.get_rank() == 0 else None


def merge(parts):
    rows = []
    factors = []
    for p in parts:
        rows += list(zip(p["sigmas"], p["save_lines"]))
        factors += list(p["factors"])
    rows.sort(key=lambda r: r[0])
    factors.sort(key=lambda f: (f[1], f[0]))
    return {"sigmas": [r[0] for r in rows], "save_lines": [r[1] for r in rows], "factors": factors}


def run_sharded(total_curves, sigma0, compute):
    """compute(first_sigma, count) -> dict like vececm(); runs this rank's shard and gathers."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    first, count = shard_range(total_curves, rank, world)
    r = compute(sigma0 + first, count) if count else {"save_lines": [], "factors": []}
    local = {"sigmas": [sigma0 + first + i for i in range(count)], "save_lines": r["save_lines"], "factors": r["factors"]}
    return gather_results(local)
