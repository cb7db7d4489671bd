"""Multi-GPU host logic: one process per GPU (torchrun), independent particle streams into a
private histogram per rank, ONE exchange step — the histogram sum — before density estimation
(SURVEY.md §8e). Animation frames are distributed round-robin with no exchange at all.

torch.distributed is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple


def rank_seed(rank: int, total_particles: int, base_seed: int = 0) -> int:
    """Rank g seeds particle slots [g*P, (g+1)*P) of one global JSF32 seed sequence: disjoint streams."""
    return base_seed + rank * total_particles


def rank_iteration_share(total_iterations: int, world: int, rank: int) -> int:
    """Strong-scaling split of a fixed iteration budget: the first `rem` ranks take one unit more."""
    base, rem = divmod(total_iterations, world)
    return base + (1 if rank < rem else 0)


def frames_of_rank(num_frames: int, world: int, rank: int) -> List[int]:
    """Frame-parallel animation (BASELINE configs[3]): frame f goes to rank f % world."""
    return list(range(rank, num_frames, world))


def row_slabs(height: int, world: int, halo: int) -> List[Tuple[int, int, int, int]]:
    """Row slabs for a sharded density estimation: (row0, row1, halo_row0, halo_row1) per rank; a slab needs
    `halo` = estimator_radius source rows on each side."""
    out = []
    base, rem = divmod(height, world)
    row = 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((row, row + n, max(0, row - halo), min(height, row + n + halo)))
        row += n
    return out


def reduce_histogram(bins, dst: int = 0):
    """Sum the per-rank float4 histograms onto rank `dst` (NCCL reduce over NVLink; 16 B per bin per GPU)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(bins, dst=dst, op=dist.ReduceOp.SUM)
    return bins


def allreduce_histogram(bins):
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bins, op=dist.ReduceOp.SUM)
    return bins


def gather_counts(value: int) -> List[int]:
    """Per-rank binned-sample counts on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return [int(value)]
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [int(t.item()) for t in out]
