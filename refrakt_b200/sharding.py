"""Host rules of the multi-GPU runs, one process per GPU. The multi-GPU layer itself is C++ (csrc/comm.cpp, the rfk_comm_*
entry points of include/refrakt_b200.h: NCCL + peer memory, rfk_render_frame_sharded); this module only states the two
rules the launchers share — which seeds a rank uses and which frames of an animation it renders — and mirrors the row-slab
geometry of the C ABI for the tests."""
from __future__ import annotations

from typing import List, Tuple


def rank_seed(rank: int, total_particles: int, base_seed: int = 0) -> int:
    """Rank g seeds particle slots [g*P, (g+1)*P) of one global JSF32 seed sequence: disjoint streams (what rfk_render --world does)."""
    return base_seed + rank * total_particles


def frames_of_rank(num_frames: int, world: int, rank: int) -> List[int]:
    """Frame-parallel animation (BASELINE configs[3], rfk_render --frame-parallel): frame f goes to rank f % world."""
    return list(range(rank, num_frames, world))


def row_slab(height: int, halo: int, rank: int, world: int) -> Tuple[int, int, int, int]:
    """rfk_comm_row_slab: (y0, y1, src_y0, src_y1) — the output rows of `rank` and the source rows their density estimation reads"""
    from . import comm_row_slab
    s = comm_row_slab(height, halo, rank, world)
    return int(s.y0), int(s.y1), int(s.src_y0), int(s.src_y1)
