"""Multi-GPU sharding of the diced-inference path: pure host logic, device-agnostic tensors.

* cubes (independent units: InstanceNorm statistics are per cube) are split into contiguous, balanced index
  ranges — one range per rank, no collective on the network path;
* the assembled (padded) volume is split into balanced z-slabs, one per rank;
* ONE exchange step: every rank ships, to each slab owner, the z-plane range of each of its cube outputs that
  intersects that slab (point-to-point over NCCL/NVLink; gloo on CPU in the tests).  The slab owner then runs the
  gather-blend in ascending cube index, which reproduces the reference's sequential fp32 accumulation
  (util/assemble_dice.py:167-173) bit for bit regardless of how many GPUs computed the cubes;
* the percentile histogram is all-reduced (3 x 128 KB), nothing else crosses GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def balanced_ranges(n: int, world: int):
    """Contiguous ranges [n*r//world, n*(r+1)//world)."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def cube_z_extent(geo, index: int):
    """Padded-volume z range [z0, z0+roi) covered by the (border-cut) output of cube `index`."""
    nz, ny, nx = geo.steps
    cz = index // (nx * ny)
    return cz * geo.step, cz * geo.step + geo.roi


def input_plane_range(geo, cube_lo: int, cube_hi: int):
    """Original-volume planes [z0, z1) a rank must hold to dice cubes [cube_lo, cube_hi): the cubes' z extent
    plus the border, mapped through numpy-'reflect' at both ends of the padded volume, clipped to the data."""
    if cube_hi <= cube_lo:
        return 0, 0
    nz, ny, nx = geo.steps
    cz_lo, cz_hi = cube_lo // (nx * ny), (cube_hi - 1) // (nx * ny)
    pz = geo.padded[0]
    lo_j = cz_lo * geo.step - geo.border
    hi_j = cz_hi * geo.step + geo.roi + geo.border - 1          # inclusive, before reflection
    min_p, max_p = max(lo_j, 0), min(hi_j, pz - 1)
    if lo_j < 0:
        max_p = max(max_p, -lo_j)                               # -i -> i
    if hi_j >= pz:
        min_p = min(min_p, 2 * (pz - 1) - hi_j)                 # pz-1+i -> pz-1-i
    z0, z1 = max(min_p, 0), min(max_p + 1, geo.size[0])
    return (z0, z1) if z1 > z0 else (0, 0)


def input_row_end(geo, cube_lo: int, cube_hi: int):
    """One past the last ORIGINAL-volume row (y) the cubes [cube_lo, cube_hi) read — their y extent plus the border;
    numpy-'reflect' at either end of the padded volume only maps onto rows below that bound, and rows beyond the data
    are zero padding.  (The y analogue of input_plane_range's upper bound; the lower bound is not needed: rows are
    uploaded in ascending order.)"""
    if cube_hi <= cube_lo:
        return 0
    nz, ny, nx = geo.steps
    lo, hi = cube_lo // nx, (cube_hi - 1) // nx               # row-of-cubes indices (z-major, then y)
    if lo // ny != hi // ny:                                   # the range wraps into the next z-layer: every row
        cy_max = ny - 1
    else:
        cy_max = hi % ny
    return min(geo.size[1], cy_max * geo.step + geo.roi + geo.border)


@dataclass
class Piece:
    cube: int   # global cube index
    p0: int     # first cube-local z plane shipped
    p1: int     # one past the last


def plan_pieces(geo, cube_ranges, slab_ranges):
    """plan[src][dst] = [Piece...] in ascending cube index."""
    world = len(cube_ranges)
    plan = [[[] for _ in range(world)] for _ in range(world)]
    for src, (c0, c1) in enumerate(cube_ranges):
        for cube in range(c0, c1):
            z0, z1 = cube_z_extent(geo, cube)
            for dst, (s0, s1) in enumerate(slab_ranges):
                a, b = max(z0, s0), min(z1, s1)
                if a < b:
                    plan[src][dst].append(Piece(cube, a - z0, b - z0))
    return plan


def piece_tables(geo, plan, dst: int, device):
    """Lookup tables for nc_blend_gather_f32 on rank `dst`: receive-buffer offset and first plane per cube."""
    off = torch.full((geo.n_cubes,), -1, dtype=torch.int64)
    z0 = torch.zeros((geo.n_cubes,), dtype=torch.int32)
    plane = geo.roi * geo.roi
    pos = 0
    sizes = []
    for src in range(len(plan)):
        n_src = 0
        for pc in plan[src][dst]:
            off[pc.cube] = pos
            z0[pc.cube] = pc.p0
            n = (pc.p1 - pc.p0) * plane
            pos += n
            n_src += n
        sizes.append(n_src)
    return off.to(device), z0.to(device), sizes, pos


def exchange_pieces(queue: torch.Tensor, cube_lo: int, geo, plan, rank: int, group=None) -> torch.Tensor:
    """queue: (n_local_cubes, roi, roi, roi) outputs of this rank's cubes.  Returns this rank's receive buffer
    (all pieces intersecting its slab, ordered by source rank = ascending cube index)."""
    world = len(plan)
    plane = geo.roi * geo.roi
    recv_sizes = [sum((pc.p1 - pc.p0) * plane for pc in plan[src][rank]) for src in range(world)]
    recv = torch.empty(sum(recv_sizes), dtype=queue.dtype, device=queue.device)
    recv_parts = list(torch.split(recv, recv_sizes))
    ops, keep = [], []
    for dst in range(world):
        pieces = plan[rank][dst]
        if not pieces:
            continue
        n = sum((pc.p1 - pc.p0) * plane for pc in pieces)
        target = recv_parts[rank] if dst == rank else torch.empty(n, dtype=queue.dtype, device=queue.device)
        pos = 0
        for pc in pieces:  # each piece is a contiguous plane range of one cube
            src_view = queue[pc.cube - cube_lo, pc.p0:pc.p1].reshape(-1)
            target[pos:pos + src_view.numel()].copy_(src_view)
            pos += src_view.numel()
        if dst != rank:
            keep.append(target)
            ops.append(dist.P2POp(dist.isend, target, dst, group))
    for src in range(world):
        if src != rank and recv_sizes[src] > 0:
            ops.append(dist.P2POp(dist.irecv, recv_parts[src], src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return recv
