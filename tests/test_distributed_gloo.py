"""CPU, world_size 2 over gloo: the one exchange step of the sharded path (pieces of cube outputs to the z-slab
owners) plus the histogram all-reduce, checked against the oracle's sequential blend of ALL cubes."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import assemble, geometry as ogeo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cube_output(geo, index):
    rng = np.random.default_rng(1000 + index)
    return rng.random((geo.roi,) * 3, dtype=np.float32)


def _blend_from_pieces(recv, off, z0, geo, s0, s1):
    """numpy statement of nc_blend_gather_f32's contract on the received pieces."""
    r, st = geo.roi, geo.step
    out = np.zeros((s1 - s0, geo.padded[1], geo.padded[2]), dtype=np.float32)
    cnt = np.zeros_like(out)
    for cube in range(geo.n_cubes):                       # ascending cube index = reference order
        if off[cube] < 0:
            continue
        oz, oy, ox = geo.origin(cube)
        a, b = max(oz, s0), min(oz + r, s1)
        if a >= b:
            continue
        n_planes = b - a
        piece = recv[off[cube]: off[cube] + n_planes * r * r].reshape(n_planes, r, r)
        assert a - oz == z0[cube]
        out[a - s0:b - s0, oy:oy + r, ox:ox + r] += piece / 8
        cnt[a - s0:b - s0, oy:oy + r, ox:ox + r] += 1
    return (out / cnt) * 8


def _worker(rank, world, port, size, roi, ov, bc, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from neuroclear_b200 import sharding
        geo = ogeo.dice_geometry(size, roi, ov, bc)
        cubes = sharding.balanced_ranges(geo.n_cubes, world)
        slabs = sharding.balanced_ranges(geo.padded[0], world)
        plan = sharding.plan_pieces(geo, cubes, slabs)
        c0, c1 = cubes[rank]
        queue = torch.from_numpy(np.stack([_cube_output(geo, i) for i in range(c0, c1)]))
        recv = sharding.exchange_pieces(queue, c0, geo, plan, rank)
        off, z0, sizes, total = sharding.piece_tables(geo, plan, rank, "cpu")
        assert recv.numel() == total
        s0, s1 = slabs[rank]
        mine = _blend_from_pieces(recv.numpy(), off.numpy(), z0.numpy(), geo, s0, s1)
        full, _ = assemble.blend_sequential([_cube_output(geo, i) for i in range(geo.n_cubes)], geo)
        assert np.array_equal(mine, full[s0:s1]), "sharded blend differs from the sequential reference order"
        # histogram all-reduce used by the percentile select
        hist = torch.bincount(torch.from_numpy((mine.reshape(-1) * 15).astype(np.int64)), minlength=16)
        dist.all_reduce(hist)
        ref = np.bincount((full.reshape(-1) * 15).astype(np.int64), minlength=16)
        assert np.array_equal(hist.numpy(), ref)
        open(os.path.join(result_dir, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_and_blend(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), (31, 40, 27), 12, 3, 2, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_three_rank_uneven(tmp_path):
    world = 3
    mp.spawn(_worker, args=(world, _free_port(), (50, 20, 33), 16, 4, 1, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def _grad_worker(rank, world, port, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from neuroclear_b200.apollo_d_path import allreduce_mean_gradients
        g = torch.Generator().manual_seed(0)
        shapes = [(64, 1, 4, 4), (64,), (128, 64, 4, 4), (7,)]
        params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
        all_grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(world)]   # same on every rank
        for p, gr in zip(params[:3], all_grads[rank][:3]):      # the last parameter has no gradient on any rank
            p.grad = gr.clone()
        allreduce_mean_gradients(params)
        for i, p in enumerate(params[:3]):
            want = sum(all_grads[r][i] for r in range(world)) / world
            assert torch.allclose(p.grad, want, atol=1e-6)
        assert params[3].grad is None
        open(os.path.join(result_dir, "gok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_allreduce(tmp_path):
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("gok%d" % r)).exists() for r in range(world))


def _digest_worker(rank, world, port, shape, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        vol = np.random.default_rng(3).integers(0, 65536, shape, dtype=np.uint16)
        z0, z1 = shape[0] * rank // world, shape[0] * (rank + 1) // world       # any partition into plane ranges
        if world == 3 and rank == 1:
            z1 = z0                                                            # a rank may own no output plane at all
        if world == 3 and rank == 2:
            z0 = shape[0] // 3
        digest, planes = bench.volume_digest(vol[z0:z1], z0, rank, world, torch.device("cpu"))
        if rank == 0:
            with open(out_path, "w") as f:
                f.write("%s %d" % (digest, planes))
    finally:
        dist.destroy_process_group()


def test_output_digest_does_not_depend_on_the_partition(tmp_path):
    """bench.py's out_sha256 (the bit-exact multi-GPU claim): the digest of the assembled volume is the same however
    the z-planes are split over ranks — world sizes 1, 2 and 3 (with an empty rank) over gloo."""
    shape = (23, 17, 31)
    got = []
    for world in (1, 2, 3):
        out = str(tmp_path / ("digest_%d.txt" % world))
        mp.spawn(_digest_worker, args=(world, _free_port(), shape, out), nprocs=world, join=True)
        got.append(open(out).read())
    assert got[0] == got[1] == got[2] and got[0].endswith(" 23")
    import hashlib
    vol = np.random.default_rng(3).integers(0, 65536, shape, dtype=np.uint16)
    h = hashlib.sha256()
    for z in range(shape[0]):
        h.update(hashlib.sha256(vol[z].tobytes()).digest())
    assert got[0].split()[0] == h.hexdigest()
