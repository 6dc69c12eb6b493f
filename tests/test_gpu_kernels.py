"""GPU: every kernel of the hot path, called through the C ABI, against the CPU oracle / golden fixtures."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import assemble, dice, geometry as ogeo, mip as omip

pytestmark = pytest.mark.gpu


def _fixture():
    return np.load(os.path.join(GOLDEN, "dice_assemble_31x40x27.npz"))


# ------------------------------------------------------------------------------------------------ dice
def test_dice_extract_golden_bit_exact(cuda):
    from neuroclear_b200.dicing import dice_extract, dice_geometry
    f = _fixture()
    roi, ov, bc = (int(v) for v in f["params"])
    vol = torch.from_numpy(f["volume"]).to(cuda)
    g = dice_geometry(f["volume"].shape, roi, ov, bc)
    got = dice_extract(vol, 0, g, 0, g.n_cubes).cpu().numpy()
    assert np.array_equal(got[:, None], f["cubes"])          # reference DiceImageDataSet output, bit for bit


@pytest.mark.parametrize("size,roi,ov,bc", [((20, 33, 25), 12, 3, 2), ((9, 9, 9), 8, 2, 1), ((130, 70, 141), 120, 15, 10),
                                            ((1, 5, 300), 16, 4, 3)])
def test_dice_extract_vs_oracle_ragged(cuda, size, roi, ov, bc):
    from neuroclear_b200.dicing import dice_extract, dice_geometry
    rng = np.random.default_rng(7)
    vol = rng.integers(0, 65536, size, dtype=np.uint16)
    g, og = dice_geometry(size, roi, ov, bc), ogeo.dice_geometry(size, roi, ov, bc)
    dev = torch.from_numpy(vol).to(cuda)
    for begin, count in [(0, min(3, g.n_cubes)), (g.n_cubes - 1, 1)]:
        got = dice_extract(dev, 0, g, begin, count).cpu().numpy()
        for k in range(count):
            assert np.array_equal(got[k][None], dice.dice_cube_gather(vol, og, begin + k))


def test_dice_extract_from_z_slab(cuda):
    """Sharded input: a rank holds only the planes sharding.input_plane_range says it needs."""
    from neuroclear_b200 import sharding
    from neuroclear_b200.dicing import dice_extract, dice_geometry
    rng = np.random.default_rng(3)
    size = (70, 30, 28)
    vol = rng.integers(0, 65536, size, dtype=np.uint16)
    g, og = dice_geometry(size, 12, 3, 2), ogeo.dice_geometry(size, 12, 3, 2)
    for c0, c1 in sharding.balanced_ranges(g.n_cubes, 3):
        z0, z1 = sharding.input_plane_range(g, c0, c1)
        dev = torch.from_numpy(vol[z0:z1].copy()).to(cuda)
        got = dice_extract(dev, z0, g, c0, c1 - c0).cpu().numpy()
        for k in (0, c1 - c0 - 1):
            assert np.array_equal(got[k][None], dice.dice_cube_gather(vol, og, c0 + k))


# ------------------------------------------------------------------------------------------------ assembly
def test_blend_percentile_rescale_golden_bit_exact(cuda):
    from neuroclear_b200.dicing import PercentileSelect, blend_gather, dice_geometry, rescale_u16_crop
    f = _fixture()
    roi, ov, bc = (int(v) for v in f["params"])
    g = dice_geometry(f["volume"].shape, roi, ov, bc)
    cubes = np.stack([assemble.crop_border(c, bc) for c in f["fake"]])
    q = torch.from_numpy(cubes).to(cuda)
    off = torch.arange(g.n_cubes, dtype=torch.int64, device=cuda) * roi ** 3
    z0 = torch.zeros(g.n_cubes, dtype=torch.int32, device=cuda)
    vis = blend_gather(q.view(-1), off, z0, g, 0, g.padded[0])
    assert np.array_equal(vis.cpu().numpy(), f["blend"])                        # reference (sum/mask)*8, bit for bit
    plain = rescale_u16_crop(vis, 0, g, None, 0, g.size[0]).cpu().numpy()
    assert np.array_equal(plain, f["final_plain"])
    sel = PercentileSelect(cuda)
    norm3, p64 = sel.run(vis, vis.numel(), (0.25, 99.75))
    assert tuple(p64.cpu().tolist()) == tuple(f["pcts"])                        # np.percentile float64, exactly
    normd = rescale_u16_crop(vis, 0, g, norm3, 0, g.size[0]).cpu().numpy()
    assert np.array_equal(normd, f["final_norm"])


def test_blend_slabs_equal_whole(cuda):
    from neuroclear_b200.dicing import blend_gather, dice_geometry
    rng = np.random.default_rng(11)
    g = dice_geometry((50, 20, 33), 16, 4, 1)
    og = ogeo.dice_geometry((50, 20, 33), 16, 4, 1)
    cubes = rng.random((g.n_cubes, 16, 16, 16), dtype=np.float32)
    ref, _ = assemble.blend_sequential(list(cubes), og)
    q = torch.from_numpy(cubes).to(cuda)
    off = torch.arange(g.n_cubes, dtype=torch.int64, device=cuda) * 16 ** 3
    z0 = torch.zeros(g.n_cubes, dtype=torch.int32, device=cuda)
    whole = blend_gather(q.view(-1), off, z0, g, 0, g.padded[0]).cpu().numpy()
    assert np.array_equal(whole, ref)
    part = blend_gather(q.view(-1), off, z0, g, 13, 21).cpu().numpy()
    assert np.array_equal(part, ref[13:34])


@pytest.mark.parametrize("n,dist", [(1 << 20, "uniform"), (999_983, "sigmoid"), (4096, "ties"), (5, "uniform")])
def test_percentile_select_exact(cuda, n, dist):
    from neuroclear_b200.dicing import PercentileSelect
    rng = np.random.default_rng(5)
    if dist == "uniform":
        a = rng.random(n, dtype=np.float32)
    elif dist == "sigmoid":
        a = (1 / (1 + np.exp(-rng.normal(0, 0.05, n)))).astype(np.float32)
    else:
        a = rng.integers(0, 4, n).astype(np.float32) / 4
    buf = torch.empty(n + 4, dtype=torch.float32, device=cuda)[:n]            # 16-byte aligned base
    buf.copy_(torch.from_numpy(a))
    for sat in [(0.25, 99.75), (0.0, 100.0), (50.0, 50.0)]:
        _, p64 = PercentileSelect(cuda).run(buf, n, sat)
        ref = np.percentile(a, sat)
        assert tuple(p64.cpu().tolist()) == (float(ref[0]), float(ref[1]))


def test_rescale_degenerate_percentile_range_follows_skimage(cuda):
    """imin == imax (e.g. a flat volume, or sat_level (50, 50)): skimage clips to [imin, imax] FIRST, so every voxel
    becomes imin, and only then to [0, 1]; the branch is chosen on the float64 percentiles."""
    from neuroclear_b200.dicing import PercentileSelect, dice_geometry, rescale_u16_crop
    from oracle import assemble, geometry as ogeo
    rng = np.random.default_rng(9)
    size = (9, 9, 9)
    g, og = dice_geometry(size, 8, 2, 1), ogeo.dice_geometry(size, 8, 2, 1)
    vis = rng.random(g.padded, dtype=np.float32)
    for sat in [(50.0, 50.0), (0.25, 99.75)]:
        ref, _ = assemble.finish(vis, og, True, sat_level=sat)
        dv = torch.from_numpy(vis).to(cuda)
        norm3, _ = PercentileSelect(cuda).run(dv.view(-1), dv.numel(), sat)
        got = rescale_u16_crop(dv, 0, g, norm3, 0, size[0]).cpu().numpy()
        assert np.array_equal(got, ref), sat
        if sat[0] == sat[1]:
            assert len(np.unique(got)) == 1


# ------------------------------------------------------------------------------------------------ network layers
def _ndhwc(t):      # (N,C,D,H,W) -> (N,D,H,W,C) contiguous
    return t.permute(0, 2, 3, 4, 1).contiguous()


def _finalize(lib, st, cin, nb, d, h, w, c, cuda):
    from neuroclear_b200._lib import call, i64, ptr, stream_ptr
    rows = lib.nc_conv3d_k3_stats_rows(cin, nb, d, h, w, c)
    mr = torch.empty(nb * 2 * c, dtype=torch.float32, device=cuda)
    scratch = torch.zeros(lib.nc_in_stats_scratch_bytes(nb, c), dtype=torch.uint8, device=cuda)
    for _ in range(2):          # the arrival counters reset themselves: a second launch must give the same answer
        call("nc_in_stats_finalize", ptr(st), nb, i64(rows // nb), c, i64(d * h * w), 1e-5, ptr(scratch), ptr(mr),
             stream_ptr())
    return mr.view(nb, 2, c)


def test_first_layer_conv_and_stats(cuda, lib):
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(0)
    nb, d, h, w = 2, 12, 20, 36
    x = torch.rand((nb, 1, d, h, w), generator=g)
    wt = torch.randn((64, 1, 3, 3, 3), generator=g) * 0.3
    ref = F.conv3d(x, wt, padding=1)
    rows = lib.nc_conv3d_k3_stats_rows(1, nb, d, h, w, 64)
    y = torch.empty((nb, d, h, w, 64), dtype=torch.float16, device=cuda)
    st = torch.empty(rows * 2 * 64, device=cuda)
    xd, wd = x.to(cuda).contiguous(), wt.to(cuda).reshape(64, 27).contiguous()   # keep alive across the launch
    packed = torch.empty(8192, dtype=torch.uint8, device=cuda)
    call("nc_pack_weights_conv3d_cin1_k3", ptr(wd), ptr(packed), stream_ptr())
    call("nc_conv3d_cin1_k3_fwd", ptr(xd), ptr(packed), nb, d, h, w, 64, ptr(y), ptr(st), stream_ptr())
    ref_h = F.conv3d(x, wt.half().float(), padding=1)      # fp16 weights, hi+lo split input (~fp32), fp16 storage
    assert torch.allclose(y.cpu().float(), _ndhwc(ref_h), atol=2e-3, rtol=2e-3)
    mr = _finalize(lib, st, 1, nb, d, h, w, 64, cuda).cpu()
    mean = ref.mean(dim=(2, 3, 4))
    rstd = 1 / torch.sqrt(ref.var(dim=(2, 3, 4), unbiased=False) + 1e-5)
    assert torch.allclose(mr[:, 0], mean, atol=1e-3) and torch.allclose(mr[:, 1], rstd, rtol=2e-3)


@pytest.mark.parametrize("cin,cout,nb,d,h,w", [(64, 64, 1, 7, 20, 12), (64, 128, 2, 6, 10, 18), (128, 128, 1, 5, 17, 9),
                                               (256, 256, 1, 4, 12, 12), (256, 128, 1, 6, 18, 10), (128, 64, 1, 8, 33, 17),
                                               (128, 128, 1, 6, 22, 20), (128, 256, 2, 9, 35, 35),
                                               (256, 128, 1, 7, 6, 9), (64, 128, 1, 5, 70, 12)])
def test_conv3d_k3_tensor_core(cuda, lib, cin, cout, nb, d, h, w):
    """fp16 operands, fp32 accumulate: compare with F.conv3d on the SAME fp16-rounded operands in fp32."""
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn((nb, cin, d, h, w), generator=g).half()
    wt = (torch.randn((cout, cin, 3, 3, 3), generator=g) * (2.0 / (27 * cin)) ** 0.5)
    ref = F.conv3d(x.float(), wt.half().float(), padding=1)
    packed = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 0), dtype=torch.uint8, device=cuda)
    wd, xd = wt.to(cuda).contiguous(), _ndhwc(x).to(cuda)                        # keep alive across the launches
    call("nc_pack_weights_conv3d_k3", ptr(wd), cout, cin, ptr(packed), stream_ptr())
    rows = lib.nc_conv3d_k3_stats_rows(cin, nb, d, h, w, cout)
    y = torch.full((nb, d, h, w, cout), float("nan"), dtype=torch.float16, device=cuda)
    st = torch.empty(rows * 2 * cout, device=cuda)
    call("nc_conv3d_k3_fwd", ptr(xd), None, nb, d, h, w, cin, ptr(packed), cout, ptr(y), ptr(st), stream_ptr())
    got = y.cpu().float()
    assert torch.isfinite(got).all()
    assert (got - _ndhwc(ref)).abs().max() <= 2e-3 * max(1.0, ref.abs().max().item())   # fp16 storage
    mr = _finalize(lib, st, cin, nb, d, h, w, cout, cuda).cpu()
    assert torch.allclose(mr[:, 0], ref.mean(dim=(2, 3, 4)), atol=2e-4)
    assert torch.allclose(mr[:, 1], 1 / torch.sqrt(ref.var(dim=(2, 3, 4), unbiased=False) + 1e-5), rtol=2e-3)


@pytest.mark.parametrize("cin,cout,d,h,w", [(256, 128, 3, 9, 5), (128, 64, 4, 18, 10)])
def test_conv_transpose_into_concat_slice(cuda, lib, cin, cout, d, h, w):
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(cin)
    x = torch.randn((1, cin, d, h, w), generator=g).half()
    wt = torch.randn((cin, cout, 2, 2, 2), generator=g) * (1.0 / cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv_transpose3d(x.float(), wt.half().float(), b, stride=2)
    packed = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 1), dtype=torch.uint8, device=cuda)
    wd, xd, bd = wt.to(cuda).contiguous(), _ndhwc(x).to(cuda), b.to(cuda)        # keep alive across the launches
    call("nc_pack_weights_convT3d_k2s2", ptr(wd), cin, cout, ptr(packed), stream_ptr())
    cat = torch.full((1, 2 * d, 2 * h, 2 * w, 2 * cout), 7.0, dtype=torch.float16, device=cuda)
    call("nc_convT3d_k2s2_fwd", ptr(xd), None, 1, d, h, w, cin, ptr(packed), ptr(bd), cout, ptr(cat), 2 * cout, cout,
         stream_ptr())
    got = cat.cpu().float()
    assert (got[..., :cout] == 7.0).all()                                     # the skip half is untouched
    assert (got[..., cout:] - _ndhwc(ref)).abs().max() <= 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("pool", [False, True])
def test_instance_norm_relu_pool_apply(cuda, lib, pool):
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(2)
    nb, c, d, h, w = 2, 64, 4, 6, 8
    raw = (torch.randn((nb, c, d, h, w), generator=g) * 3 + 1).half().float()     # raw conv outputs are stored fp16
    mean = raw.mean(dim=(2, 3, 4))
    rstd = 1 / torch.sqrt(raw.var(dim=(2, 3, 4), unbiased=False) + 1e-5)
    mr = torch.stack([mean, rstd], 1).contiguous().to(cuda)
    ref = F.relu((raw - mean[:, :, None, None, None]) * rstd[:, :, None, None, None])
    y = torch.zeros((nb, d, h, w, 2 * c), dtype=torch.float16, device=cuda)
    pooled = torch.zeros((nb, d // 2, h // 2, w // 2, c), dtype=torch.float16, device=cuda) if pool else None
    rawd = _ndhwc(raw).half().to(cuda)
    call("nc_in_relu_apply", ptr(rawd), ptr(mr), nb, d, h, w, c, ptr(y), 2 * c, c, ptr(pooled), stream_ptr())
    got = y.cpu().float()
    assert (got[..., :c] == 0).all()
    assert torch.equal(got[..., c:], _ndhwc(ref).half().float())          # same fp32 expression, RN to fp16
    if pool:
        assert torch.equal(pooled.cpu().float(), _ndhwc(F.max_pool3d(ref, 2)).half().float())


def test_head_1x1_sigmoid_with_border_cut(cuda, lib):
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(4)
    nb, c, d, h, w, crop = 2, 64, 10, 12, 14, 3
    raw = torch.randn((nb, c, d, h, w), generator=g).half().float()
    mean = raw.mean(dim=(2, 3, 4))
    rstd = 1 / torch.sqrt(raw.var(dim=(2, 3, 4), unbiased=False) + 1e-5)
    w1, b1 = torch.randn(c, generator=g) * 0.2, torch.tensor(0.05)
    w2, b2 = torch.tensor(1.7), torch.tensor(-0.1)
    act = F.relu(F.instance_norm(raw, eps=1e-5))
    ref = torch.sigmoid(w2 * ((act * w1[None, :, None, None, None]).sum(1) + b1) + b2)
    hp = torch.cat([w1, b1[None], w2[None], b2[None]]).to(cuda)
    mr = torch.stack([mean, rstd], 1).contiguous().to(cuda)
    for cr in (0, crop):
        y = torch.empty((nb, d - 2 * cr, h - 2 * cr, w - 2 * cr), device=cuda)
        rawd = _ndhwc(raw).half().to(cuda)
        call("nc_head_1x1_sigmoid_fwd", ptr(rawd), ptr(mr), ptr(hp), nb, d, h, w, c, cr, ptr(y), stream_ptr())
        want = ref[:, cr:d - cr, cr:h - cr, cr:w - cr]
        assert (y.cpu() - want).abs().max() <= 1e-5


# ------------------------------------------------------------------------------------------------ projection
def test_mip_forward_backward(cuda, lib):
    from neuroclear_b200.projection import Volume
    f = np.load(os.path.join(GOLDEN, "mip_12.npz"))
    vol = torch.from_numpy(f["vol"]).to(cuda)
    np.random.seed(5)
    state = np.random.get_state()
    for axis in range(3):
        np.random.set_state(state)
        got = Volume(vol, cuda).get_projection(4, axis)
        assert got.shape == (1, 1, 12, 12) and np.array_equal(got.cpu().numpy(), f[f"proj{axis}"])   # reference output
    # gradient routing = torch.max(dim)[0] autograd on the CPU
    np.random.set_state(state)
    v = vol.clone().requires_grad_(True)
    p = Volume(v, cuda).get_projection(5, 1)
    gout = torch.rand(p.shape, generator=torch.Generator().manual_seed(9)).to(cuda)
    p.backward(gout)
    np.random.set_state(state)
    vc = torch.from_numpy(f["vol"]).requires_grad_(True)
    pc, _ = omip.get_projection(vc, 5, 1)
    pc.backward(gout.cpu())
    assert torch.equal(v.grad.cpu(), vc.grad)


def test_stats_finalize_shared_scratch_across_channel_counts(cuda, lib):
    """One scratch buffer serves calls with different (NB, C), as UnetDeconvEngine does layer after layer."""
    from neuroclear_b200._lib import call, i64, ptr, stream_ptr
    g = torch.Generator().manual_seed(3)
    scratch = torch.zeros(lib.nc_in_stats_scratch_bytes(3, 256), dtype=torch.uint8, device=cuda)
    for rep in range(2):
        for nb, c, rows in [(3, 64, 500), (2, 256, 77), (1, 128, 1000), (3, 64, 9)]:
            part = torch.rand((nb, rows, 2, c), generator=g) + 0.5
            pd = part.to(cuda)
            mr = torch.empty((nb, 2, c), device=cuda)
            n = rows * 384
            call("nc_in_stats_finalize", ptr(pd), nb, i64(rows), c, i64(n), 1e-5, ptr(scratch), ptr(mr), stream_ptr())
            s1, s2 = part[:, :, 0].double().sum(1), part[:, :, 1].double().sum(1)
            mean = s1 / n
            var = (s2 / n - mean * mean).clamp_min(0)
            got = mr.cpu().double()
            assert torch.allclose(got[:, 0], mean, rtol=1e-6) and torch.allclose(got[:, 1], 1 / torch.sqrt(var + 1e-5), rtol=1e-5)


@pytest.mark.parametrize("cin,cout,nb,d,h,w", [(64, 64, 2, 9, 20, 12), (128, 128, 1, 5, 17, 9), (256, 256, 1, 4, 12, 12),
                                               (128, 128, 2, 7, 22, 19), (256, 256, 1, 9, 35, 35)])
def test_conv3d_k3_fused_instance_norm_input(cuda, lib, cin, cout, nb, d, h, w):
    """in_mean_rstd != NULL: the kernel normalises + ReLUs the RAW input planes in shared memory.  Must equal the
    two-pass path (nc_in_relu_apply, then conv) BIT FOR BIT: same fp32 expression, same fp16 rounding."""
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(cin * 3 + cout)
    raw = (torch.randn((nb, d, h, w, cin), generator=g) * 2 + 0.5).half().to(cuda)
    mean = raw.float().mean(dim=(1, 2, 3))
    rstd = 1 / torch.sqrt(raw.float().var(dim=(1, 2, 3), unbiased=False) + 1e-5)
    mr = torch.stack([mean, rstd], 1).contiguous()
    wt = (torch.randn((cout, cin, 3, 3, 3), generator=g) * (2.0 / (27 * cin)) ** 0.5).to(cuda)
    packed = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 0), dtype=torch.uint8, device=cuda)
    call("nc_pack_weights_conv3d_k3", ptr(wt), cout, cin, ptr(packed), stream_ptr())
    rows = lib.nc_conv3d_k3_stats_rows(cin, nb, d, h, w, cout)
    outs = []
    for fused in (False, True):
        y = torch.full((nb, d, h, w, cout), float("nan"), dtype=torch.float16, device=cuda)
        st = torch.zeros(rows * 2 * cout, device=cuda)
        if fused:
            call("nc_conv3d_k3_fwd", ptr(raw), ptr(mr), nb, d, h, w, cin, ptr(packed), cout, ptr(y), ptr(st), stream_ptr())
        else:
            act = torch.empty_like(raw)
            call("nc_in_relu_apply", ptr(raw), ptr(mr), nb, d, h, w, cin, ptr(act), cin, 0, None, stream_ptr())
            call("nc_conv3d_k3_fwd", ptr(act), None, nb, d, h, w, cin, ptr(packed), cout, ptr(y), ptr(st), stream_ptr())
        torch.cuda.synchronize()
        outs.append((y.clone(), st.clone()))
    assert torch.isfinite(outs[1][0].float()).all()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_conv_transpose_fused_instance_norm_input(cuda, lib):
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(21)
    cin, cout, d, h, w = 128, 64, 4, 18, 10
    raw = (torch.randn((1, d, h, w, cin), generator=g) * 2 - 0.3).half().to(cuda)
    mean = raw.float().mean(dim=(1, 2, 3))
    rstd = 1 / torch.sqrt(raw.float().var(dim=(1, 2, 3), unbiased=False) + 1e-5)
    mr = torch.stack([mean, rstd], 1).contiguous()
    wt = (torch.randn((cin, cout, 2, 2, 2), generator=g) * (1.0 / cin) ** 0.5).to(cuda)
    b = (torch.randn(cout, generator=g) * 0.1).to(cuda)
    packed = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 1), dtype=torch.uint8, device=cuda)
    call("nc_pack_weights_convT3d_k2s2", ptr(wt), cin, cout, ptr(packed), stream_ptr())
    outs = []
    for fused in (False, True):
        cat = torch.zeros((1, 2 * d, 2 * h, 2 * w, 2 * cout), dtype=torch.float16, device=cuda)
        if fused:
            call("nc_convT3d_k2s2_fwd", ptr(raw), ptr(mr), 1, d, h, w, cin, ptr(packed), ptr(b), cout, ptr(cat), 2 * cout,
                 cout, stream_ptr())
        else:
            act = torch.empty_like(raw)
            call("nc_in_relu_apply", ptr(raw), ptr(mr), 1, d, h, w, cin, ptr(act), cin, 0, None, stream_ptr())
            call("nc_convT3d_k2s2_fwd", ptr(act), None, 1, d, h, w, cin, ptr(packed), ptr(b), cout, ptr(cat), 2 * cout,
                 cout, stream_ptr())
        torch.cuda.synchronize()
        outs.append(cat.clone())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("cin,cout", [(64, 64), (128, 128), (128, 64)])
def test_conv3d_persistent_grid_cap_is_bit_identical(cuda, lib, cin, cout):
    """nc_debug_set_max_ctas(2) makes every CTA walk many tiles (plane / weight ring wrap-around, TMEM ping-pong):
    the result must not change by a bit."""
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(cin + 7 * cout)
    nb, d, h, w = 2, 10, 21, 19
    x = torch.randn((nb, d, h, w, cin), generator=g).half().to(cuda)
    wt = (torch.randn((cout, cin, 3, 3, 3), generator=g) * (2.0 / (27 * cin)) ** 0.5).to(cuda)
    packed = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 0), dtype=torch.uint8, device=cuda)
    call("nc_pack_weights_conv3d_k3", ptr(wt), cout, cin, ptr(packed), stream_ptr())
    rows = lib.nc_conv3d_k3_stats_rows(cin, nb, d, h, w, cout)
    outs = []
    try:
        for cap in (0, 2, 5):
            lib.nc_debug_set_max_ctas(cap)
            y = torch.zeros((nb, d, h, w, cout), dtype=torch.float16, device=cuda)
            st = torch.zeros(rows * 2 * cout, device=cuda)
            call("nc_conv3d_k3_fwd", ptr(x), None, nb, d, h, w, cin, ptr(packed), cout, ptr(y), ptr(st), stream_ptr())
            torch.cuda.synchronize()
            outs.append((y, st))
    finally:
        lib.nc_debug_set_max_ctas(0)
    for y, st in outs[1:]:
        assert torch.equal(y, outs[0][0]) and torch.equal(st, outs[0][1])


@pytest.mark.parametrize("cin,cout,nb,d,h,w,fused", [(128, 128, 2, 7, 22, 19, False), (256, 256, 1, 9, 35, 35, True),
                                                     (64, 128, 1, 6, 70, 17, False), (128, 128, 1, 5, 5, 9, True)])
def test_remainder_pair_kernel_is_bit_identical_to_the_regular_tiles(cuda, lib, cin, cout, nb, d, h, w, fused):
    """H mod 16 in 1..6: the remainder strip of two consecutive planes shares one accumulator (conv3d_rp_kernel).
    Same K order, same epilogue: the raw fp16 outputs must equal the regular kernel's (hook off) bit for bit; the
    statistics are the same sums split over different tiles, so mean / rstd agree to fp32 rounding."""
    from neuroclear_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(cin + cout + h)
    x = (torch.randn((nb, d, h, w, cin), generator=g) * 1.5 + 0.2).half().to(cuda)
    mr_in = None
    if fused:
        mean = x.float().mean(dim=(1, 2, 3))
        rstd = 1 / torch.sqrt(x.float().var(dim=(1, 2, 3), unbiased=False) + 1e-5)
        mr_in = torch.stack([mean, rstd], 1).contiguous()
    wt = (torch.randn((cout, cin, 3, 3, 3), generator=g) * (2.0 / (27 * cin)) ** 0.5).to(cuda)
    packed = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 0), dtype=torch.uint8, device=cuda)
    call("nc_pack_weights_conv3d_k3", ptr(wt), cout, cin, ptr(packed), stream_ptr())
    res = []
    try:
        for on in (1, 0):
            lib.nc_debug_set_remainder_pairs(on)
            rows = lib.nc_conv3d_k3_stats_rows(cin, nb, d, h, w, cout)
            y = torch.full((nb, d, h, w, cout), float("nan"), dtype=torch.float16, device=cuda)
            st = torch.full((rows * 2 * cout,), float("nan"), device=cuda)
            call("nc_conv3d_k3_fwd", ptr(x), ptr(mr_in), nb, d, h, w, cin, ptr(packed), cout, ptr(y), ptr(st), stream_ptr())
            mr = _finalize(lib, st, cin, nb, d, h, w, cout, cuda)
            torch.cuda.synchronize()
            res.append((y, mr.clone(), rows))
    finally:
        lib.nc_debug_set_remainder_pairs(1)
    assert res[0][2] < res[1][2]                                   # fewer tiles with the remainder pairs
    assert torch.isfinite(res[0][0].float()).all() and torch.equal(res[0][0], res[1][0])
    assert torch.allclose(res[0][1], res[1][1], rtol=1e-5, atol=1e-6)
