"""CPU: the C-ABI library loads and exports everything the header declares; host-side integer logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import geometry as ogeo


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "neuroclear_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nc_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(lib):
    from neuroclear_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table and header disagree"
    assert lib.nc_abi_version() == 1


def test_no_cpu_fallback(lib):
    """Without a GPU every compute path must fail loudly instead of computing on the host."""
    from neuroclear_b200 import _lib, networks
    from neuroclear_b200.unet_engine import UnetDeconvEngine
    with pytest.raises(_lib.NeuroclearError):
        UnetDeconvEngine("cpu")
    import io
    from contextlib import redirect_stdout
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [], dimension=3)
    with torch.no_grad(), pytest.raises(_lib.NeuroclearError):
        net(torch.zeros(1, 1, 8, 8, 8))


def test_define_g_state_dict_is_the_references():
    import io
    from contextlib import redirect_stdout
    from neuroclear_b200 import networks
    from oracle import unet
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [], dimension=3)
    sd = net.state_dict()
    assert list(sd.keys()) == list(unet.STATE_DICT_SHAPES.keys())
    assert {k: tuple(v.shape) for k, v in sd.items()} == unet.STATE_DICT_SHAPES
    assert sum(p.numel() for p in net.parameters()) == 7_077_251
    assert all(float(v.abs().sum()) == 0.0 for k, v in sd.items() if k.endswith("bias"))   # init_weights zeroes biases
    net.load_state_dict(unet.random_state_dict(0))                                          # reference checkpoints load
    with pytest.raises(NotImplementedError):
        networks.define_G(1, 1, 64, "resnet_9blocks", "instance")


@pytest.mark.parametrize("size,roi,ov", [((128, 128, 128), 120, 15), ((900, 900, 900), 120, 15),
                                         ((1024, 2048, 2048), 120, 15), ((31, 40, 27), 12, 3), ((1, 1, 1), 8, 2),
                                         ((105, 210, 119), 120, 15)])
def test_c_geometry_matches_oracle(lib, size, roi, ov):
    from neuroclear_b200.dicing import dice_geometry
    g, o = dice_geometry(size, roi, ov, 1), ogeo.dice_geometry(size, roi, ov, 1)
    assert (g.padded, g.steps, g.n_cubes) == (o.padded, o.steps, o.n_cubes)
    assert all(p > n for p, n in zip(g.padded, g.size))          # pad is always >= 1 (util/util.py:207-209)


def test_c_geometry_rejects_bad_arguments(lib):
    from neuroclear_b200 import _lib
    from neuroclear_b200.dicing import dice_geometry
    with pytest.raises(_lib.NeuroclearError):
        dice_geometry((10, 10, 10), 8, 8, 1)
    with pytest.raises(_lib.NeuroclearError):
        dice_geometry((10, 0, 10), 8, 2, 1)


@pytest.mark.parametrize("n", [2, 7, 1000, 39 * 48 * 39, 884_736_000])
def test_percentile_rank_arithmetic_matches_numpy(n):
    from neuroclear_b200.dicing import percentile_ranks
    ranks, t_lo, t_hi = percentile_ranks(n, (0.25, 99.75))
    if n <= 100_000:
        rng = np.random.default_rng(n)
        a = rng.random(n, dtype=np.float32)
        s = np.sort(a)

        def lerp(lo, hi, t):   # numpy _lerp with float32 order statistics and float64 gamma
            diff = np.float32(s[hi] - s[lo])
            r = np.float64(s[lo]) + np.float64(diff) * t
            return np.float64(s[hi]) - np.float64(diff) * (1 - t) if t >= 0.5 else r
        mine = (lerp(ranks[0], ranks[1], t_lo), lerp(ranks[2], ranks[3], t_hi))
        ref = np.percentile(a, (0.25, 99.75))
        assert ref.dtype == np.float64 and mine[0] == ref[0] and mine[1] == ref[1]
    assert 0 <= ranks[0] <= ranks[1] < n and ranks[2] <= ranks[3] < n


def test_sharding_plans_cover_everything():
    from neuroclear_b200 import sharding
    for size, world in [((900, 900, 900), 8), ((128, 128, 128), 2), ((31, 40, 27), 3), ((1024, 2048, 2048), 8)]:
        roi, ov, bc = (12, 3, 2) if size[0] < 100 else (120, 15, 10)
        g = ogeo.dice_geometry(size, roi, ov, bc)
        cubes = sharding.balanced_ranges(g.n_cubes, world)
        slabs = sharding.balanced_ranges(g.padded[0], world)
        assert cubes[0][0] == 0 and cubes[-1][1] == g.n_cubes and all(a[1] == b[0] for a, b in zip(cubes, cubes[1:]))
        assert max(c1 - c0 for c0, c1 in cubes) - min(c1 - c0 for c0, c1 in cubes) <= 1
        plan = sharding.plan_pieces(g, cubes, slabs)
        # every plane of every cube lands on exactly one slab owner
        planes = np.zeros((g.n_cubes, roi), dtype=np.int32)
        for src in range(world):
            for dst in range(world):
                for pc in plan[src][dst]:
                    assert cubes[src][0] <= pc.cube < cubes[src][1]
                    z0, _ = sharding.cube_z_extent(g, pc.cube)
                    assert slabs[dst][0] <= z0 + pc.p0 and z0 + pc.p1 <= slabs[dst][1]
                    planes[pc.cube, pc.p0:pc.p1] += 1
        assert (planes == 1).all()
        # input plane ranges contain everything reflect indexing can touch
        for c0, c1 in cubes:
            lo, hi = sharding.input_plane_range(g, c0, c1)
            for cube in (c0, c1 - 1):
                oz = g.origin(cube)[0]
                j = np.arange(oz - bc, oz - bc + g.edge)
                j = np.where(j < 0, -j, j)
                j = np.where(j >= g.padded[0], 2 * (g.padded[0] - 1) - j, j)
                j = j[j < size[0]]
                assert j.size == 0 or (lo <= j.min() and j.max() < hi)


def test_deep_linear_tail_fold_is_exact():
    """deeplinear_engine.fold_tail: the k3 layer followed by the three 1x1 layers == ONE 64->1 k3 stencil, borders
    included (only pointwise ops follow the k3 layer), checked against the layered evaluation on the CPU."""
    import torch.nn.functional as F
    from neuroclear_b200.deeplinear_engine import FLOP_PER_VOXEL, KEYS, SHAPES, fold_tail
    from oracle import deeplinear
    assert SHAPES == deeplinear.STATE_DICT_SHAPES and tuple(KEYS) == tuple(deeplinear.STATE_DICT_SHAPES)
    assert FLOP_PER_VOXEL * 108 ** 3 == pytest.approx(1630.4e9, rel=1e-3)
    sd = deeplinear.random_state_dict(seed=1)
    g = torch.Generator().manual_seed(2)
    h2 = torch.randn((2, 64, 5, 6, 7), generator=g)
    layered = h2
    for k in KEYS[2:]:
        layered = F.conv3d(layered, sd[k], padding=1 if sd[k].shape[-1] == 3 else 0)
    K = fold_tail(*[sd[k] for k in KEYS[2:]])
    assert tuple(K.shape) == (64, 27)
    folded = F.conv3d(h2, K.reshape(1, 64, 3, 3, 3), padding=1)
    assert float((folded - layered).abs().max()) <= 1e-5 * float(layered.abs().max())


def test_generator_modules_keep_the_reference_state_dicts():
    import io
    from contextlib import redirect_stdout
    from neuroclear_b200 import networks
    from oracle import deeplinear
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "deep_linear_gen", "instance", False, "kaiming", 0.02, [], dimension=3)
    sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == deeplinear.STATE_DICT_SHAPES
    assert sum(p.numel() for p in net.parameters()) == deeplinear.N_PARAMS
    net.load_state_dict(deeplinear.random_state_dict(0))
    with pytest.raises(Exception):                      # CPU tensors: no fallback
        net(torch.zeros((1, 1, 8, 8, 8)))


def test_fused_adam_is_a_torch_optimizer_with_adam_semantics():
    """FusedAdam's host logic (param groups, state, pointer table, version bump, lr schedulers) with the kernel
    launch replaced by the same arithmetic on host memory: must track torch.optim.Adam step for step."""
    import ctypes as C
    from argparse import Namespace
    from neuroclear_b200 import networks
    from neuroclear_b200.apollo_d_path import FusedAdam

    class HostAdam(FusedAdam):
        def _launch(self, table, offset, count, group, step):
            b1, b2 = group["betas"]
            bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
            for p_, g_, m_, v_, n in table[offset:offset + count].tolist():
                arr = lambda a: np.ctypeslib.as_array((C.c_float * n).from_address(a))
                p, g, m, v = arr(p_), arr(g_), arr(m_), arr(v_)
                m[:] = b1 * m + (1 - b1) * g
                v[:] = b2 * v + (1 - b2) * g * g
                p -= (group["lr"] / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + group["eps"])

    g = torch.Generator().manual_seed(0)
    init = [torch.randn(s, generator=g) for s in [(7, 3), (5,), (2, 2, 2)]]
    mine = [torch.nn.Parameter(t.clone()) for t in init]
    ref = [torch.nn.Parameter(t.clone()) for t in init]
    o_mine = HostAdam([{"params": mine[:2]}, {"params": mine[2:], "lr": 3e-3}], lr=1e-2, betas=(0.1, 0.999))
    o_ref = torch.optim.Adam([{"params": ref[:2]}, {"params": ref[2:], "lr": 3e-3}], lr=1e-2, betas=(0.1, 0.999))
    opt = Namespace(lr_policy="linear", epoch_count=1, n_epochs=1, n_epochs_decay=3)
    s_mine, s_ref = networks.get_scheduler(o_mine, opt), networks.get_scheduler(o_ref, opt)
    assert o_mine.params == mine
    for it in range(4):
        for a, b in zip(mine, ref):
            grad = torch.randn(a.shape, generator=g)
            a.grad, b.grad = grad.clone(), grad.clone()
        if it == 2:
            mine[1].grad = ref[1].grad = None                      # parameters without gradient are skipped
        v0 = mine[0]._version
        o_mine.step()
        o_ref.step()
        assert mine[0]._version > v0                                # weight caches / autograd see the update
        s_mine.step()
        s_ref.step()
        assert [gr["lr"] for gr in o_mine.param_groups] == [gr["lr"] for gr in o_ref.param_groups]
        for a, b in zip(mine, ref):
            assert float((a - b).abs().max()) <= 1e-6
    o_mine.zero_grad()
    assert all(p.grad is None for p in mine)
    for policy, kw in (("constant", {}), ("step", {"lr_decay_iters": 2}), ("cosine", {"n_epochs": 5})):
        networks.get_scheduler(o_mine, Namespace(lr_policy=policy, **kw)).step()
    with pytest.raises(NotImplementedError):
        networks.get_scheduler(o_mine, Namespace(lr_policy="nope"))


def test_save_and_load_networks_round_trip(tmp_path):
    """'<epoch>_net_<name>.pth' files as BaseModel.save_networks writes them (base_model.py:146-201)"""
    import io
    from contextlib import redirect_stdout
    from neuroclear_b200 import networks
    from neuroclear_b200.apollo_model import load_networks, save_networks
    from oracle import deeplinear
    with redirect_stdout(io.StringIO()):
        a = networks.define_G(1, 1, 64, "deep_linear_gen", "instance", False, "kaiming", 0.02, [], dimension=3)
        b = networks.define_G(1, 1, 64, "deep_linear_gen", "instance", False, "kaiming", 0.02, [], dimension=3)
    a.load_state_dict(deeplinear.random_state_dict(5))
    save_networks({"G_B": torch.nn.DataParallel(a) if False else a}, str(tmp_path), "latest")
    assert (tmp_path / "latest_net_G_B.pth").exists()
    with redirect_stdout(io.StringIO()):
        load_networks({"G_B": b}, str(tmp_path), "latest")
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the package (nor __graft_entry__.build) may import it, and in
    bench.py only the cpu_baseline / --impl reference legs do."""
    import ast
    pkg = os.path.join(ROOT, "neuroclear_b200")
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            mods = [a.name for a in node.names] if isinstance(node, ast.Import) else \
                [node.module or ""] if isinstance(node, ast.ImportFrom) else []
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), fn
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"CpuSample", "train_step_cpu_baseline"}     # the cpu_baseline / --impl reference legs
    for top in tree.body:
        if isinstance(top, (ast.FunctionDef, ast.ClassDef)):
            for node in ast.walk(top):
                if isinstance(node, (ast.ImportFrom, ast.Import)):
                    mods = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""]
                    if any(m.split(".")[0] == "oracle" for m in mods):
                        assert top.name in allowed, top.name
    for node in tree.body:                                              # nothing at module level
        assert not (isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle")


def test_flip_tta_helpers_match_the_reference_semantics():
    """Assemble_Dice.varycubeinput / combinecube (assemble_dice.py:79-128): for a flip-equivariant 'network' the
    combined result equals the plain one; for a general one it is the mean over the four un-flipped outputs."""
    from collections import OrderedDict
    from neuroclear_b200.dicing import Assemble_Dice
    helper = Assemble_Dice.__new__(Assemble_Dice)           # the helpers use no instance state
    g = torch.Generator().manual_seed(0)
    x = torch.rand((1, 1, 4, 5, 6), generator=g)
    copies = helper.varycubeinput(OrderedDict([("A", x), ("A_paths", "7")]))
    assert len(copies) == 4 and all(c["A_paths"] == "7" for c in copies)
    assert torch.equal(copies[0]["A"], x) and all(torch.equal(copies[1 + i]["A"], x.flip(2 + i)) for i in range(3))
    w = torch.rand((1, 1, 4, 5, 6), generator=g)
    net = lambda t: t * w                                    # not flip-equivariant
    outs = [OrderedDict([("real", c["A"]), ("fake", net(c["A"]))]) for c in copies]
    comb = helper.combinecube(outs)
    want = (x * w + sum((x.flip(2 + i) * w).flip(2 + i) for i in range(3))) / 4
    assert torch.allclose(comb["fake"], want) and torch.allclose(comb["real"], x)


def test_host_side_sizing_functions_of_the_training_abi(lib):
    """pure host functions of the library (no launch): workspace / scratch sizing used by the training path"""
    # PatchGAN workspace: conv outputs + IN outputs + statistics of the five layers + two gradient ping-pong buffers
    n = lib.nc_patchgan_ws_floats(1, 108, 108, 64, 3)
    sizes = [(64, 54), (128, 27), (256, 13), (512, 12), (1, 11)]
    acts = [c * s * s for c, s in sizes]
    want = sum(acts) + sum(acts[1:4]) + 2 * (128 + 256 + 512) + 2 * max(max(acts), 108 * 108)
    assert n == want
    assert lib.nc_patchgan_ws_floats(1, 16, 16, 64, 3) == -1 and b"too small" in lib.nc_last_error()
    # weight-gradient split-K scratch: splits x taps x Cin x Cout floats, splits = #SMs / work items (148 without a GPU)
    for ks, taps, groups in ((3, 27, 1), (5, 125, 2), (1, 1, 1), (71, 7, 1)):
        items = 1 * 1 * (7 if ks == 71 else ks) * groups
        b = lib.nc_conv3d_wgrad_scratch_bytes(ks, 1, 108, 108, 108, 64, 64)
        assert b == (148 // items) * taps * 64 * 64 * 4, ks
    assert lib.nc_conv3d_wgrad_scratch_bytes(3, 1, 2, 8, 8, 256, 256) == 2 * 27 * 256 * 256 * 4     # 2 plane tiles only
    assert lib.nc_bwd_scratch_bytes(2) == 148 * 8 * 2 * 512 * 4


def test_input_row_end_covers_every_row_a_cube_range_reads(lib):
    """sharding.input_row_end (the bound the y-chunked upload waits for) against a brute-force enumeration of the
    reflected / zero-padded rows every cube of the range reads."""
    from neuroclear_b200 import sharding
    from neuroclear_b200.dicing import dice_geometry
    for size, roi, ov, bc in [((900, 900, 900), 120, 15, 10), ((40, 41, 58), 24, 6, 4), ((31, 40, 27), 12, 3, 2)]:
        geo = dice_geometry(size, roi, ov, bc)
        nz, ny, nx = geo.steps
        P = geo.padded[1]
        rng = np.random.default_rng(0)
        spans = [(0, geo.n_cubes), (0, 1), (geo.n_cubes - 1, geo.n_cubes)]
        spans += [tuple(sorted(rng.integers(0, geo.n_cubes + 1, 2))) for _ in range(40)]
        for c0, c1 in spans:
            if c1 <= c0:
                continue
            need = 0
            for c in range(c0, c1):
                cy = (c % (nx * ny)) // nx
                js = np.arange(cy * geo.step - bc, cy * geo.step + roi + bc)
                ks = np.where(js < 0, -js, np.where(js >= P, 2 * (P - 1) - js, js))
                ks = ks[ks < size[1]]                       # rows beyond the data are zero padding, never read
                need = max(need, int(ks.max()) + 1 if ks.size else 0)
            got = sharding.input_row_end(geo, int(c0), int(c1))
            assert need <= got <= size[1], (size, c0, c1, need, got)


def test_parity_class_data_gradient_index_math():
    """The stride-2 data gradient of the PatchGAN convolutions is computed per parity class of the input pixel
    (csrc/disc2d.cu conv2d_k4_dgrad_s2_kernel): class (h % 2, w % 2) is reached only through the taps kh = (h + 1) % 2
    (+ 2), kw = (w + 1) % 2 (+ 2), from the output pixel ((h + 1 - kh) / 2, (w + 1 - kw) / 2).  The same index math in
    numpy against torch's autograd, odd and even image sizes (classes of different size)."""
    import numpy as np
    import torch
    rng = np.random.default_rng(0)
    for (cin, cout, H, W) in [(3, 5, 9, 12), (2, 4, 8, 7), (1, 2, 5, 5)]:
        x = torch.tensor(rng.standard_normal((1, cin, H, W)), dtype=torch.float64, requires_grad=True)
        w = torch.tensor(rng.standard_normal((cout, cin, 4, 4)), dtype=torch.float64)
        y = torch.nn.functional.conv2d(x, w, stride=2, padding=1)
        dy = torch.tensor(rng.standard_normal(tuple(y.shape)), dtype=torch.float64)
        y.backward(dy)
        Ho, Wo = y.shape[2:]
        dx = np.zeros((cin, H, W))
        dyn, wn = dy.numpy()[0], w.numpy()
        for cls in range(4):
            ph, pw = cls >> 1, cls & 1
            kh0, kw0 = (ph + 1) & 1, (pw + 1) & 1
            Hq, Wq = (H - ph + 1) // 2, (W - pw + 1) // 2
            for p in range(Hq * Wq):
                i, j = divmod(p, Wq)
                h, ww = 2 * i + ph, 2 * j + pw
                for tap in range(4):
                    kh, kw = kh0 + 2 * (tap >> 1), kw0 + 2 * (tap & 1)
                    th, tw = h + 1 - kh, ww + 1 - kw
                    assert th % 2 == 0 and tw % 2 == 0
                    if th < 0 or th // 2 >= Ho or tw < 0 or tw // 2 >= Wo:
                        continue
                    dx[:, h, ww] += wn[:, :, kh, kw].T @ dyn[:, th // 2, tw // 2]
        assert np.abs(dx - x.grad.numpy()[0]).max() <= 1e-12 * max(1.0, np.abs(dx).max())


def test_three_tf32_products_keep_fp32_accuracy():
    """The PatchGAN convolutions evaluate an fp32 product as lo*hi + hi*lo + hi*hi with hi = tf32(x), lo = tf32(x - hi)
    (cvt.rna.tf32: 10 mantissa bits, round half away; csrc/disc2d.cu).  Emulated in numpy: the dropped lo*lo term and
    the rounding of lo leave ~2^-21 per product, i.e. fp32-level results, where a single TF32 product is 2^-11."""
    import numpy as np

    def tf32(v):
        b = v.astype(np.float32).view(np.uint32)
        return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)

    rng = np.random.default_rng(1)
    a = rng.standard_normal((64, 4096)).astype(np.float32)
    b = rng.standard_normal((4096,)).astype(np.float32)
    ah, bh = tf32(a), tf32(b)
    al, bl = tf32(a - ah), tf32(b - bh)
    f64 = lambda v: v.astype(np.float64)
    exact = f64(a) @ f64(b)
    three = f64(al) @ f64(bh) + f64(ah) @ f64(bl) + f64(ah) @ f64(bh)
    one = f64(ah) @ f64(bh)
    scale = np.abs(f64(a)) @ np.abs(f64(b))
    assert (np.abs(three - exact) / scale).max() <= 2.0 ** -20
    assert (np.abs(one - exact) / scale).max() >= 2.0 ** -16      # what a plain TF32 GEMM would give
