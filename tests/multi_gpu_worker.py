"""Run under torchrun (one rank per GPU, NCCL): the sharded pipeline must reproduce the single-GPU result BIT FOR BIT
on every rank's output slab (cube-range sharding + piece exchange + histogram all-reduce change no value)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def data_parallel_discriminator_step(rank, world, dev):
    """an N-GPU discriminator step == a single-process step on the same N crops with the
    gradients averaged.  Every rank also replays all crops locally and compares the updated parameters."""
    import io
    from argparse import Namespace
    from contextlib import redirect_stdout
    from neuroclear_b200.apollo_d_path import ApolloDiscriminatorPath
    from oracle import discriminator as odisc
    opt = Namespace(gan_mode="lsgan", randomize_projection_depth=False, projection_depth=6, min_projection_depth=2,
                    lambda_plane=[1, 1, 1], input_nc=1, output_nc=1, ndf=64, netD="basic", n_layers_D=3,
                    norm="instance", init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1, lambda_A=5.0)
    names = ["D_A_axial", "D_A_lateral", "D_B_axial", "D_B_lateral"]

    def make(distributed):
        with redirect_stdout(io.StringIO()):
            p = ApolloDiscriminatorPath(opt, dev, distributed=distributed)
        for i, n in enumerate(names):
            getattr(p, "net" + n).module.load_state_dict(odisc.random_state_dict(seed=20 + i))
        return p

    def crops(r):
        g = torch.Generator().manual_seed(100 + r)
        return [torch.rand((1, 1, 32, 32, 32), generator=g).to(dev) for _ in range(3)]

    dp = make(True)
    np.random.seed(50 + rank)
    dp.optimize_D(*crops(rank))
    # single-process replay: gradients of every crop (same per-rank RNG streams), averaged, one Adam step
    sp = make(False)
    params = sp.optimizer_D.params
    acc = [torch.zeros_like(p) for p in params]
    for r in range(world):
        sp.optimizer_D.zero_grad()
        np.random.seed(50 + r)
        real, fake, rec = crops(r)
        sp.backward_D_all(real, fake, rec)
        for a, p in zip(acc, params):
            a += p.grad
    for a, p in zip(acc, params):
        p.grad = a / world
    sp.optimizer_D.step()
    worst = max((a.detach() - b.detach()).abs().max().item() for a, b in zip(dp.optimizer_D.params, params))
    print("rank %d/%d data-parallel D step vs single-process replay: max |dparam| %.3g" % (rank, world, worst), flush=True)
    return worst <= 2e-6


def data_parallel_apollo_step(rank, world, dev):
    """Full apollo iteration, data parallel (BASELINE.json configs[3]): every rank trains on its own crop, the
    gradients of both optimisers are averaged with one all-reduce each, so all ranks must hold identical weights
    afterwards and those weights must equal a single-process replay with averaged gradients."""
    import io
    from argparse import Namespace
    from contextlib import redirect_stdout
    from neuroclear_b200.apollo_model import AxialToLateralGANApolloModel
    from oracle import deeplinear, discriminator as odisc, unet
    opt = Namespace(isTrain=True, gpu_ids=[dev.index], gan_mode="lsgan", randomize_projection_depth=False,
                    projection_depth=6, min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1,
                    ngf=64, ndf=64, netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3,
                    norm="instance", no_dropout=True, init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1,
                    direction="AtoB", lambda_A=5.0)
    names = ["D_A_axial", "D_A_lateral", "D_B_axial", "D_B_lateral"]

    def make(distributed):
        with redirect_stdout(io.StringIO()):
            m = AxialToLateralGANApolloModel(opt, dev, distributed=distributed)
        m.netG_A.module.load_state_dict(unet.random_state_dict(seed=41, bias_std=0.05))
        m.netG_B.module.load_state_dict(deeplinear.random_state_dict(seed=42))
        for i, n in enumerate(names):
            getattr(m, "net" + n).module.load_state_dict(odisc.random_state_dict(seed=50 + i))
        return m

    crop = lambda r: torch.rand((1, 1, 24, 24, 24), generator=torch.Generator().manual_seed(200 + r))
    dp = make(True)
    np.random.seed(70 + rank)
    dp.set_input({"A": crop(rank), "A_paths": "x"})
    dp.optimize_parameters()
    # single-process replay of the generator update: gradients of every rank's crop, averaged, one Adam step
    sp = make(False)
    params = sp.optimizer_G.params
    acc = [torch.zeros_like(p) for p in params]
    for r in range(world):
        sp.optimizer_G.zero_grad()
        np.random.seed(70 + r)
        sp.set_input({"A": crop(r), "A_paths": "x"})
        sp.forward()
        sp.backward_G()
        for a, p in zip(acc, params):
            a += p.grad
    for a, p in zip(acc, params):
        p.grad = a / world
    sp.optimizer_G.step()
    worst = max((a.detach() - b.detach()).abs().max().item() for a, b in zip(dp.optimizer_G.params, params))
    # all ranks hold the same generator weights
    flat = torch.cat([p.detach().reshape(-1) for p in dp.optimizer_G.params])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    print("rank %d/%d data-parallel apollo step: generators vs single-process replay max |dparam| %.3g, "
          "identical across ranks %s" % (rank, world, worst, same), flush=True)
    return worst <= 2e-6 and same


def main():
    from neuroclear_b200.pipeline import DicedInference
    from oracle import unet as ounet
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = ounet.random_state_dict(seed=0, bias_std=0.1)
    ok = True
    for shape, (roi, ov, bc), normalize in [((40, 58, 38), (24, 6, 4), True), ((70, 41, 58), (24, 6, 4), False),
                                            ((128, 128, 128), (120, 15, 10), True)]:
        vol = (np.random.default_rng(1).random(shape) ** 3 * 65535).astype(np.uint16)
        sharded = DicedInference(sd, dev, roi, ov, bc, normalize_intensity=normalize, batch=3)
        got, (z0, z1) = sharded.run(vol)
        single = DicedInference(sd, dev, roi, ov, bc, normalize_intensity=normalize, batch=3, distributed=False)
        ref, _ = single.run(vol)
        same = np.array_equal(got, ref[z0:z1])
        pc = (tuple(sharded.last["percentiles"].cpu().tolist()) == tuple(single.last["percentiles"].cpu().tolist())
              if normalize else True)
        print("rank %d/%d shape %s slab [%d,%d): identical=%s percentiles_identical=%s" %
              (rank, world, shape, z0, z1, same, pc), flush=True)
        ok = ok and same and pc
    ok = data_parallel_discriminator_step(rank, world, dev) and ok
    ok = data_parallel_apollo_step(rank, world, dev) and ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
