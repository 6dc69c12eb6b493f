"""Run under torchrun (one rank per GPU, NCCL): the sharded pipeline must reproduce the single-GPU result BIT FOR BIT
on every rank's output slab (cube-range sharding + piece exchange + histogram all-reduce change no value)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


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
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
