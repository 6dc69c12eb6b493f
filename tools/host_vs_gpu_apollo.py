"""Is the apollo iteration host-bound or GPU-bound?  Enqueues K iterations back to back and reports the host time to
enqueue them (no synchronisation inside) next to the GPU time between two events around the same region.
Usage: python tools/host_vs_gpu_apollo.py [crop=108] [iters=20]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

crop = int(sys.argv[1]) if len(sys.argv) > 1 else 108
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
ms, launches, _, model = bench.time_apollo_iterations(dev, crop, 10, 6, False)
g = torch.Generator().manual_seed(7)
crops = [torch.rand((1, 1, crop, crop, crop), generator=g).pin_memory() for _ in range(2)]
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(iters):
        model.set_input({"A": crops[i % 2], "A_paths": "synthetic"})
        model.optimize_parameters()
    e1.record()
    t_host = (time.perf_counter() - t0) * 1e3 / iters
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) * 1e3 / iters
    print("crop %d: host enqueue %.2f ms/iter, GPU %.2f ms/iter, wall incl. drain %.2f ms/iter (bench: %.2f)" %
          (crop, t_host, e0.elapsed_time(e1) / iters, t_all, ms))
