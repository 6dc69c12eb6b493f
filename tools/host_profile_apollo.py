"""Host-side cost of one apollo iteration: cProfile over K iterations at a SMALL crop (GPU work negligible, the loop
runs at the host's pace).  Usage: python tools/host_profile_apollo.py [crop=40] [iters=30]"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

crop = int(sys.argv[1]) if len(sys.argv) > 1 else 40
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)
ms, launches, _, model = bench.time_apollo_iterations(dev, crop, 10, 6, False)
crops = [torch.rand((1, 1, crop, crop, crop)).pin_memory() for _ in range(2)]


def loop():
    for i in range(iters):
        model.set_input({"A": crops[i % 2], "A_paths": "synthetic"})
        model.optimize_parameters()


torch.cuda.synchronize()
t0 = time.perf_counter()
loop()
t_host = (time.perf_counter() - t0) * 1e3 / iters
torch.cuda.synchronize()
print("crop %d: host %.2f ms/iter (bench back-to-back %.2f ms/iter, %d library calls)" % (crop, t_host, ms, launches))
pr = cProfile.Profile()
pr.enable()
loop()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:48]))
