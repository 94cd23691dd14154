"""GPU box: milliseconds the raster kernel of the library in DYNENV_B200_LIB takes for the 100 camera images of 1024 RoboCup
IMAGE environments (5 ticks x 10 robots x 8 bit-planes each) a few steps into the episode; prints the plane bandwidth"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dynenv_b200

n, P = 1024, 10
sim = dynenv_b200.BatchedRoboCup(n, 5, dynenv_b200.ObservationType.IMAGE, 1, 2.0, allowHeadTurn=True, canFall=True, auto_reset=True, seed=42)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1234)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 25):
    a = torch.rand((n, P, 4), generator=g, device="cuda")  # the bench's image workload: head turn uniform in [-3, 3)
    a[..., :3] = (a[..., :3] * torch.tensor([5, 3, 3], device="cuda")).floor()
    a[..., 3] = a[..., 3] * 6.0 - 3.0
    sim.step(a.contiguous())
out = torch.empty((256, 5, P, 8, 480, 20), dtype=torch.int32, device="cuda")
sim.render_images(0, 256, out)
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c in range(4):
        sim.render_images(256 * c, 256 * (c + 1), out)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("%-28s %.2f ms per 1024 envs = %.0f GB/s of bit-planes, checksum %d" % (
    os.path.basename(os.environ.get("DYNENV_B200_LIB", "default")), best, n * 5 * P * 8 * 38400 / best / 1e6, int(out.sum().item())))
