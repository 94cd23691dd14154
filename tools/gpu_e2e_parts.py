"""Dev experiment (GPU box): where the end-to-end step of bench.py spends its time.  Every variant runs one whole RoboCup
episode (240 steps) from a fresh reset with the same pinned action pool."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dynenv_b200
from dynenv_b200 import layout as L

E, P, A, STEPS = 65536, 10, 4, 240
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42)
g = torch.Generator().manual_seed(1)
h_act = [((torch.rand((E, P, A), generator=g) * torch.tensor([5, 3, 3, 7])).floor()).pin_memory() for _ in range(32)]
d_pool = [h.cuda() for h in h_act]
d_act = torch.empty((E, P, A), device="cuda")
n_res = L.RC_I_DONE + 1
h_res = torch.empty((E, n_res), dtype=torch.float32).pin_memory()
st = torch.cuda.current_stream()


def run(name, fn):
    sim.reset()
    for s in range(3):
        sim.step(d_pool[s])
    torch.cuda.synchronize()
    t = time.perf_counter()
    for s in range(STEPS):
        fn(s)
    torch.cuda.synchronize()
    print("%-52s %.3f ms / step" % (name, (time.perf_counter() - t) / STEPS * 1e3), flush=True)


def full(s):
    d_act.copy_(h_act[s % 32], non_blocking=True)
    o = sim.step(d_act)
    h_res.copy_(o.info[:, :n_res], non_blocking=True)
    st.synchronize()


def h2d_step_sync(s):
    d_act.copy_(h_act[s % 32], non_blocking=True)
    sim.step(d_act)
    st.synchronize()


def h2d_step(s):
    d_act.copy_(h_act[s % 32], non_blocking=True)
    sim.step(d_act)


run("step, device-resident actions, no sync", lambda s: sim.step(d_pool[s % 32]))
run("step + sync per step", lambda s: (sim.step(d_pool[s % 32]), st.synchronize()))
run("H2D + step, no sync", h2d_step)
run("H2D + step + sync per step", h2d_step_sync)
run("H2D + step + D2H + sync (bench.py e2e)", full)
run("step, device-resident actions, no sync (again)", lambda s: sim.step(d_pool[s % 32]))
