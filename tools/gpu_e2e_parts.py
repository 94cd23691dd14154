"""Dev experiment (GPU box): where the end-to-end step of bench.py spends its time (pinned H2D of the actions, the step,
D2H of rewards + done, the per-step synchronisation)."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dynenv_b200

kind = sys.argv[1] if len(sys.argv) > 1 else "driving"
E = 65536
if kind == "driving":
    sim = dynenv_b200.BatchedDriving(E, 10, 1, 1, 2.0, continuousActions=True, auto_reset=True, seed=42)
    P, A = 10, 2
    mk = lambda: torch.rand((E, P, A)) * 6 - 3
else:
    sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=42)
    P, A = 10, 4
    mk = lambda: (torch.rand((E, P, A)) * torch.tensor([5, 3, 3, 7])).floor()
sim.reset()
h_act = [mk().pin_memory() for _ in range(8)]
d_act = torch.empty((E, P, A), device="cuda")
h_rew = torch.empty((E, P), dtype=torch.float32).pin_memory()
h_done = torch.empty((E,), dtype=torch.int32).pin_memory()
h_both = torch.empty((E, P + 1), dtype=torch.float32).pin_memory()
st = torch.cuda.current_stream()
for s in range(30):
    d_act.copy_(h_act[s % 8], non_blocking=True)
    out = sim.step(d_act)
torch.cuda.synchronize()


def timed(name, fn, n=40):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for s in range(n):
        fn(s)
    torch.cuda.synchronize()
    print("%-46s %.3f ms / step" % (name, (time.perf_counter() - t) / n * 1e3))


def full(s):
    d_act.copy_(h_act[s % 8], non_blocking=True)
    o = sim.step(d_act)
    h_rew.copy_(o.rewards, non_blocking=True)
    h_done.copy_(o.done, non_blocking=True)
    st.synchronize()


def full_one_copy(s):
    d_act.copy_(h_act[s % 8], non_blocking=True)
    o = sim.step(d_act)
    h_both.copy_(o.info[:, :P + 1], non_blocking=True)
    st.synchronize()


def full_contig(s):
    d_act.copy_(h_act[s % 8], non_blocking=True)
    o = sim.step(d_act)
    h_both.copy_(o.info[:, :P + 1].contiguous(), non_blocking=True)
    st.synchronize()


timed("step only (no sync per step)", lambda s: sim.step(d_act))
timed("step + sync per step", lambda s: (sim.step(d_act), st.synchronize()))
timed("H2D actions only", lambda s: (d_act.copy_(h_act[s % 8], non_blocking=True), st.synchronize()))
timed("D2H rewards + done only", lambda s: (h_rew.copy_(out.rewards, non_blocking=True), h_done.copy_(out.done, non_blocking=True), st.synchronize()))
timed("D2H one strided copy", lambda s: (h_both.copy_(out.info[:, :P + 1], non_blocking=True), st.synchronize()))
timed("D2H gather kernel + contiguous copy", lambda s: (h_both.copy_(out.info[:, :P + 1].contiguous(), non_blocking=True), st.synchronize()))
timed("end to end as bench.py", full)
timed("end to end, one strided copy", full_one_copy)
timed("end to end, gather + contiguous copy", full_contig)
