"""GPU box: one step of golden Driving scenarios on the GPU and in the oracle; dumps both for inspection."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import drv_golden_cfg, drv_oracle_step_batch, load_golden
from oracle.oracle import drv_f32_record_to_words, drv_words_to_f32_record
import dynenv_b200

for name in sys.argv[1:]:
    g = load_golden(name, "driving_")
    cfg = drv_golden_cfg(g)
    words = np.concatenate([g["state0"][None], g["states"][:-1]], 0)
    recs = drv_words_to_f32_record(words)
    seed = 2000 + len(name)
    sim = dynenv_b200.BatchedDriving(len(recs), cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"],
                                     continuousActions=cfg["continuous"], seed=seed)
    sim.set_state(torch.from_numpy(recs.view(np.int32)))
    out = sim.step(torch.from_numpy(np.ascontiguousarray(g["actions"], dtype=np.float32)).cuda())
    torch.cuda.synchronize()
    gw = drv_f32_record_to_words(sim.get_state().cpu().numpy().view(np.uint32))
    outs, ow = drv_oracle_step_batch(cfg, recs, g["actions"], seed)
    d = {"gw": gw, "ow": ow, "rec": out.record.cpu().numpy(), "info": out.info.cpu().numpy()}
    for k in ("self", "car", "obs", "ped", "lane", "counts", "rewards"):
        d["o_" + k] = np.stack([o[k] for o in outs])
    np.savez_compressed("gpurun_out/drvdump_%s.npz" % name, **d)
