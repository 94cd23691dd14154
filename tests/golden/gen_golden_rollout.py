"""Generates tests/golden/rollout.npz with the REFERENCE's own RolloutStorage._discount_rewards
(DynEnv/models/storage.py:170-205), executed from the reference's source text (the package import pulls matplotlib / h5py /
gym, which this image lacks; the one relative import of storage.py is a loss dataclass the method does not touch).
Run here only: python tests/golden/gen_golden_rollout.py"""
import os

import numpy as np
import torch

SRC = "/root/reference/DynEnv/models/storage.py"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_storage():
    txt = open(SRC).read().replace("from ..models.loss_descriptors import A2CLosses", "A2CLosses = None")
    ns = {}
    exec(compile(txt, SRC, "exec"), ns)
    return ns["RolloutStorage"]


if __name__ == "__main__":
    RolloutStorage = load_storage()
    rec = {}
    for k, (T, E, A, disc, seed) in enumerate([(5, 4, 10, 0.99, 1), (20, 3, 2, 0.9, 2), (1, 7, 10, 0.99, 3), (64, 2, 10, 0.999, 4)]):
        st = RolloutStorage(T, E, A, 4, 1, feature_size=8, is_cuda=False)
        g = torch.Generator().manual_seed(seed)
        st.rewards.copy_(torch.randn((T, E * A), generator=g) * 0.3)
        st.dones.copy_((torch.rand((T, E), generator=g) < 0.2).float())  # the reference ignores them (:194)
        fv = torch.randn((E * A,), generator=g)
        out = st._discount_rewards(fv, disc)
        rec.update({"rewards%d" % k: st.rewards.numpy().copy(), "dones%d" % k: st.dones.numpy().astype(np.uint8), "final%d" % k: fv.numpy(),
                    "out%d" % k: out.numpy(), "shape%d" % k: np.array([T, E, A]), "discount%d" % k: np.array(disc)})
    np.savez_compressed(os.path.join(OUT, "rollout.npz"), **rec)
    print("cases", k + 1)
