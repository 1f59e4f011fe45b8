"""Bit-level regression aid: 40 TF32 steps (4 members, K = 8 per call) -> losses + arena checksums, saved to the given .npz."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jsrl_corl_b200 import IQLEnsemble, ReplayBuffer
from jsrl_corl_b200.synthetic import synthetic_dataset

out = sys.argv[1]
res = {}
for name, S, A, L, det in (("hc", 17, 6, 2, False), ("ant", 29, 8, 3, False)):
    ens = IQLEnsemble(4, S, A, 256, L, 256, deterministic=det, math_mode="tf32", seeds=[1, 2, 3, 4], max_steps_per_call=8)
    rb = ReplayBuffer(S, A, 20000, "cuda"); rb.load_d4rl_dataset(synthetic_dataset(20000, S, A, 0)); ens.bind_replay(rb)
    losses = np.concatenate([ens.train_steps(8).cpu().numpy() for _ in range(5)], axis=1)
    res[name + "_losses"] = losses
    res[name + "_params"] = ens.engine.params.cpu().numpy()
    res[name + "_target"] = ens.engine.target.cpu().numpy()
    res[name + "_m"] = ens.engine.exp_avg.cpu().numpy()
np.savez_compressed(out, **res)
print("saved", out, {k: float(np.abs(v).sum()) for k, v in res.items() if "losses" in k})
