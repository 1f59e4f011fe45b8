#!/usr/bin/env python
"""us per iteration of the reference's ONLINE loop body as the host sees it (jsrl_w_iql.py:445-548 without the env):
`action = actor.act(state)` -> `replay_buffer.add_transition(...)` -> `batch = replay_buffer.sample(B)` ->
`log_dict = trainer.train(batch)`, through the drop-in classes; observations come from a precomputed array
(a real env.step sits between act and add_transition and is not part of this library)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import jsrl_corl_b200 as J
from jsrl_corl_b200.synthetic import synthetic_dataset


def run(S, A, L, det, n=3000, warm=300, B=256):
    torch.manual_seed(0)
    q, v = J.TwinQ(S, A, 256, L), J.ValueFunction(S, 256, L)
    actor = (J.DeterministicPolicy if det else J.GaussianPolicy)(S, A, 1.0, 256, L)
    tr = J.ImplicitQLearning(1.0, actor, torch.optim.Adam(actor.parameters(), lr=3e-4), q, torch.optim.Adam(q.parameters(), lr=3e-4),
                             v, torch.optim.Adam(v.parameters(), lr=3e-4), device="cuda")
    rb = J.ReplayBuffer(S, A, 200_000, "cuda")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        rb.load_d4rl_dataset(synthetic_dataset(100_000, S, A, 0))
    actor.eval()
    obs = np.random.RandomState(1).randn(n + warm + 1, S).astype(np.float32)
    np.random.seed(0)
    parts = np.zeros(4)
    t_all = 0.0
    for i in range(n + warm):
        t0 = time.perf_counter()
        a = actor.act(obs[i], "cuda")
        t1 = time.perf_counter()
        rb.add_transition(obs[i], a, 1.0, obs[i + 1], False)
        t2 = time.perf_counter()
        batch = rb.sample(B)
        t3 = time.perf_counter()
        tr.train(batch)
        t4 = time.perf_counter()
        if i >= warm:
            parts += (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
            t_all += t4 - t0
    us = parts / n * 1e6
    return {"act_us": round(us[0], 1), "add_transition_us": round(us[1], 1), "sample_us": round(us[2], 1), "train_us": round(us[3], 1),
            "iteration_us": round(t_all / n * 1e6, 1), "iterations_per_s": round(n / t_all)}


if __name__ == "__main__":
    out = {"hopper_det_2x256": run(11, 3, 2, True), "antmaze_gauss_3x256": run(29, 8, 3, False)}
    print(json.dumps(out))
