#!/usr/bin/env python
"""Soak of the host-step path: N iterations of act -> add_transition -> sample -> train with periodic checkpoints, a
resume in the middle, an in-place edit of a sampled batch now and then (staged-batch path) and a second buffer; checks
that every log dict is finite, counters add up and nothing times out."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import jsrl_corl_b200 as J
from jsrl_corl_b200.synthetic import synthetic_dataset

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
S, A, B = 11, 3, 256
torch.manual_seed(0)
np.random.seed(0)
q, v, actor = J.TwinQ(S, A), J.ValueFunction(S), J.GaussianPolicy(S, A, 1.0)
tr = J.ImplicitQLearning(1.0, actor, torch.optim.Adam(actor.parameters(), lr=3e-4), q, torch.optim.Adam(q.parameters(), lr=3e-4),
                         v, torch.optim.Adam(v.parameters(), lr=3e-4), device="cuda", max_steps=N)
off = J.ReplayBuffer(S, A, 120_000, "cuda")
off.load_d4rl_dataset(synthetic_dataset(100_000, S, A, 0))
on = J.ReplayBuffer(S, A, 5_000, "cuda")
obs = np.random.RandomState(1).randn(4096, S).astype(np.float32)
actor.eval()
t0 = time.perf_counter()
bad = 0
for i in range(N):
    o = obs[i & 4095]
    a = actor.act(o, "cuda")
    on.add_transition(o, a, 1.0, obs[(i + 1) & 4095], (i % 1000) == 999)
    off.add_transition(o, a, 1.0, obs[(i + 1) & 4095], False)
    rb = on if (i % 3 == 2 and on._size >= 1) else off
    batch = rb.sample(B)
    if i % 97 == 0:
        batch[2].mul_(1.0)  # an in-place edit: the staged-batch path
    log = tr.train(batch)
    if not all(np.isfinite(x) for x in log.values()):
        bad += 1
    if i % 20_000 == 19_999:
        sd = tr.state_dict()
        assert sd["total_it"] == i + 1 and float(sd["q_optimizer"]["state"][0]["step"]) == i + 1
        if i == 39_999:
            tr.load_state_dict({k: (v if not isinstance(v, dict) else v) for k, v in sd.items()})
        print(f"{i + 1} iterations, {(time.perf_counter() - t0) / (i + 1) * 1e6:.1f} us each, last {log}", flush=True)
torch.cuda.synchronize()
assert bad == 0 and tr.total_it == N and tr._path_counts[0] + tr._path_counts[1] == N
print(f"soak ok: {N} iterations, {tr._path_counts} (engine-gather, staged) steps, {(time.perf_counter() - t0):.1f} s")
