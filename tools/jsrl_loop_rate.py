#!/usr/bin/env python
"""Iterations/s of the whole offline -> online JSRL loop (`jsrl_w_iql.train_loop`, mirror of the reference's
jsrl_w_iql.py:373-604) on a trivial env of antmaze-umaze shape (obs 29, act 8, 3x256 Gaussian actor, beta 10, tau 0.9):
offline updates, then per env step guide / learner arbitration + act + env.step + add_transition + sample + train, with
periodic evaluations and the horizon curriculum.  BASELINE.json configs[1] as a host-visible loop."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from jsrl_corl_b200 import ReplayBuffer
from jsrl_corl_b200.jsrl_utils import JsrlTrainConfig
from jsrl_corl_b200.jsrl_w_iql import train_loop
from jsrl_corl_b200.synthetic import synthetic_dataset

S, A, T = 29, 8, 100


class FakeEnv:
    """gym-style 4-tuple env: random-walk observations, reward -1 until the last step of the episode."""

    def __init__(self, seed=0):
        self.rng = np.random.RandomState(seed)
        self.spec = type("Spec", (), {"id": "FakeAnt-v0"})()

    def seed(self, s):
        self.rng = np.random.RandomState(s)

    def reset(self):
        self.x = self.rng.randn(S).astype(np.float32)
        self.t = 0
        return self.x

    def step(self, action):
        self.x = (0.99 * self.x + 0.01 * self.rng.randn(S)).astype(np.float32)
        self.t += 1
        return self.x, -1.0, self.t >= T, {}


def main():
    n_off, n_on = int(sys.argv[1]) if len(sys.argv) > 1 else 3000, int(sys.argv[2]) if len(sys.argv) > 2 else 6000
    cfg = JsrlTrainConfig(device="cuda", env="FakeAnt-v0", seed=0, eval_freq=1000, n_episodes=2, offline_iterations=n_off,
                          online_iterations=n_on, batch_size=256, n_curriculum_stages=5, rolling_mean_n=1, tolerance=0.05,
                          horizon_fn="time_step", online_buffer_size=10000, checkpoints_path=None, iql_deterministic=False,
                          beta=10.0, iql_tau=0.9, normalize_reward=False)
    cfg.n_hidden = 3 if hasattr(cfg, "n_hidden") else None
    torch.manual_seed(0)
    np.random.seed(0)
    rb = ReplayBuffer(S, A, 200_000, "cuda")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        rb.load_d4rl_dataset(synthetic_dataset(100_000, S, A, 0, antmaze_rewards=True))
    stamps = {}

    def log(d, step):
        now = time.perf_counter()
        if "offline_iter" in d:
            stamps.setdefault("off0", (now, d["offline_iter"]))
            stamps["off1"] = (now, d["offline_iter"])
        elif "online_iter" in d:
            stamps.setdefault("on0", (now, d["online_iter"]))
            stamps["on1"] = (now, d["online_iter"])

    t0 = time.perf_counter()
    trainer, cfg, history = train_loop(cfg, FakeEnv(0), FakeEnv(1), rb, S, A, 1.0, max_steps=T, log=log)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    off = (stamps["off1"][1] - stamps["off0"][1]) / (stamps["off1"][0] - stamps["off0"][0])
    on = (stamps["on1"][1] - stamps["on0"][1]) / (stamps["on1"][0] - stamps["on0"][0])
    print(json.dumps({"shape": "antmaze-umaze (obs 29, act 8), Gaussian actor, batch 256", "offline_iterations_per_s": round(off),
                      "online_iterations_per_s": round(on), "online_us_per_iteration": round(1e6 / on, 1),
                      "evaluations": len(history), "curriculum_stage_idx": int(cfg.curriculum_stage_idx), "wall_s": round(total, 2),
                      "what": "jsrl_w_iql.train_loop on a trivial env: includes eval rollouts (2 episodes x 100 steps every 1000 "
                              "iterations), guide / learner arbitration, exploration, ring inserts, sample + train every step"}))


if __name__ == "__main__":
    main()
