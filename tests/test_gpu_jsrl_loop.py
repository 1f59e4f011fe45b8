"""Offline -> online JSRL loop on a fake env: curriculum bookkeeping, online ring
inserts, the update gate, fresh online learner, checkpoints -- all through the CUDA engine."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class PointEnv:
    """gym-style (4-tuple) toy env: 3-dim state (x, y, t/T), 2-dim action, reward = -|pos|, T = 20 steps."""
    T = 20

    def __init__(self, seed=0):
        self.rng = np.random.RandomState(seed)
        self.spec = type("Spec", (), {"id": "Point-v0"})()

    def seed(self, s):
        self.rng = np.random.RandomState(s)

    def reset(self):
        self.pos = self.rng.uniform(-1, 1, 2)
        self.t = 0
        return self._obs()

    def _obs(self):
        return np.array([self.pos[0], self.pos[1], self.t / self.T], dtype=np.float32)

    def step(self, action):
        self.pos = np.clip(self.pos + 0.1 * np.asarray(action, dtype=np.float64).reshape(-1)[:2], -2, 2)
        self.t += 1
        return self._obs(), float(-np.linalg.norm(self.pos)), self.t >= self.T, {}


def test_jsrl_offline_to_online_loop(tmp_path):
    from jsrl_corl_b200 import ReplayBuffer
    from jsrl_corl_b200.jsrl_utils import JsrlTrainConfig
    from jsrl_corl_b200.jsrl_w_iql import train_loop

    cfg = JsrlTrainConfig(device="cuda", env="Point-v0", seed=0, eval_freq=50, n_episodes=2, offline_iterations=40,
                          online_iterations=260, batch_size=32, n_curriculum_stages=3, rolling_mean_n=1, tolerance=0.5,
                          horizon_fn="time_step", online_buffer_size=100, checkpoints_path=str(tmp_path / "ckpt"),
                          iql_deterministic=True)
    rng = np.random.RandomState(1)
    n = 500
    data = {"observations": rng.uniform(-1, 1, (n, 3)).astype(np.float32), "actions": rng.uniform(-1, 1, (n, 2)).astype(np.float32),
            "rewards": rng.uniform(-1, 0, n).astype(np.float32), "next_observations": rng.uniform(-1, 1, (n, 3)).astype(np.float32),
            "terminals": np.zeros(n, bool)}
    torch.manual_seed(0)
    np.random.seed(0)
    rb = ReplayBuffer(3, 2, n, "cuda")
    rb.load_d4rl_dataset(data)
    logs = []
    trainer, cfg, history = train_loop(cfg, PointEnv(0), PointEnv(1), rb, 3, 2, 1.0, max_steps=PointEnv.T,
                                       log=lambda d, step: logs.append((step, d)))
    # offline phase: 40 updates with the offline trainer (its own total_it), then a FRESH learner whose
    # total_it continues from offline_iterations and has no LR schedule
    assert trainer.total_it == 40 + 260 and trainer.actor_lr_schedule is None
    offline = [d for _, d in logs if "offline_iter" in d]
    online = [d for _, d in logs if "online_iter" in d]
    assert len(offline) == 40 - 32  # reference quirk: the gate `t >= batch_size` uses the GLOBAL t
    assert len(online) == 260 and all(np.isfinite(d["q_loss"]) for d in online)
    # evaluations at t = 49, 99, ... ; curriculum metrics only in the online phase; stages linspace(init, 0, 3)
    assert [h["t"] for h in history] == [49, 99, 149, 199, 249, 299]
    assert all("eval/jsrl/curriculum_stage" in h for h in history)
    assert cfg.curriculum_stage_idx >= 1 and len(cfg.all_curriculum_stages) == 3 and cfg.all_curriculum_stages[-1] == 0
    assert np.isclose(cfg.all_curriculum_stages[0], (PointEnv.T - 1) / 2)  # mean step index of the guide-only eval
    assert sorted(os.listdir(tmp_path / "ckpt" / os.path.basename(cfg.checkpoints_path)) if False else os.listdir(cfg.checkpoints_path))[0].startswith("checkpoint_")
    sd = torch.load(os.path.join(cfg.checkpoints_path, "checkpoint_299.pt"), map_location="cuda")
    assert sd["total_it"] == 300 and sd["actor_lr_schedule"] == {}
    episodes = [d for d in online if "train/episode_length" in d]
    assert len(episodes) == 260 // PointEnv.T and all(d["train/episode_length"] == PointEnv.T for d in episodes)
