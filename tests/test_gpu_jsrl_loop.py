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


def _make_trainer(det, dropout=0.0, S=5, A=3, device="cuda"):
    from jsrl_corl_b200 import iql as facade

    torch.manual_seed(3)
    q, v = facade.TwinQ(S, A), facade.ValueFunction(S)
    actor = (facade.DeterministicPolicy if det else facade.GaussianPolicy)(S, A, 0.8, dropout=dropout)
    return facade.ImplicitQLearning(0.8, actor, torch.optim.Adam(actor.parameters(), lr=3e-4), q,
                                    torch.optim.Adam(q.parameters(), lr=3e-4), v, torch.optim.Adam(v.parameters(), lr=3e-4),
                                    device=device, max_steps=100)


def _batch(S=5, A=3, B=32, device="cuda"):
    g = torch.Generator().manual_seed(0)
    return [torch.randn(B, S, generator=g).to(device), torch.rand(B, A, generator=g).to(device) * 2 - 1,
            torch.randn(B, 1, generator=g).to(device), torch.randn(B, S, generator=g).to(device), torch.zeros(B, 1, device=device)]


@pytest.mark.parametrize("det", [True, False])
def test_policy_act_runs_on_the_engine_kernel(det):
    """once the module aliases the engine arena, ``act`` (iql.py:371-379, 403-413) is the fused
    act kernel -- same numbers as the stock torch forward on the same (updated) parameters."""
    trainer = _make_trainer(det)
    for _ in range(3):
        trainer.train(_batch())
    actor, eng = trainer.actor, trainer._engine
    actor.eval()
    state = np.linspace(-1, 1, 5).astype(np.float32)
    calls0 = eng.act_calls
    a = actor.act(state, "cuda")
    assert eng.act_calls == calls0 + 1 and a.shape == (3,) and a.dtype == np.float32
    with torch.no_grad():
        out = actor(torch.tensor(state.reshape(1, -1), device="cuda"))
        mean = out if det else out.mean
        want = torch.clamp(0.8 * mean, -0.8, 0.8).cpu().numpy().flatten()
    np.testing.assert_allclose(a, want, rtol=0, atol=2e-6)
    actor.train()
    if not det:
        # training mode (`dist.sample()`, iql.py:377): mean and std come from the engine kernel, the sample is drawn on the
        # host from torch's CPU generator -- reproducible under torch.manual_seed, distributed as Normal(mean, std)
        calls1 = eng.act_calls
        torch.manual_seed(7)
        s1 = actor.act(state, "cuda")
        assert eng.act_calls == calls1 + 1 and s1.shape == (3,) and s1.dtype == np.float32
        torch.manual_seed(7)
        assert np.array_equal(actor.act(state, "cuda"), s1)
        m_eng, std_eng = eng.act_host_gaussian(0, state)
        with torch.no_grad():
            dist = actor(torch.tensor(state.reshape(1, -1), device="cuda"))
        np.testing.assert_allclose(m_eng, dist.mean.cpu().numpy().flatten(), rtol=0, atol=2e-6)
        np.testing.assert_allclose(std_eng, dist.stddev.cpu().numpy().flatten(), rtol=1e-6)
        torch.manual_seed(7)
        eps = torch.randn(3).numpy()
        np.testing.assert_allclose(s1, np.clip(0.8 * (m_eng + std_eng * eps), -0.8, 0.8), rtol=1e-6, atol=1e-7)
        draws = np.stack([actor.act(state, "cuda") for _ in range(400)])
        inside = np.abs(0.8 * (m_eng + 3 * std_eng)) < 0.8  # components whose +-3 sigma range is not clipped
        if inside.any():
            assert np.all(np.abs(draws.mean(0)[inside] - 0.8 * m_eng[inside]) < 0.25 * 0.8 * std_eng[inside])
        # with active dropout the stock torch path keeps the masks
        drop = _make_trainer(False, dropout=0.1)
        drop.train(_batch())
        drop.actor.train()
        c0 = drop._engine.act_calls
        drop.actor.act(state, "cuda")
        assert drop._engine.act_calls == c0
    import copy
    clone = copy.deepcopy(actor).eval()  # a copy owns its parameters: no engine behind it
    assert clone._engine_ref is None
    np.testing.assert_allclose(clone.act(state, "cuda"), want, rtol=0, atol=2e-6)


def test_jsrl_loop_issues_act_kernel_launches():
    from jsrl_corl_b200 import ReplayBuffer
    from jsrl_corl_b200.jsrl_utils import JsrlTrainConfig
    from jsrl_corl_b200.jsrl_w_iql import train_loop
    from jsrl_corl_b200.engine import EnsembleEngine

    cfg = JsrlTrainConfig(device="cuda", env="Point-v0", seed=0, eval_freq=40, n_episodes=1, offline_iterations=40,
                          online_iterations=40, batch_size=32, n_curriculum_stages=2, rolling_mean_n=1, tolerance=0.5,
                          horizon_fn="time_step", online_buffer_size=100, iql_deterministic=True)
    rng = np.random.RandomState(1)
    n = 200
    data = {"observations": rng.uniform(-1, 1, (n, 3)).astype(np.float32), "actions": rng.uniform(-1, 1, (n, 2)).astype(np.float32),
            "rewards": rng.uniform(-1, 0, n).astype(np.float32), "next_observations": rng.uniform(-1, 1, (n, 3)).astype(np.float32),
            "terminals": np.zeros(n, bool)}
    rb = ReplayBuffer(3, 2, n, "cuda")
    rb.load_d4rl_dataset(data)
    count = {"n": 0}
    orig = EnsembleEngine.act_host

    def counting(self, *a, **k):
        count["n"] += 1
        return orig(self, *a, **k)

    EnsembleEngine.act_host = counting
    try:
        train_loop(cfg, PointEnv(0), PointEnv(1), rb, 3, 2, 1.0, max_steps=PointEnv.T)
    finally:
        EnsembleEngine.act_host = orig
    # guide-only probe (1 episode) + two evaluations (1 episode each) + guide steps of the online phase: every one of
    # them an engine act-kernel launch (learner steps with deterministic policies too)
    assert count["n"] >= 3 * PointEnv.T


def test_engine_on_a_non_current_device():
    """device='cuda:N' while the process's current device is another one."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(0)
    trainer = _make_trainer(True, device="cuda:1")
    logs = [trainer.train(_batch(device="cuda:1")) for _ in range(3)]
    assert torch.cuda.current_device() == 0 and all(np.isfinite(v) for d in logs for v in d.values())
    ref = _make_trainer(True, device="cuda:0")
    logs0 = [ref.train(_batch(device="cuda:0")) for _ in range(3)]
    for a, b in zip(logs, logs0):
        for k in a:
            assert a[k] == b[k]  # same kernels, same inputs: bit-identical across devices
    # the host-caller paths (by-value sample / insert, host step from sample()'s indices, act through pinned memory)
    from jsrl_corl_b200 import ReplayBuffer
    outs = []
    for dev, tr in (("cuda:1", trainer), ("cuda:0", ref)):
        rb = ReplayBuffer(5, 3, 64, dev)
        rng = np.random.RandomState(0)
        for i in range(40):
            rb.add_transition(rng.randn(5), rng.uniform(-1, 1, 3), float(i), rng.randn(5), i % 5 == 0)
        np.random.seed(4)
        log = [tr.train(rb.sample(32)) for _ in range(3)]
        tr.actor.eval()
        act = tr.actor.act(np.array([0.1, -0.2, 0.3, 0.4, -0.5], np.float32), dev)
        assert tr._path_counts[0] == 3 and torch.cuda.current_device() == 0
        outs.append((log, act))
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])


def test_wide_row_insert_and_unsynchronised_inserts():
    from jsrl_corl_b200 import ReplayBuffer

    S, A = 600, 8  # row_floats > 1024: the insert kernel loops over the row
    rb = ReplayBuffer(S, A, 16, "cuda")
    rng = np.random.RandomState(0)
    rows = []
    for i in range(20):  # wraps the ring; consecutive inserts reuse the two pinned staging rows
        s, a, s2 = rng.randn(S).astype(np.float32), rng.randn(A).astype(np.float32), rng.randn(S).astype(np.float32)
        rb.add_transition(s, a, float(i), s2, i % 3 == 0)
        rows.append((s, a, float(i), s2, float(i % 3 == 0)))
    torch.cuda.synchronize()
    assert rb._size == 16 and rb._pointer == 4
    for slot in range(16):
        s, a, r, s2, d = rows[slot + 16] if slot < 4 else rows[slot]
        np.testing.assert_array_equal(rb._states[slot].cpu().numpy(), s)
        np.testing.assert_array_equal(rb._actions[slot].cpu().numpy(), a)
        np.testing.assert_array_equal(rb._next_states[slot].cpu().numpy(), s2)
        assert float(rb._rewards[slot]) == r and float(rb._dones[slot]) == d


def test_optimizer_state_republished_after_loading_an_empty_checkpoint():
    """a step-0 checkpoint loaded into a stepped trainer must not leave `_published` stale."""
    fresh = _make_trainer(False)
    sd0 = fresh.state_dict()
    trainer = _make_trainer(False)
    for _ in range(2):
        trainer.train(_batch())
    trainer.load_state_dict(sd0)
    assert trainer.q_optimizer.state_dict()["state"] == {}
    trainer.train(_batch())
    st = trainer.q_optimizer.state_dict()["state"]
    assert len(st) == 12 and float(st[0]["step"]) == 1.0 and st[0]["exp_avg"].abs().sum() > 0
