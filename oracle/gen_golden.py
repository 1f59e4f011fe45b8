"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the LIVE reference.

Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

Every fixture is produced by the unmodified reference classes
(``algorithms/finetune/iql.py``: ReplayBuffer, TwinQ, ValueFunction,
GaussianPolicy / DeterministicPolicy, ImplicitQLearning) on CPU, with
``torch.manual_seed`` initial weights, synthetic data from
``oracle.iql_numpy.synthetic_dataset`` and numpy's global MT19937 index stream.
The reference ships no golden vectors of its own, so
these files are what pins the oracle and the CUDA path.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.iql_numpy import synthetic_dataset  # noqa: E402
from oracle.ref_loader import load_reference_iql  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class _MaskDropout(torch.nn.Module):
    """Oracle-side hook: replaces nn.Dropout by a module fed with explicit keep-masks."""

    def __init__(self, p):
        super().__init__()
        self.p = p
        self.mask = None

    def forward(self, x):
        if not self.training:
            return x
        return x * self.mask.to(x.dtype) / (1.0 - self.p)


def flat_state(q, v, actor, prefix):
    out = {}
    for grp, mod in (("qf", q), ("vf", v), ("actor", actor)):
        for k, t in mod.state_dict().items():
            out[f"{prefix}/{grp}/{k}"] = t.detach().double().numpy().astype(np.float32 if t.dtype == torch.float32 else np.float64)
    return out


def flat_opt(opt, mod, grp, prefix):
    out = {}
    names = [n for n, _ in mod.named_parameters()]
    sd = opt.state_dict()["state"]
    for i, n in enumerate(names):
        if i in sd:
            out[f"{prefix}/{grp}/{n}/exp_avg"] = sd[i]["exp_avg"].numpy().copy()
            out[f"{prefix}/{grp}/{n}/exp_avg_sq"] = sd[i]["exp_avg_sq"].numpy().copy()
    return out


def run_reference(name, S, A, H, L, det, B, steps, snapshots, *, n_rows=10000, beta=3.0, iql_tau=0.7, tau=0.005,
                  discount=0.99, lr=3e-4, dropout=0.0, seed=0, idx_seed=1, antmaze=False, max_steps=None,
                  with_fp64=False, store_indices=True, lean=False):
    ref = load_reference_iql("finetune")
    max_steps = steps if max_steps is None else max_steps

    def build(dtype):
        torch.manual_seed(seed)
        q = ref.TwinQ(S, A, H, L)
        v = ref.ValueFunction(S, H, L)
        actor = (ref.DeterministicPolicy if det else ref.GaussianPolicy)(S, A, 1.0, H, L, dropout=dropout)
        if dtype == torch.float64:
            q.double(); v.double(); actor.double()
        vo = torch.optim.Adam(v.parameters(), lr=lr)
        qo = torch.optim.Adam(q.parameters(), lr=lr)
        ao = torch.optim.Adam(actor.parameters(), lr=lr)
        tr = ref.ImplicitQLearning(1.0, actor, ao, q, qo, v, vo, iql_tau=iql_tau, beta=beta, max_steps=max_steps,
                                   discount=discount, tau=tau, device="cpu")
        return q, v, actor, qo, vo, ao, tr

    data = synthetic_dataset(n_rows, S, A, 0, antmaze_rewards=antmaze)
    out = {}
    meta = dict(S=S, A=A, H=H, L=L, det=int(det), B=B, steps=steps, n_rows=n_rows, beta=beta, iql_tau=iql_tau, tau=tau,
                discount=discount, lr=lr, dropout=dropout, seed=seed, idx_seed=idx_seed, antmaze=int(antmaze),
                max_steps=max_steps)
    out["meta_keys"] = np.array(list(meta.keys()))
    out["meta_vals"] = np.array([float(v) for v in meta.values()], dtype=np.float64)

    results = {}
    for dtype in ([torch.float32, torch.float64] if with_fp64 else [torch.float32]):
        q, v, actor, qo, vo, ao, tr = build(dtype)
        drops = []
        if dropout > 0.0:
            seq = actor.net.net
            for i, mod in enumerate(seq):
                if isinstance(mod, torch.nn.Dropout):
                    seq[i] = _MaskDropout(dropout)
                    drops.append(seq[i])
        rb = ref.ReplayBuffer(S, A, n_rows, "cpu")
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            rb.load_d4rl_dataset(data)
        np.random.seed(idx_seed)
        mask_rng = np.random.RandomState(1234)
        if dtype == torch.float32:
            init = flat_state(q, v, actor, "init")
            if lean:  # init is reproducible from torch.manual_seed: keep a checksum only
                h = hashlib.sha256()
                for k in sorted(init):
                    h.update(init[k].tobytes())
                out["init_sha256"] = np.array(h.hexdigest())
            else:
                out.update(init)
        losses, indices, masks = [], [], []
        for t in range(1, steps + 1):
            state = np.random.get_state()
            batch = rb.sample(B)
            np.random.set_state(state)
            idx = np.random.randint(0, rb._size, size=B)
            indices.append(idx)
            if drops:
                mk = (mask_rng.uniform(size=(L, B, H)) >= dropout)
                masks.append(mk)
                for li, dmod in enumerate(drops):
                    dmod.mask = torch.from_numpy(mk[li])
            batch = [b.to(dtype) for b in batch]
            log = tr.train(batch)
            losses.append([log["value_loss"], log["q_loss"], log["actor_loss"]])
            if dtype == torch.float32 and t in snapshots:
                out.update(flat_state(q, v, actor, f"step{t}"))
                if not lean:
                    out.update({f"step{t}/q_target/{k}": x.numpy().copy() for k, x in tr.q_target.state_dict().items()})
                    for grp, mod, opt in (("qf", q, qo), ("vf", v, vo), ("actor", actor, ao)):
                        out.update(flat_opt(opt, mod, grp, f"step{t}/opt"))
                out[f"step{t}/actor_lr"] = np.float64(ao.param_groups[0]["lr"])
        results[dtype] = (np.array(losses, dtype=np.float64), {k: x.detach().double().numpy() for grp, mod in (("qf", q), ("vf", v), ("actor", actor)) for k, x in ((f"{grp}/{kk}", vv) for kk, vv in mod.state_dict().items())})
        if dtype == torch.float32:
            out["losses"] = np.array(losses, dtype=np.float32)
            idx_arr = np.array(indices, dtype=np.int64)
            out["indices_sha256"] = np.array(hashlib.sha256(idx_arr.tobytes()).hexdigest())
            if store_indices:
                out["indices"] = idx_arr.astype(np.int32)
            if masks:
                out["dropout_masks"] = np.packbits(np.array(masks, dtype=np.uint8).reshape(-1))
    if with_fp64:
        l32, w32 = results[torch.float32]
        l64, w64 = results[torch.float64]
        # the reference's own fp32-vs-fp64 divergence: the noise floor of this chaotic trajectory
        for k in w32:
            out[f"noise/{k}"] = np.float64(np.linalg.norm(w32[k] - w64[k]) / max(np.linalg.norm(w32[k]), 1e-30))
        out["noise/losses_rel"] = np.abs(l32 - l64) / np.maximum(np.abs(l32), 1e-30)
        out["losses_fp64"] = l64
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  final losses {out['losses'][-1]}")


def sampler_golden():
    """ReplayBuffer.sample of the reference (iql.py:171-178): numpy MT19937 index stream + gathers."""
    ref = load_reference_iql("finetune")
    out = {}
    for tag, (S, A, n, B) in {"hopper": (11, 3, 5000, 256), "pen": (45, 24, 777, 64), "one_row": (3, 2, 1, 8)}.items():
        data = synthetic_dataset(n, S, A, 3)
        rb = ref.ReplayBuffer(S, A, n + 5, "cpu")
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            rb.load_d4rl_dataset(data)
        np.random.seed(42)
        state = np.random.get_state()
        b1 = rb.sample(B)
        b2 = rb.sample(B)
        np.random.set_state(state)
        out[f"{tag}/indices"] = np.stack([np.random.randint(0, n, size=B), np.random.randint(0, n, size=B)]).astype(np.int64)
        for j, b in enumerate((b1, b2)):
            for nm, t in zip(("s", "a", "r", "s2", "d"), b):
                out[f"{tag}/batch{j}/{nm}"] = t.numpy().copy()
        out[f"{tag}/dims"] = np.array([S, A, n, B])
    path = os.path.join(OUT, "sampler.npz")
    np.savez_compressed(path, **out)
    print(f"sampler: {os.path.getsize(path) / 1e6:.2f} MB")


def resume_golden():
    """A checkpoint written by the reference (torch.save(trainer.state_dict())) after 20 steps, and the losses of
    the next 20 steps of a FRESH reference trainer that loaded it (load_state_dict re-clones q_target, iql.py:584)."""
    ref = load_reference_iql("finetune")
    S, A, H, L, B, n_rows = 11, 3, 64, 2, 32, 10000

    def build():
        torch.manual_seed(0)
        q, v = ref.TwinQ(S, A, H, L), ref.ValueFunction(S, H, L)
        actor = ref.GaussianPolicy(S, A, 1.0, H, L)
        vo, qo, ao = (torch.optim.Adam(m.parameters(), lr=3e-4) for m in (v, q, actor))
        return ref.ImplicitQLearning(1.0, actor, ao, q, qo, v, vo, max_steps=40, device="cpu")

    data = synthetic_dataset(n_rows, S, A, 0)
    rb = ref.ReplayBuffer(S, A, n_rows, "cpu")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        rb.load_d4rl_dataset(data)
    np.random.seed(1)
    tr = build()
    for _ in range(20):
        tr.train(rb.sample(B))
    path = os.path.join(OUT, "reference_checkpoint_19.pt")
    torch.save(tr.state_dict(), path)
    tr2 = build()
    tr2.load_state_dict(torch.load(path))
    losses = []
    for _ in range(20):
        log = tr2.train(rb.sample(B))
        losses.append([log["value_loss"], log["q_loss"], log["actor_loss"]])
    np.savez_compressed(os.path.join(OUT, "resume.npz"), losses_after_resume=np.array(losses, dtype=np.float32),
                        dims=np.array([S, A, H, L, B, n_rows]), total_it=np.array(tr2.total_it))
    print(f"resume: checkpoint {os.path.getsize(path) / 1e6:.2f} MB, final losses {losses[-1]}")


def late_snapshot(name, S, A, H, L, det, B, at, *, n_rows=10000, beta=3.0, iql_tau=0.7, tau=0.005, discount=0.99, lr=3e-4,
                  seed=0, idx_seed=1, antmaze=False, max_steps=None):
    """Teacher-forcing fixture for the LATE trajectory (nothing tight between step 40 and 1000): the full
    reference state after `at` free-running steps -- weights, target network, Adam moments, step counts, schedule epoch
    -- then ONE more reference step from it: its batch indices, its three losses and the post-step weights / target
    (every 16th element: enough to pin an independent one-step computation, 1/16 of the bytes).  At at = 999 with
    T_max = 1000 this is the large-step-count / bias-correction ~ 1 / cosine-LR ~ 0 regime."""
    ref = load_reference_iql("finetune")
    max_steps = at + 1 if max_steps is None else max_steps
    torch.manual_seed(seed)
    q, v = ref.TwinQ(S, A, H, L), ref.ValueFunction(S, H, L)
    actor = (ref.DeterministicPolicy if det else ref.GaussianPolicy)(S, A, 1.0, H, L)
    vo, qo, ao = (torch.optim.Adam(m.parameters(), lr=lr) for m in (v, q, actor))
    tr = ref.ImplicitQLearning(1.0, actor, ao, q, qo, v, vo, iql_tau=iql_tau, beta=beta, max_steps=max_steps, discount=discount,
                               tau=tau, device="cpu")
    data = synthetic_dataset(n_rows, S, A, 0, antmaze_rewards=antmaze)
    rb = ref.ReplayBuffer(S, A, n_rows, "cpu")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        rb.load_d4rl_dataset(data)
    np.random.seed(idx_seed)
    losses = []
    for _ in range(at):
        log = tr.train(rb.sample(B))
        losses.append([log["value_loss"], log["q_loss"], log["actor_loss"]])
    meta = dict(S=S, A=A, H=H, L=L, det=int(det), B=B, steps=at + 1, n_rows=n_rows, beta=beta, iql_tau=iql_tau, tau=tau,
                discount=discount, lr=lr, dropout=0.0, seed=seed, idx_seed=idx_seed, antmaze=int(antmaze), max_steps=max_steps)
    out = {"meta_keys": np.array(list(meta.keys())), "meta_vals": np.array([float(x) for x in meta.values()], dtype=np.float64)}
    pre = f"step{at}"
    out.update(flat_state(q, v, actor, pre))
    out.update({f"{pre}/q_target/{k}": x.numpy().copy() for k, x in tr.q_target.state_dict().items()})
    for grp, mod, opt in (("qf", q, qo), ("vf", v, vo), ("actor", actor, ao)):
        out.update(flat_opt(opt, mod, grp, f"{pre}/opt"))
    out[f"{pre}/actor_lr"] = np.float64(ao.param_groups[0]["lr"])
    out[f"{pre}/sched_epoch"] = np.int64(tr.actor_lr_schedule.last_epoch)
    state = np.random.get_state()
    batch = rb.sample(B)
    np.random.set_state(state)
    idx = np.random.randint(0, rb._size, size=B)
    log = tr.train(batch)
    losses.append([log["value_loss"], log["q_loss"], log["actor_loss"]])
    out["next_indices"] = idx.astype(np.int64)
    out["losses"] = np.array(losses, dtype=np.float32)
    out["indices_sha256"] = np.array(hashlib.sha256(idx.astype(np.int64).tobytes()).hexdigest())
    post = f"step{at + 1}"
    for k, a in flat_state(q, v, actor, post).items():
        out[k + "@16"] = a.reshape(-1)[::16].copy()
    for k, x in tr.q_target.state_dict().items():
        out[f"{post}/q_target/{k}@16"] = x.numpy().reshape(-1)[::16].copy()
    out[f"{post}/actor_lr"] = np.float64(ao.param_groups[0]["lr"])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  losses at step {at + 1}: {losses[-1]}  actor lr {ao.param_groups[0]['lr']:.3e}")


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--pen-only" in sys.argv:
        torch.set_num_threads(8)
        run_reference("pen_2x256_dropout", 45, 24, 256, 2, False, 256, 6, {6}, dropout=0.1, iql_tau=0.8, store_indices=False, lean=True)
        return
    if "--late-only" in sys.argv:
        torch.set_num_threads(8)
        late_snapshot("hopper_late_999", 11, 3, 256, 2, True, 256, 999, tau=0.001)
        late_snapshot("antmaze_late_999", 29, 8, 256, 3, False, 256, 999, beta=10.0, iql_tau=0.9, antmaze=True)
        return
    resume_golden()
    torch.set_num_threads(8)
    sampler_golden()
    # small Gaussian config with full optimizer state (teacher-forced single-step checks)
    run_reference("small_gauss", 11, 3, 64, 2, False, 32, 40, {1, 20, 40})
    run_reference("small_det", 5, 2, 32, 1, True, 16, 20, {1, 20})
    # dropout actor (pen-like, small) with injected masks
    run_reference("small_dropout", 45, 24, 64, 2, False, 32, 12, {1, 12}, dropout=0.1, iql_tau=0.8)
    # BASELINE configs[3] at full shape: pen-human (obs 45, act 24), dropout actor p = 0.1, injected masks
    run_reference("pen_2x256_dropout", 45, 24, 256, 2, False, 256, 6, {6}, dropout=0.1, iql_tau=0.8, store_indices=False, lean=True)
    # antmaze shape, 3 hidden layers, beta 10
    run_reference("antmaze_3x256", 29, 8, 256, 3, False, 256, 30, {30}, beta=10.0, iql_tau=0.9, antmaze=True, store_indices=False, lean=True)
    # halfcheetah shape, Gaussian, full width
    run_reference("halfcheetah_2x256", 17, 6, 256, 2, False, 256, 30, {30}, store_indices=False, lean=True)
    # config 1: hopper-medium, deterministic, polyak 0.001, 1000 steps with fp64 noise floor
    run_reference("hopper_1000", 11, 3, 256, 2, True, 256, 1000, {1000}, tau=0.001, with_fp64=True, store_indices=False, lean=True)
    # late-trajectory teacher-forcing snapshots (same runs as hopper_1000 / a 1000-step antmaze 3x256 beta 10 run)
    late_snapshot("hopper_late_999", 11, 3, 256, 2, True, 256, 999, tau=0.001)
    late_snapshot("antmaze_late_999", 29, 8, 256, 3, False, 256, 999, beta=10.0, iql_tau=0.9, antmaze=True)


if __name__ == "__main__":
    main()
