"""Checkpoint interchange in the other direction: a checkpoint written by the B200 engine's drop-in trainer
(tools/write_engine_checkpoint.py, run on the GPU box; fixture committed under tests/golden/) is loaded by the
UNMODIFIED reference trainer on CPU, which must continue with the same losses."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN


def test_reference_trainer_loads_engine_checkpoint():
    from oracle.iql_numpy import synthetic_dataset
    from oracle.ref_loader import load_reference_iql, reference_available

    ck = os.path.join(GOLDEN, "engine_checkpoint_19.pt")
    if not reference_available() or not os.path.exists(ck):
        pytest.skip("needs the reference tree and the engine-written fixture")
    ref = load_reference_iql("finetune")
    z = np.load(os.path.join(GOLDEN, "engine_checkpoint_next_losses.npz"))
    S, A, H, L, B, n_rows = [int(x) for x in z["dims"]]
    torch.manual_seed(99)
    q, v, actor = ref.TwinQ(S, A, H, L), ref.ValueFunction(S, H, L), ref.GaussianPolicy(S, A, 1.0, H, L)
    vo, qo, ao = (torch.optim.Adam(m.parameters(), lr=3e-4) for m in (v, q, actor))
    tr = ref.ImplicitQLearning(1.0, actor, ao, q, qo, v, vo, max_steps=40, device="cpu")
    sd = torch.load(ck, map_location="cpu")
    assert list(sd.keys()) == list(tr.state_dict().keys())
    tr.load_state_dict(sd)
    assert tr.total_it == 20 and tr.actor_lr_schedule.last_epoch == 20
    assert float(qo.state_dict()["state"][0]["step"]) == 20.0
    import contextlib, io
    rb = ref.ReplayBuffer(S, A, n_rows, "cpu")
    with contextlib.redirect_stdout(io.StringIO()):
        rb.load_d4rl_dataset(synthetic_dataset(n_rows, S, A, 0))
    np.random.seed(1)
    for _ in range(20):
        np.random.randint(0, n_rows, size=B)
    losses = []
    for _ in range(5):
        log = tr.train(rb.sample(B))
        losses.append([log["value_loss"], log["q_loss"], log["actor_loss"]])
    np.testing.assert_allclose(np.array(losses), z["losses"], rtol=2e-5, atol=1e-8)
