"""Rate of the reference-shaped loop `batch = rb.sample(B); log = trainer.train(batch)` through the drop-in
facade (S = 1, K = 1, one host sync per step), next to the fused K-step ensemble call."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import jsrl_corl_b200 as J  # noqa: E402
from jsrl_corl_b200.synthetic import synthetic_dataset  # noqa: E402


def run(S, A, L, det, beta, steps=600, warm=100, B=256):
    torch.manual_seed(0)
    q, v = J.TwinQ(S, A, 256, L).cuda(), J.ValueFunction(S, 256, L).cuda()
    actor = (J.DeterministicPolicy if det else J.GaussianPolicy)(S, A, 1.0, 256, L).cuda()
    opts = [torch.optim.Adam(m.parameters(), lr=3e-4) for m in (actor, q, v)]
    tr = J.ImplicitQLearning(1.0, actor, opts[0], q, opts[1], v, opts[2], beta=beta, max_steps=10 ** 6, device="cuda")
    rb = J.ReplayBuffer(S, A, 1_000_000, "cuda")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        rb.load_d4rl_dataset(synthetic_dataset(1_000_000, S, A, 0))
    np.random.seed(0)
    for _ in range(warm):
        tr.train(rb.sample(B))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        log = tr.train(rb.sample(B))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return steps / dt, log


if __name__ == "__main__":
    for name, args in {"hopper 2x256 det": (11, 3, 2, True, 3.0), "antmaze 3x256 gauss": (29, 8, 3, False, 10.0)}.items():
        r, log = run(*args)
        print(f"{name}: {r:.0f} sample+train steps/s through ReplayBuffer.sample + ImplicitQLearning.train  last={log}")
