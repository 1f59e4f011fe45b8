"""Run on the GPU box: train the drop-in trainer for 20 steps, write its checkpoint in the reference's format to
gpurun_out/engine_checkpoint_19.pt, reload it into the same trainer (q_target <- qf, as the reference's
load_state_dict does) and record the losses of the next 5 steps.  tests/test_checkpoint_interop.py then loads the
file into the REFERENCE trainer on CPU and must reproduce those losses."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jsrl_corl_b200 as J  # noqa: E402
from jsrl_corl_b200.synthetic import synthetic_dataset  # noqa: E402

S, A, H, L, B, n_rows = 11, 3, 64, 2, 32, 10000
torch.manual_seed(0)
q, v, actor = J.TwinQ(S, A, H, L).cuda(), J.ValueFunction(S, H, L).cuda(), J.GaussianPolicy(S, A, 1.0, H, L).cuda()
vo, qo, ao = (torch.optim.Adam(m.parameters(), lr=3e-4) for m in (v, q, actor))
tr = J.ImplicitQLearning(1.0, actor, ao, q, qo, v, vo, max_steps=40, device="cuda", math_mode="fp32")
rb = J.ReplayBuffer(S, A, n_rows, "cuda")
rb.load_d4rl_dataset(synthetic_dataset(n_rows, S, A, 0))
np.random.seed(1)
for _ in range(20):
    tr.train(rb.sample(B))
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
path = os.path.join(out, "engine_checkpoint_19.pt")
sd = tr.state_dict()
cpu = lambda o: ({k: cpu(x) for k, x in o.items()} if isinstance(o, dict) else [cpu(x) for x in o] if isinstance(o, list)
                 else o.detach().cpu().clone() if torch.is_tensor(o) else o)
torch.save(cpu(sd), path)
tr.load_state_dict(torch.load(path, map_location="cuda"))
losses = []
for _ in range(5):
    log = tr.train(rb.sample(B))
    losses.append([log["value_loss"], log["q_loss"], log["actor_loss"]])
np.savez(os.path.join(out, "engine_checkpoint_next_losses.npz"), losses=np.array(losses, dtype=np.float64),
         dims=np.array([S, A, H, L, B, n_rows]))
print("wrote", path, losses[-1])
