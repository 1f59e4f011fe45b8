"""Where the time of the S = 1 drop-in loop goes: wall-clock per component of `rb.sample(B); trainer.train(batch)`."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jsrl_corl_b200 import iql as facade, ReplayBuffer
from jsrl_corl_b200.synthetic import synthetic_dataset

S, A, B = 11, 3, 256
torch.manual_seed(0)
q, v, actor = facade.TwinQ(S, A), facade.ValueFunction(S), facade.DeterministicPolicy(S, A, 1.0)
tr = facade.ImplicitQLearning(1.0, actor, torch.optim.Adam(actor.parameters(), lr=3e-4), q, torch.optim.Adam(q.parameters(), lr=3e-4),
                              v, torch.optim.Adam(v.parameters(), lr=3e-4), device="cuda")
rb = ReplayBuffer(S, A, 100000, "cuda"); rb.load_d4rl_dataset(synthetic_dataset(100000, S, A, 0))
for _ in range(300): tr.train(rb.sample(B))
torch.cuda.synchronize()
n = 3000
t_s = t_t = 0.0
for _ in range(n):
    t0 = time.perf_counter(); b = rb.sample(B); t1 = time.perf_counter(); tr.train(b); t2 = time.perf_counter()
    t_s += t1 - t0; t_t += t2 - t1
print(f"sample {t_s / n * 1e6:.1f} us   train {t_t / n * 1e6:.1f} us   total {(t_s + t_t) / n * 1e6:.1f} us  -> {n / (t_s + t_t):.0f} steps/s")
# GPU-only time of the same step: K = 1 calls without the host sync
eng = tr._engine
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b = rb.sample(B)
e0.record()
for _ in range(500): eng.train_on_batch(b)
e1.record(); torch.cuda.synchronize()
print(f"GPU time of load_batch + 1 step (back to back, no sync): {e0.elapsed_time(e1) / 500 * 1e3:.1f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): tr.train(rb.sample(B))
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
