"""Fine-grained wall-clock of ReplayBuffer.sample() pieces (host side), S = 1 drop-in loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from jsrl_corl_b200 import ReplayBuffer, _lib
from jsrl_corl_b200.synthetic import synthetic_dataset
S, A, B = 11, 3, 256
rb = ReplayBuffer(S, A, 100000, "cuda"); rb.load_d4rl_dataset(synthetic_dataset(100000, S, A, 0))
for _ in range(200): rb.sample(B)
torch.cuda.synchronize()
N = 3000
def timeit(f):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(N): f()
    dt = (time.perf_counter() - t0) / N * 1e6; torch.cuda.synchronize(); return dt
print("sample()                         %.1f us" % timeit(lambda: rb.sample(B)))
print("torch.empty                      %.1f us" % timeit(lambda: torch.empty(B * (2 * S + A + 2), dtype=torch.float32, device=rb._device)))
flat = torch.empty(B * (2 * S + A + 2), dtype=torch.float32, device=rb._device)
def views():
    s, a, r, s2, d = flat.split((B * S, B * A, B, B * S, B))
    return [s.view(B, S), a.view(B, A), r.view(B, 1), s2.view(B, S), d.view(B, 1)]
print("split + 5 views                  %.1f us" % timeit(views))
print("np.random.randint                %.1f us" % timeit(lambda: np.random.randint(0, 100000, size=B)))
pin = torch.empty(B, dtype=torch.int64).pin_memory(); dev = torch.empty(B, dtype=torch.int64, device="cuda")
idxh = np.random.randint(0, 100000, size=B)
def stage():
    pin.numpy()[:] = idxh
print("pin.numpy()[:] = idx             %.1f us" % timeit(stage))
print("idx.copy_(pin, non_blocking)     %.1f us" % timeit(lambda: dev.copy_(pin, non_blocking=True)))
ev = torch.cuda.Event()
print("ev.record()                      %.1f us" % timeit(lambda: ev.record()))
print("ev.synchronize()                 %.1f us" % timeit(lambda: ev.synchronize()))
print("raw stream query                 %.1f us" % timeit(lambda: torch._C._cuda_getCurrentRawStream(0)))
print("current_device                   %.1f us" % timeit(lambda: torch.cuda.current_device()))
out = views(); L = _lib.lib()
def launch():
    _lib.check(L.iql_replay_sample(rb._rows.data_ptr(), rb._lay_ref, 100000, B, dev.data_ptr(), 0, 0, flat.data_ptr(), out[1].data_ptr(),
                                   out[2].data_ptr(), out[3].data_ptr(), out[4].data_ptr(), None, torch._C._cuda_getCurrentRawStream(0)), None, "x")
print("ctypes iql_replay_sample launch  %.1f us" % timeit(launch))
