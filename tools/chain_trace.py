"""IQL_CHAIN_TRACE=1 python tools/chain_trace.py [members]: where the time of a bwd_chain task goes (CTA pair 0)."""
import ctypes as C, os, sys
os.environ["IQL_B200_DEBUG"] = "1"
os.environ["IQL_CHAIN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jsrl_corl_b200 import IQLEnsemble, ReplayBuffer, _lib
from jsrl_corl_b200.synthetic import synthetic_dataset
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ens = IQLEnsemble(S, 17, 6, 256, 2, 256, math_mode="tf32", seeds=list(range(S)), max_steps_per_call=4, init=False)
ens.init_member(0, 0); ens.engine.params[1:] = ens.engine.params[0]; ens.engine.target[1:] = ens.engine.target[0]
rb = ReplayBuffer(17, 6, 200000, "cuda"); rb.load_d4rl_dataset(synthetic_dataset(200000, 17, 6, 0)); ens.bind_replay(rb)
for _ in range(3): ens.train_steps(1)
torch.cuda.synchronize()
NP, NT = 9, 4
buf = (C.c_longlong * (3 * NT * NP * 4))()
n = _lib.lib().iql_debug_chain_trace(buf, len(buf))
a = np.array(buf[:], dtype=np.int64).reshape(3, NT, NP, 4)
t0 = a[a > 0].min()
names = ["TMA", "MMA", "EPI"]
for ti in range(NT):
    for r in range(3):
        for pi in range(NP):
            row = a[r, ti, pi]
            if row.max() > 0:
                print(f"task {ti} {names[r]} phase {pi}: " + " ".join(f"{(x - t0) / 1e3:8.2f}" if x > 0 else "       -" for x in row))
