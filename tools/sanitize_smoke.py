"""Tiny engine run for compute-sanitizer (memcheck / racecheck / synccheck): 2 members, 2 steps, every tcgen05 path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jsrl_corl_b200 import IQLEnsemble, ReplayBuffer
from jsrl_corl_b200.synthetic import synthetic_dataset

path = sys.argv[1] if len(sys.argv) > 1 else "auto"
ens = IQLEnsemble(2, 17, 6, 256, 2, 256, math_mode="tf32", seeds=[1, 2], max_steps_per_call=2, step_path=path)
rb = ReplayBuffer(17, 6, 4096, "cuda")
rb.load_d4rl_dataset(synthetic_dataset(4096, 17, 6, 0))
ens.bind_replay(rb)
os.environ["IQL_B200_DEBUG"] = "1"
os.environ["IQL_B200_GRAPHS"] = "0"
out = ens.train_steps(2).cpu()
torch.cuda.synchronize()
print(path, ens.engine.paths, out[0, -1].tolist())
