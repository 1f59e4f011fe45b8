"""Diagnostic: error growth of the TF32 (tcgen05) and FP32 (SIMT) paths against the numpy
oracle along a free-running trajectory (same init, same index stream)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Golden, batch_from, tree_max_rel  # noqa: E402
from oracle.iql_numpy import NumpyIQL  # noqa: E402
from test_gpu_parity import _cpu_tree, _make_engine  # noqa: E402


def main(name="halfcheetah_2x256", steps=30):
    g = Golden(name)
    data, idx = g.dataset(), g.indices()
    orc = NumpyIQL(g.oracle_config(), g.init_tree(), np.float32)
    engs = {m: _make_engine(g, m)[0] for m in ("fp32", "tf32")}
    for t in range(steps):
        lo = orc.train(batch_from(data, idx[t]))
        ref = np.array([lo["value_loss"], lo["q_loss"], lo["actor_loss"]])
        line = f"step {t + 1:3d}"
        for m, eng in engs.items():
            ii = torch.from_numpy(idx[t:t + 1]).unsqueeze(0).contiguous()
            l = eng.train_steps(1, mode="indices", indices=ii).cpu().numpy()[0, 0]
            rel = np.abs(l - ref) / np.abs(ref)
            worst, where = tree_max_rel(_cpu_tree(eng.param_views(0)), orc.state())
            line += f" | {m}: loss rel {rel[0]:.1e} {rel[1]:.1e} {rel[2]:.1e} w {worst:.1e} ({where})"
        if t < 5 or (t + 1) % 5 == 0:
            print(line, flush=True)


if __name__ == "__main__":
    main(*(sys.argv[1:2]))
