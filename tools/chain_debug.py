"""Debug aid: one update on two engines (per-phase backward vs chained backward), per-tensor differences."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from helpers import Golden
from jsrl_corl_b200 import EnsembleEngine, ReplayBuffer

name = sys.argv[1] if len(sys.argv) > 1 else "halfcheetah_2x256"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = Golden(name)
m = g.meta
engs = {}
for path in ("phases", "chain"):
    eng = EnsembleEngine(1, m["S"], m["A"], m["H"], m["L"], m["B"], bool(m["det"]), "tf32", "cuda", 64, step_path=path)
    rb = ReplayBuffer(m["S"], m["A"], m["n_rows"], "cuda")
    rb.load_d4rl_dataset(g.dataset())
    eng.load_params(0, g.init_tree(), dropout_keys=m["dropout"] > 0)
    eng.set_hparams(0, beta=m["beta"], iql_tau=m["iql_tau"], discount=m["discount"], tau=m["tau"], vf_lr=m["lr"], qf_lr=m["lr"],
                    actor_lr=m["lr"], actor_dropout=m["dropout"], cosine_t_max=m["max_steps"], seed=0)
    eng.bind_replay(0, rb.rows, m["n_rows"])
    idx = torch.from_numpy(g.indices()[:steps]).unsqueeze(0).contiguous()
    losses = eng.train_steps(steps, mode="indices", indices=idx).cpu().numpy()
    print(path, eng.paths, losses[0, -1])
    engs[path] = (eng, rb)
a, b = engs["phases"][0], engs["chain"][0]
def cmp(title, va, vb):
    for grp in va:
        for k in va[grp]:
            x, y = va[grp][k].cpu().numpy(), vb[grp][k].cpu().numpy()
            d = np.abs(x - y).max()
            rel = d / (np.abs(x).max() + 1e-30)
            flag = "  <<<<" if rel > 1e-5 else ""
            if flag or "-v" in sys.argv:
                bad = np.argwhere(np.abs(x - y) > 1e-5 * (np.abs(x).max() + 1e-30))
                print(f"{title:8s} {grp}/{k:24s} maxabs {d:.3e} rel {rel:.3e} nbad {len(bad)} first {bad[:3].tolist()} last {bad[-2:].tolist()}{flag}")
cmp("param", a.param_views(0), b.param_views(0))
ma, va = a.moment_views(0); mb, vb = b.moment_views(0)
cmp("exp_avg", ma, mb)
cmp("exp_sq", va, vb)
cmp("target", {"qf": a.target_views(0)}, {"qf": b.target_views(0)})
P = a.layout.param_floats
print("raw arenas equal:", torch.equal(a.params, b.params), torch.equal(a.exp_avg, b.exp_avg), torch.equal(a.target, b.target))
