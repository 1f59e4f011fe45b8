#!/usr/bin/env python
"""us per env step of `policy.act(state, device)` (finetune/iql.py:371-379, 403-413): the engine's fused act kernel with
pinned staging (what the JSRL loop and eval_actor use once the policy is engine-backed) vs the same call in stock torch
(tensor from numpy -> 3 Linear + activations -> clamp -> .cpu().numpy())."""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from jsrl_corl_b200 import iql as facade


def main():
    out = {}
    for name, S, A, L, det in (("hopper_det_2x256", 11, 3, 2, True), ("antmaze_gauss_3x256", 29, 8, 3, False)):
        torch.manual_seed(0)
        q, v = facade.TwinQ(S, A, 256, L), facade.ValueFunction(S, 256, L)
        actor = (facade.DeterministicPolicy if det else facade.GaussianPolicy)(S, A, 1.0, 256, L)
        tr = facade.ImplicitQLearning(1.0, actor, torch.optim.Adam(actor.parameters(), lr=3e-4), q, torch.optim.Adam(q.parameters(), lr=3e-4),
                                      v, torch.optim.Adam(v.parameters(), lr=3e-4), device="cuda")
        g = torch.Generator().manual_seed(0)
        batch = [torch.randn(256, S, generator=g).cuda(), (torch.rand(256, A, generator=g) * 2 - 1).cuda(), torch.randn(256, 1, generator=g).cuda(),
                 torch.randn(256, S, generator=g).cuda(), torch.zeros(256, 1).cuda()]
        tr.train(batch)
        actor.eval()
        stock = copy.deepcopy(actor).eval()  # owns its parameters: stock torch path
        states = np.random.RandomState(0).randn(2000, S).astype(np.float32)
        res = {}
        for label, pol in (("engine_act_kernel", actor), ("stock_torch", stock)):
            for s in states[:200]:
                pol.act(s, "cuda")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for s in states:
                a = pol.act(s, "cuda")
            torch.cuda.synchronize()
            res[label + "_us"] = round((time.perf_counter() - t0) / len(states) * 1e6, 2)
        np.testing.assert_allclose(actor.act(states[0], "cuda"), stock.act(states[0], "cuda"), atol=2e-6)
        res["speedup"] = round(res["stock_torch_us"] / res["engine_act_kernel_us"], 2)
        out[name] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
