#!/usr/bin/env python
"""Experiment: does splitting the ensemble into G independent member groups, each with its own engine, CUDA
graph and stream, buy anything?  Members never exchange data, so the groups are free to run out of phase: the
tensor-bound fused forward of one group can overlap the HBM-bound backward / optimizer of another, and the
partially filled last wave of a persistent kernel is covered by the other group's CTAs.

    python tools/group_overlap.py [--groups 1 2 4] [--members 64] [--inner 50] [--steps 10]

Prints one line per G: steps/s summed over all members (CUDA events on a stream that joins all groups)."""
import argparse
import contextlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from bench import N_ROWS, WORKLOADS
    from jsrl_corl_b200 import IQLEnsemble, ReplayBuffer
    from jsrl_corl_b200.synthetic import synthetic_dataset

    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--workload", default="halfcheetah_ens64")
    ap.add_argument("--members", type=int, default=0)
    ap.add_argument("--inner", type=int, default=50)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rows", type=int, default=N_ROWS)
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    S_total = args.members or w["members"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    data = synthetic_dataset(args.rows, w["S"], w["A"], 0, antmaze_rewards=w["antmaze"])
    rb = ReplayBuffer(w["S"], w["A"], args.rows, dev)
    with contextlib.redirect_stdout(sys.stderr):
        rb.load_d4rl_dataset(data)
    for G in args.groups:
        sizes = [S_total // G + (1 if g < S_total % G else 0) for g in range(G)]
        groups, first = [], 0
        for n in sizes:
            seeds = list(range(first, first + n))
            hp = [dict(beta=w["beta"], iql_tau=w["iql_tau"], tau=w["tau"], cosine_t_max=1_000_000) for _ in seeds]
            ens = IQLEnsemble(n, w["S"], w["A"], w["H"], w["L"], w["B"], deterministic=w["det"], actor_dropout=w["dropout"],
                              math_mode="tf32", device=dev, max_steps_per_call=args.inner, seeds=seeds, hparams=hp, init=False)
            ens.init_member(0, seeds[0])
            ens.engine.params[1:] = ens.engine.params[0]
            ens.engine.target[1:] = ens.engine.target[0]
            ens.bind_replay(rb)
            out = torch.empty(n, args.inner, 3, dtype=torch.float32, device=dev)
            groups.append((ens, torch.cuda.Stream(device=dev), out))
            first += n
        torch.cuda.synchronize(dev)

        def round_():
            for ens, st, out in groups:
                with torch.cuda.stream(st):
                    ens.engine.train_steps(args.inner, out=out)

        for _ in range(3):
            round_()
        torch.cuda.synchronize(dev)
        main_st = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_st)
        for _, st, _ in groups:
            st.wait_stream(main_st)
        for _ in range(args.steps):
            round_()
        for _, st, _ in groups:
            main_st.wait_stream(st)
        e1.record(main_st)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        finite = all(bool(torch.isfinite(o).all()) for _, _, o in groups)
        print(f"groups={G} sizes={sizes} {S_total * args.inner * args.steps / (ms * 1e-3):.0f} steps/s "
              f"({ms / args.steps:.3f} ms per {args.inner}-step round) finite={finite}", flush=True)
        del groups
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
