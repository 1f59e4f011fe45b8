"""Timeline of the fused forward kernel on one SM (CTA 0, first 8 tiles): clock64 stamps of the TMA producer, the
MMA issuer and epilogue warp 2.  Run on the GPU box:  IQL_FUSED_TRACE=1 python tools/fused_trace.py [workload]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("IQL_B200_DEBUG", "1")
os.environ.setdefault("IQL_FUSED_TRACE", "1")

import bench  # noqa: E402
from jsrl_corl_b200 import _lib  # noqa: E402


def main():
    from jsrl_corl_b200 import IQLEnsemble, ReplayBuffer
    from jsrl_corl_b200.synthetic import synthetic_dataset

    name = sys.argv[1] if len(sys.argv) > 1 else "halfcheetah_ens64"
    w = bench.WORKLOADS[name]
    S = w["members"]
    ens = IQLEnsemble(S, w["S"], w["A"], w["H"], w["L"], w["B"], deterministic=w["det"], actor_dropout=w["dropout"],
                      math_mode="tf32", device="cuda", max_steps_per_call=8, seeds=list(range(S)), init=False)
    ens.init_member(0, 0)
    ens.engine.params[1:] = ens.engine.params[0]
    ens.engine.target[1:] = ens.engine.target[0]
    n_rows = 200_000
    rb = ReplayBuffer(w["S"], w["A"], n_rows, "cuda")
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        rb.load_d4rl_dataset(synthetic_dataset(n_rows, w["S"], w["A"], 0, antmaze_rewards=w["antmaze"]))
    ens.bind_replay(rb)
    ens.train_steps(8)
    torch.cuda.synchronize()
    ens.engine.profile_step(1)  # eager launches: the last fused launch wrote the trace
    torch.cuda.synchronize()
    buf = np.zeros(5 * 8 * 4 * 4, dtype=np.int64)
    n = _lib.lib().iql_debug_fused_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
    if n <= 0:
        raise SystemExit(f"no trace (rc={n})")
    t = buf.reshape(5, 8, 4, 4).astype(np.float64)
    L = w["L"]
    t0 = t[:4][t[:4] > 0].min()
    us = lambda x: (x - t0) / 1965.0  # noqa: E731  (SM clock 1965 MHz)
    print("times in us since the first stamp of CTA 0; one line per (tile, layer)")
    print(f"{'tile':>4} {'l':>2} | {'tma first':>9} {'tma last':>9} | {'mma start':>9} {'A ready':>9} {'kb0 full':>9} {'issued':>9} | "
          f"{'epi start':>9} {'acc full':>9} {'A stored':>9} {'epi end':>9} | first chunk: {'ld done':>8} {'math':>8} {'A st+arr':>8} {'TMA st':>8} {'bits':>8} {'ld 2nd':>8}")
    for ti in range(8):
        for l in range(L):
            p, m, e = t[0, ti, l], t[1, ti, l], t[2, ti, l]
            if m[0] == 0:
                continue
            print(f"{ti:>4} {l:>2} | {us(p[0]):9.2f} {us(p[1]):9.2f} | {us(m[0]):9.2f} {us(m[1]):9.2f} {us(m[2]):9.2f} {us(m[3]):9.2f} | "
                  f"{us(e[0]):9.2f} {us(e[1]):9.2f} {us(e[2]):9.2f} {us(e[3]):9.2f} |              "
                  f"{us(t[3, ti, l, 0]):8.2f} {us(t[3, ti, l, 1]):8.2f} {us(t[3, ti, l, 2]):8.2f} {us(t[3, ti, l, 3]):8.2f} "
                  f"{us(t[4, ti, l, 1]):8.2f} {us(t[4, ti, l, 0]):8.2f}")


    life = t[4, 7, 3]
    if life[0] > 0:
        print(f"kernel life of CTA 0 (us on the same axis): entry {us(life[0]):.2f}, set-up done {us(life[1]):.2f}, "
              f"predecessor done {us(life[2]):.2f}, exit {us(life[3]):.2f}")


if __name__ == "__main__":
    main()
