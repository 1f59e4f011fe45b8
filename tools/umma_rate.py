"""Tensor-pipe rate of the tcgen05 TF32 GEMM building block per operand layout (mode 0 K-major x K-major,
1 K-major x MN-major, 2 MN-major x MN-major), single CTA vs CTA pair, on one large compute-bound problem.
Run on the GPU box: python tools/umma_rate.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsrl_corl_b200 import _lib  # noqa: E402


def run(mode, M, N, K, single, reps=5):
    L = _lib.lib()
    A = torch.randn((M, K) if mode != 2 else (K, M), device="cuda")
    B = torch.randn((N, K) if mode == 0 else (K, N), device="cuda")
    Cout = torch.empty(M, N, device="cuda")
    scratch = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.iql_selftest_umma_gemm(mode | (0x100 if single else 0), M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(),
                                      B.stride(0), Cout.data_ptr(), N, scratch.data_ptr(), scratch.numel(), st.cuda_stream)
        _lib.check(rc, None, "selftest")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * M * N * K / (best * 1e-3) / 1e12, best


if __name__ == "__main__":
    for (M, N, K) in ((8192, 4096, 4096), (4096, 4096, 1024), (16384, 256, 256)):
        for mode in (0, 1, 2):
            for single in (True, False):
                tf, ms = run(mode, M, N, K, single)
                print(f"M={M} N={N} K={K} mode={mode} {'single' if single else 'pair  '}: {tf:7.1f} TFLOP/s  ({ms * 1e3:.0f} us)", flush=True)
    a = torch.randn(8192, 4096, device="cuda")
    b = torch.randn(4096, 4096, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = True
    for _ in range(3):
        a @ b
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    print(f"cuBLAS TF32 8192x4096x4096: {5 * 2.0 * 8192 * 4096 * 4096 / (e0.elapsed_time(e1) * 1e-3) / 1e12:.1f} TFLOP/s")
