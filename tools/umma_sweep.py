"""Bring-up aid: run the tcgen05 GEMM self-test for each operand-layout mode and,
for the MN-major modes, over candidate descriptor encodings (env overrides read
by umma_gemm.cu).  Prints the relative error of every variant."""
import itertools
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
from test_gpu_umma import _run
import torch
for mode in (0, 1, 2):
    try:
        err, out, ref = _run(mode, 256, 256, 256)
        print("mode", mode, "rel_err %%.3e" %% err, "nan" if not torch.isfinite(out).all() else "")
    except Exception as e:
        print("mode", mode, "FAILED", repr(e)[:200])
''' % (ROOT, ROOT)


def main():
    variants = [dict()]
    for layout, swz in ((1, 4), (2, 3)):
        for lbo, sbo in ((4096, 512), (512, 4096), (4096, 1024), (1024, 4096), (128, 4096), (4096, 128)):
            variants.append({"IQL_UMMA_MN_LAYOUT": str(layout), "IQL_UMMA_MN_TMA_SWIZZLE": str(swz),
                             "IQL_UMMA_MN_LBO": str(lbo >> 4), "IQL_UMMA_MN_SBO": str(sbo >> 4)})
    for v in variants:
        env = dict(os.environ, **v)
        try:
            out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
            txt = out.stdout.strip().replace("\n", " | ") + (" ERR: " + out.stderr.strip()[-300:] if out.returncode else "")
        except subprocess.TimeoutExpired:
            txt = "TIMEOUT"
        print(v or "default", "->", txt, flush=True)
        if not v and "FAILED" not in txt and all(float(t.split()[0]) < 2e-3 for t in txt.split("rel_err ")[1:]):
            print("default encoding passes all modes; sweep skipped")
            return


if __name__ == "__main__":
    main()
