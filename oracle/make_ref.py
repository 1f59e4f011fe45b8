"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the two UNMODIFIED reference files the hot path lives in

    /root/reference/algorithms/offline/iql.py   -> oracle/_ref/offline_iql.py
    /root/reference/algorithms/finetune/iql.py  -> oracle/_ref/finetune_iql.py

so that ``bench.py --impl reference`` can time the reference's own classes on the GPU box's host cores
(``cpu_baseline.kind == "reference"``) and on its GPU in stock eager mode (``torch_eager_b200``).  ``oracle/_ref/`` is
git-ignored (no reference source enters the history) but not gpurun-ignored, like the built ``.so`` files.  Called by
``__graft_entry__.build()``; a no-op where /root/reference does not exist (the GPU box uses the staged copy).
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("JSRL_REFERENCE_ROOT", "/root/reference")


def make_ref() -> bool:
    out = os.path.join(HERE, "_ref")
    done = False
    for variant in ("offline", "finetune"):
        src = os.path.join(REFERENCE_ROOT, "algorithms", variant, "iql.py")
        if not os.path.isfile(src):
            continue
        os.makedirs(out, exist_ok=True)
        dst = os.path.join(out, f"{variant}_iql.py")
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        with open(dst + ".sha256", "w") as f:
            f.write(f"{digest}  {src}\n")
        done = True
    return done


if __name__ == "__main__":
    print("staged" if make_ref() else "reference tree not present; nothing staged")
