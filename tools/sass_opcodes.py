#!/usr/bin/env python
"""Per-kernel histogram of the Blackwell-specific SASS opcodes in libiql_b200.so (tcgen05 MMA / TMEM / TMA / cluster
barriers), from `cuobjdump -sass`.  Evidence that the GEMMs are tcgen05 + TMA code (UTCHMMA / UTMALDG / LDTM ...); the one
user of warp-level HMMA is the act_dim <= 8 policy head inside the fused forward's epilogue (16 m16n8k8 MMAs per chunk on an
operand the warp has just parked in shared memory -- a 32 x 8 output is below the smallest tcgen05 tile):

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "jsrl_corl_b200", "libiql_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTCCP", "SYNCS", "UBLKCP",
        "HMMA", "IMMA", "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "UCGABAR", "CCTL", "ACQBULK", "ELECT", "FFMA", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("iql::", "")
            if name.endswith(")"):  # drop the trailing parameter list (balanced scan: template arguments contain "(int)3")
                depth = 0
                for i in range(len(name) - 1, -1, -1):
                    depth += name[i] == ")"
                    depth -= name[i] == "("
                    if depth == 0:
                        name = name[:i]
                        break
            name = re.sub(r"^void ", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            cur[op] += 1
    arch = subprocess.run(["cuobjdump", "-lelf", SO], capture_output=True, text=True).stdout
    print(f"# {os.path.relpath(SO, ROOT)}: {len(kernels)} kernels; ELF images: {sorted(set(re.findall(r'sm_[0-9a-z]+', arch)))}")
    print("# counts of full opcodes (with modifiers) whose base mnemonic is in the watch list; _total = all instructions\n")
    for name, c in kernels.items():
        rows = [(op, n) for op, n in sorted(c.items()) if op != "_total" and any(op.split(".")[0] == k for k in KEYS[:13])]
        tc = sum(n for op, n in rows)
        print(f"## {name}  (instructions: {c['_total']}, tcgen05/TMA/TMEM opcodes: {tc})")
        for op, n in rows:
            print(f"    {op:48s} {n}")
        other = {k: sum(n for op, n in c.items() if op.split('.')[0] == k) for k in KEYS[13:]}
        print("    " + "  ".join(f"{k}:{v}" for k, v in other.items() if v))
        print()


if __name__ == "__main__":
    main()
