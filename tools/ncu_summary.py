"""Print the key roofline metrics of every kernel in an .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}
    print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|" + "---|" * len(WANT))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0][-48:]
        vals = []
        for key, _ in WANT:
            if key in idx:
                v = r[idx[key]]
                u = U[idx[key]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {u}".strip())
            else:
                vals.append("-")
        print(f"| {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
