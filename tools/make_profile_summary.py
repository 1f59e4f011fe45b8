"""Build profiles/<tag>_summary.md, profiles/<tag>_*.json and profiles/traffic.json from gpurun_out/ artifacts."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def last_json(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def main(tag="r01"):
    os.makedirs(P, exist_ok=True)
    b = last_json(os.path.join(G, f"bench_{tag}.json"))
    json.dump(b, open(os.path.join(P, f"{tag}_bench.json"), "w"))
    ref = last_json(os.path.join(G, f"bench_{tag}_ref.json"))
    json.dump(ref, open(os.path.join(P, f"{tag}_bench_reference_arm.json"), "w"))
    shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches.csv"))
    ls = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(P, f"{tag}_launches.csv")],
                        capture_output=True, text=True).stdout
    ncu = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"prof_{tag}.ncu-rep")],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_ncu_full_summary.md"), "w").write(ncu)
    # per-kernel DRAM traffic (dram read + write of one launch) keyed by the bench's kernel labels
    label_of = {"adam_polyak_kernel": "adam_polyak", "gather_kernel": "gather", "loss_kernel": "loss", "last_bwd_kernel": "last_bwd_wgrad", "last_bwd_v4_kernel": "last_bwd_wgrad",
                "fused_fwd_kernel": "fused_fwd"}
    traffic, plain = {}, []
    for line in ncu.splitlines()[2:]:
        cells = [c.strip() for c in line.strip("|").split("|")]
        if len(cells) < 4:
            continue
        name = cells[0]

        def mb(x):
            v, u = x.split()[:2]
            return float(v) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)

        tot = mb(cells[2]) + mb(cells[3])
        if "umma_gemm_kernel<2, 0" in name:
            traffic.setdefault("first_fwd_fwd", tot)
        elif "umma_gemm_kernel<2, 1" in name:
            traffic.setdefault("hidden_fwd", tot)
        elif "umma_gemm_kernel<3, 0" in name:
            traffic.setdefault("hidden_dgrad", tot)
        elif "umma_gemm_kernel<0, 0" in name:
            plain.append(tot)  # two wgrad launches per step: hidden layer (larger) and input layer
        else:
            for k, lab in label_of.items():
                if k in name and lab not in traffic:
                    traffic[lab] = tot
    if plain:
        traffic["hidden_wgrad"] = max(plain)
        traffic["first_wgrad_wgrad"] = min(plain)
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    rows = "\n".join(f"| {k['kernel']} | {k['us']} | {k['bound']} | {k['gbs']} | {k['tflops']} | {k['frac']} | "
                     f"{traffic.get(k['kernel'], 0) / 1e6:.0f} |" for k in b["kernels"])
    others = []
    for w in ("hopper_ens64", "pen_sweep256", "stress", "hopper_single", "antmaze_jsrl"):
        f = os.path.join(G, f"bench_{tag}_{w}.json")
        if os.path.exists(f):
            d = last_json(f)
            json.dump(d, open(os.path.join(P, f"{tag}_bench_{w}.json"), "w"))
            others.append(f"| {d['config']['workload']} | {d['config']['members_per_gpu']} | {d['config']['batch']} | {d['config']['hidden']} | "
                          f"{d['value']:.0f} | {d['e2e']['value']:.0f} | {100 * d['step_roofline']['whole_step_frac_of_tensor_peak']:.1f} % | "
                          f"{d['step_roofline']['whole_step_hbm_gbs']:.0f} |")
    sr = b["step_roofline"]
    ft = os.path.join(G, f"fused_trace_{tag}.txt")
    if os.path.exists(ft):
        shutil.copy(ft, os.path.join(P, f"{tag}_fused_trace.txt"))
    extra = ""
    for name, title in (("fused_probe", "Fused-forward epilogue probes (`tools/fused_probe.sh`, `IQL_FUSED_DBG`: 1 no sign bits, 2 no activation "
                                        "stores, 4 no staging wait, 8 no head / policy math, 15 all of them, 16 reversed tile order; results are "
                                        "wrong in these runs, only the times are read)"),
                        ("group_overlap", "Member groups on their own engines / graphs / streams (`tools/group_overlap.py`): splitting the "
                                          "64-member ensemble loses throughput, the persistent kernels of two streams queue instead of sharing SMs")):
        src = os.path.join(G, f"{name}_{tag}.txt")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, f"{tag}_{name}.txt"))
            body = "".join(l for l in open(src) if l.startswith(("dbg=", "groups=")))
            extra += f"## {title}\n\n```\n{body}```\n\n"
    dr = os.path.join(G, f"dropin_rate_{tag}.txt")
    dropin = ""
    if os.path.exists(dr):
        lines = [l.split("  last=")[0] for l in open(dr).read().splitlines() if "steps/s" in l]
        dropin = ("Drop-in loop `batch = rb.sample(256); log = trainer.train(batch)` (S = 1, K = 1, one host sync per step, "
                  "`tools/dropin_rate.py`): " + "; ".join(lines) + ".\n")
    md = f"""# Round 1 profile summary (B200, TF32 tcgen05 path)

Default bench workload = BASELINE.json configs[2]: halfcheetah-medium-replay shape (obs 17, act 6, 2x256, Gaussian
actor, batch 256), 64-member ensemble, 1M-row synthetic buffer, Philox sampling in-kernel, 50 update steps per
engine call (one CUDA graph launch).

## Bench lines (`{tag}_bench.json`, `{tag}_bench_reference_arm.json`)

* **{b['value']:.0f} gradient steps/s** summed over 64 members ({b['ms_per_step']:.2f} ms per 50-step call, {b['gpu_launches']} kernel launches in the timed region);
  **e2e {b['e2e']['value']:.0f} steps/s** (host-drawn int64 indices H2D {b['e2e']['h2d_bytes_per_step']} B + losses D2H {b['e2e']['d2h_bytes_per_step']} B and one sync per call).
* reference arm (`--impl reference`): numpy port of the reference update on the box's {ref['cpu_baseline']['cores']} host threads, one member:
  {ref['value']:.0f} steps/s ({ref['cpu_baseline']['sample']}).  `cpu_baseline` inside the GPU line: {b['cpu_baseline']['value']:.0f} steps/s.
* clocks sampled during the timed region: {b['clocks']}
* whole step: {sr['whole_step_tflops']} TFLOP/s of GEMM work = {100 * sr['whole_step_frac_of_tensor_peak']:.1f} % of the live-measured cuBLAS TF32 peak ({sr['tf32_peak_tflops']} TFLOP/s);
  algorithmic HBM bytes of all kernels / sum of kernel times = {sr['whole_step_hbm_gbs']} GB/s = {100 * sr['whole_step_hbm_gbs'] / sr['hbm_peak_gbs']:.0f} % of the measured HBM peak
  ({sr['hbm_peak_gbs']} GB/s, MEASURED_PEAKS.json).  The forward runs as ONE fused launch (hidden layers chained through
  tensor memory, heads in the epilogue, CTA pairs); the backward and the optimizer still move every activation
  gradient and the whole optimizer state through HBM once per step, which is what bounds the step (DESIGN.md section 7).

## Per-kernel roofline: CUDA events inside bench.py (`kernels`), algorithmic bytes/flops per launch; traffic = ncu DRAM bytes

| kernel | us | bound | GB/s | TFLOP/s | frac of peak | ncu dram MB |
|---|---|---|---|---|---|---|
{rows}

Dominant kernel: `{b['roofline']['kernel']}` ({b['roofline']['kernel_us']} us, {100 * b['roofline']['kernel_share_of_step']:.0f} % of the step):
{b['roofline']['algorithmic_bytes_per_launch'] / 1e6:.0f} MB algorithmic per launch at {b['roofline']['achieved']} GB/s = {100 * b['roofline']['frac']:.0f} % of measured HBM peak.

## Other BASELINE configs (same bench, `--workload`; steps/s summed over members; 1 GPU)

| workload | members | batch | hidden | steps/s | e2e steps/s | whole-step % of TF32 peak | algorithmic GB/s |
|---|---|---|---|---|---|---|---|
{chr(10).join(others)}

{dropin}
The stress shape (4x1024, batch 4096) is the compute-bound regime of the per-layer kernels (hidden width 1024 does not
fit the fused forward): `tools/umma_rate.py` measures the tcgen05 GEMM building block at 575-587 TFLOP/s on one CTA per
tile and 644-655 TFLOP/s on CTA pairs for a 8192 x 4096 x 4096 problem, all three operand layouts, against 788 TFLOP/s
for cuBLAS TF32 on the same box.

{extra}## ncu launch list (`{tag}_launches.csv`: `--metrics gpu__time_duration.sum --clock-control none`, graphs off, cold cache, serialised)

```
{ls}```

## ncu --set full, one launch per kernel (`{tag}_ncu_full_summary.md`)

{ncu}
Reading (see also DESIGN.md section 5 and `{tag}_cta_pair.md`): the backward GEMM launches saturate no single unit
(DRAM 35-55 %, L2 15-35 %, tensor pipe 10-27 %); switching parts of the kernel off (`tools/umma_probe.sh`) shows they
are bound by the HBM traffic of the phase, not by issue efficiency.  The fused forward is bound by a dependency
cycle through tensor memory (`{tag}_fused_trace.txt`: clock64 timeline of one CTA pair; DESIGN.md section 3): 4.4 us of MMAs
per tile inside a ~7.5 us period.  `adam_polyak` runs at 77 % DRAM utilisation under ncu (cold, serialised) and at
99 % of the measured copy peak inside the step.
"""
    open(os.path.join(P, f"{tag}_summary.md"), "w").write(md)
    print(md[:1500])


if __name__ == "__main__":
    main(*sys.argv[1:2])
