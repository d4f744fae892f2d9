#!/usr/bin/env python
"""profiles/r02_launches.csv + profiles/r02_roofline_traffic.json from an ncu launch list with DRAM counters.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches_raw.csv python tools/prof_infer.py 2
  python tools/roofline_traffic.py gpurun_out/launches_raw.csv 275

Keeps the LAST forward (the last N repo-kernel launches), prints the per-kernel table, and writes the per-launch DRAM bytes
bench.py quotes in its `roofline.traffic` field.
"""
import collections
import csv
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize_ncu import short  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(src: str, per_forward: int):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10][1:]
    byid = collections.OrderedDict()
    for r in rows:
        d = byid.setdefault(r[0], {"kernel": short(r[4]), "grid": r[8], "block": r[7]})
        d[r[12]] = float(r[14].replace(",", ""))
    ours = [(i, d) for i, d in byid.items() if not d["kernel"].startswith("at::")]
    last = ours[-per_forward:]
    dst = os.path.join(ROOT, "profiles", "r02_launches.csv")
    fam = collections.OrderedDict()
    kinds = {"conv_gemm_bf16": [0, 0.0, 0.0], "conv_gemm_tf32": [0, 0.0, 0.0], "fused_mlp": [0, 0.0, 0.0]}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "duration_us", "dram_read_MB", "dram_write_MB"])
        for i, d in last:
            us = d.get("gpu__time_duration.sum", 0.0) / 1e3
            rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
            w.writerow([i, d["kernel"], d["grid"], d["block"], f"{us:.2f}", f"{rd / 1e6:.3f}", f"{wr / 1e6:.3f}"])
            k = re.sub(r"<.*", "", d["kernel"])
            a = fam.setdefault(k, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += us; a[2] += rd / 1e9; a[3] += wr / 1e9
            kk = None
            if k == "conv_gemm_kernel":      # first template argument: operand kind, MSPI_BF16 = 0, MSPI_F32 (tf32) = 1
                kk = "conv_gemm_bf16" if d["kernel"].startswith("conv_gemm_kernel<0") else "conv_gemm_tf32"
            elif k == "fused_mlp_kernel":
                kk = "fused_mlp"
            if kk:
                kinds[kk][0] += 1; kinds[kk][1] += rd + wr; kinds[kk][2] += us / 1e3
    tot = sum(a[1] for a in fam.values())
    print("| kernel | launches | total ms | share | dram read GB | dram write GB |\n|---|---|---|---|---|---|")
    for k, a in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {a[0]} | {a[1] / 1e3:.3f} | {a[1] / tot * 100:.1f}% | {a[2]:.3f} | {a[3]:.3f} |")
    print(f"| **total** | {sum(a[0] for a in fam.values())} | {tot / 1e3:.3f} | 100% | {sum(a[2] for a in fam.values()):.3f} | "
          f"{sum(a[3] for a in fam.values()):.3f} |")
    out = {"source": "profiles/r02_launches.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, "
                     "one eager B=32 forward; tools/roofline_traffic.py)"}
    for kk, (n, b, ms) in kinds.items():
        out[kk] = {"launches_per_step": n, "dram_bytes_per_step": b, "dram_bytes_per_launch": b / max(n, 1), "ms_per_step_ncu": ms}
    with open(os.path.join(ROOT, "profiles", "r02_roofline_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 275)
