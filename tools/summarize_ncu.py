#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/) into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.csv
        per-launch list (kernel, grid, duration, dram bytes) + a per-kernel-family table on stdout / .md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_prof_x.csv
        selected raw metrics of every captured launch of an `ncu --set full` report
"""
import collections
import csv
import re
import subprocess
import sys


def short(name: str) -> str:
    name = name.replace("void ", "").replace("mspi::<unnamed>::", "").replace("mspi::(anonymous namespace)::", "")
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("mspi::", "")
    m = re.match(r"([A-Za-z0-9_]+)(<[^(]*?>)?\(", name)
    if not m:
        return name[:60]
    t = m.group(2) or ""
    t = t.replace("(int)", "").replace("__nv_bfloat16", "bf16").replace(" ", "")
    return m.group(1) + t


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    data = rows[1:]
    byid = collections.OrderedDict()
    for r in data:
        d = byid.setdefault(r[0], {"kernel": short(r[4]), "grid": r[8], "block": r[7]})
        d[r[12]] = float(r[14].replace(",", ""))
    fam = collections.OrderedDict()
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "duration_us", "dram_read_MB", "dram_write_MB"])
        for i, d in byid.items():
            us = d.get("gpu__time_duration.sum", 0.0) / 1e3
            rd, wr = d.get("dram__bytes_read.sum", 0.0) / 1e6, d.get("dram__bytes_write.sum", 0.0) / 1e6
            w.writerow([i, d["kernel"], d["grid"], d["block"], f"{us:.2f}", f"{rd:.3f}", f"{wr:.3f}"])
            k = re.sub(r"<.*", "", d["kernel"])
            a = fam.setdefault(k, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += us; a[2] += rd; a[3] += wr
    tot = sum(a[1] for a in fam.values())
    lines = ["| kernel | launches | total ms | share | dram read GB | dram write GB |", "|---|---|---|---|---|---|"]
    for k, a in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {a[0]} | {a[1] / 1e3:.3f} | {a[1] / tot * 100:.1f}% | {a[2] / 1e3:.3f} | {a[3] / 1e3:.3f} |")
    lines.append(f"| **total** | {sum(a[0] for a in fam.values())} | {tot / 1e3:.3f} | 100% | "
                 f"{sum(a[2] for a in fam.values()) / 1e3:.3f} | {sum(a[3] for a in fam.values()) / 1e3:.3f} |")
    md = "\n".join(lines)
    print(md)
    with open(dst.replace(".csv", ".md"), "w") as f:
        f.write(f"ncu launch list summary (`{src}`; per-launch times are cold-cache and serialised: compare shares)\n\n{md}\n")


KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_bytes.sum", "lts__t_bytes.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or any(h.startswith(k) for k in ("sm__pipe_tensor", "sm__inst_executed_pipe_tensor"))]
    name_i = hdr.index("Kernel Name")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [hdr[i] for i in cols])
        w.writerow(["(unit)"] + [rows[1][i] for i in cols])
        for r in rows[2:]:
            w.writerow([short(r[name_i])] + [r[i] for i in cols])
    print(open(dst).read()[:3000])


def traffic(src, dst):
    """profiles/*_launches.csv -> the roofline traffic json bench.py reads (dram bytes per launch of the tensor-core kernels)."""
    import json
    rows = list(csv.DictReader(open(src)))
    out = {"source": f"{src} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, one eager B=32 forward)"}
    for key, pred in (("conv_gemm_bf16", lambda k: k.startswith("conv_gemm_kernel<0") or k.startswith("conv_gemm_kernel<bf16")),
                      ("conv_gemm_tf32", lambda k: k.startswith("conv_gemm_kernel<1")),
                      ("fused_mlp", lambda k: k.startswith("fused_mlp"))):
        sel = [r for r in rows if pred(r["kernel"])]
        b = sum((float(r["dram_read_MB"]) + float(r["dram_write_MB"])) * 1e6 for r in sel)
        out[key] = {"launches_per_step": len(sel), "dram_bytes_per_step": b, "dram_bytes_per_launch": b / max(len(sel), 1),
                    "ms_per_step_ncu": sum(float(r["duration_us"]) for r in sel) / 1e3}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
