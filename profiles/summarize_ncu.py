#!/usr/bin/env python
"""Turn an `ncu --set full` report into the short text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_<name>.txt [profiles/r01_traffic.json]

With a third argument, the per-launch DRAM traffic of every kernel in the report (mean over
its instances, keyed by kernel name + grid size) is merged into that JSON file; bench.py reads
`roofline.traffic` from it.

Reads the report on the CPU box (`ncu -i ... --page raw/source --csv`); numbers taken under
the profiler are evidence about the kernel's behaviour, never bench values."""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", "--print-units", "base"], capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = [f"ncu summary of {rep} (ncu --set full --clock-control none --import-source on)", ""]
    raw = page(rep, "raw")
    hdr, units, rows = raw[0], raw[1], raw[2:]
    for r in rows:
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:72s} {r[i]:>16s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                stalls.append((num(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        lines.append("  warp stall reasons (warps per issue-active cycle): " +
                     ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
        wr, rd = num(r[hdr.index("dram__bytes_write.sum")]), num(r[hdr.index("dram__bytes_read.sum")])
        lines.append(f"  dram traffic per launch (read+write, units as above): {rd + wr:.1f}")
        lines.append("")
    src = page(rep, "source")
    if len(src) > 2:
        h2 = src[1]
        data = [r for r in src[2:] if len(r) > 5]
        i_s, i_i, i_src = h2.index("Warp Stall Sampling (All Samples)"), h2.index("Instructions Executed"), h2.index("Source")
        tot = sum(num(r[i_s]) for r in data) or 1.0
        lines.append(f"top stall-sampled SASS (first kernel instance; {len(data)} instructions, {int(tot)} samples)")
        for r in sorted(data, key=lambda r: -num(r[i_s]))[:16]:
            lines.append(f"  {100 * num(r[i_s]) / tot:5.1f}%  exec={r[i_i]:>9s}  {r[i_src].strip()[:90]}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))
    if len(sys.argv) > 3:
        import json
        import os
        acc = {}
        for r in rows:
            key = f"{r[hdr.index('Kernel Name')]} grid={r[hdr.index('launch__grid_size')]}"
            a = acc.setdefault(key, {"n": 0, "rd": 0.0, "wr": 0.0, "ns": 0.0})
            a["n"] += 1
            a["rd"] += num(r[hdr.index("dram__bytes_read.sum")])
            a["wr"] += num(r[hdr.index("dram__bytes_write.sum")])
            a["ns"] += num(r[hdr.index("gpu__time_duration.sum")])
        path = sys.argv[3]
        doc = json.load(open(path)) if os.path.exists(path) else {}
        for key, a in acc.items():
            doc[key] = {"dram_bytes_read": a["rd"] / a["n"], "dram_bytes_write": a["wr"] / a["n"],
                        "traffic_bytes": (a["rd"] + a["wr"]) / a["n"], "gpu_time_ns_under_ncu": a["ns"] / a["n"],
                        "instances": a["n"], "report": os.path.basename(rep)}
        if len(sys.argv) > 5 and sys.argv[4] == "--alias":   # e.g. --alias bench_rollout: bench.py reads this kernel's traffic
            first = f"{rows[0][hdr.index('Kernel Name')]} grid={rows[0][hdr.index('launch__grid_size')]}"
            doc.setdefault("_alias", {})[sys.argv[5]] = first
        json.dump(doc, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
