"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the small, tracked
summaries under profiles/.

  python tools/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep> [more.ncu-rep ...]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

METRICS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "smsp__inst_executed.sum"]


def launches(tag, path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    seq = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("b200::<unnamed>::", "").replace("unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        seq.append((name, r[gi], v))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: launch list of `bench.py --steps 2 --warmup 3` under ncu\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n")
        f.write(f"{len(seq)} launches captured, {tot / 1e3:.1f} us total.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t / 1e3:.1f} | {t / tot:.3f} |\n")
        f.write("\n## per launch (in issue order)\n\n| # | kernel | grid | us |\n|---:|---|---|---:|\n")
        for i, (k, g, v) in enumerate(seq):
            f.write(f"| {i} | `{k}` | {g} | {v / 1e3:.1f} |\n")


def full(tag, rep):
    name = os.path.splitext(os.path.basename(rep))[0]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(m, hdr.index(m)) for m in METRICS if m in hdr]

    def norm(i, v):
        """ncu scales the unit per report (byte / Kbyte / Mbyte, ns / us ...): bring bytes to MB and times to us."""
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i])
        if scale is None:
            return v
        try:
            return f"{float(v.replace(',', '')) * scale:.6f}"
        except ValueError:
            return v

    stall = [(i, h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
             if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    with open(os.path.join(OUT, f"{tag}_{name}.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` -- {name}\n\n")
        f.write("Units normalised (time us, bytes MB, percentages).  `traffic` = dram read + write per launch.\n\n")
        for r in rows[2:]:
            f.write("## " + r[hdr.index("Kernel Name")][:110] + "\n\n| metric | value |\n|---|---|\n")
            for m, i in idx[1:]:
                f.write(f"| {m} | {norm(i, r[i])} |\n")
            st = []
            for i, h in stall:
                try:
                    st.append((float(r[i]), h))
                except ValueError:
                    pass
            tot = sum(v for v, _ in st) or 1.0
            f.write("| warp-stall samples (top) | " + ", ".join(f"{h} {100 * v / tot:.0f}%" for v, h in sorted(st, reverse=True)[:5]) + " |\n\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    for rep in sys.argv[3:]:
        full(tag, rep)
