"""Turns what `tools/profile_session.sh <tag>` brought back in gpurun_out/ into the tracked artefacts under profiles/:
bench lines, the ncu launch list, the ncu --set full summaries (+ MRF DRAM traffic), per-op times, MRF timelines,
config 4 / 5 results -- and the SASS opcode table of the built library (cuobjdump, runs here: no GPU needed).

   python tools/profile_collect.py <tag>          e.g. r2"""
import collections
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarize_profiles as sp  # noqa: E402

G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def sass_table(tag):
    so = os.path.join(ROOT, "beatrice_vst_b200", "csrc", "libbeatrice_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    ops = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "HMMA", "FFMA"]
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("b200::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for o in ops:
                if op == o or op.startswith(o + "."):
                    cur[o] += 1
            cur["_all"] += 1
    with open(os.path.join(P, f"{tag}_sass_opcodes.md"), "w") as f:
        f.write(f"# {tag}: SASS opcode counts per kernel of libbeatrice_b200.so (`cuobjdump -sass`, sm_100a)\n\n")
        f.write("`UTCHMMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit, `LDTM` / `STTM` = tcgen05.ld / st (TMEM), `UBLKCP` = cp.async.bulk (TMA "
                "bulk copies; the kernels use bulk copies of pre-laid-out operand tiles, no tensor maps, hence no `UTMALDG`), `SYNCS` = "
                "mbarrier traffic, `LDGSTS` = cp.async, `HMMA` = legacy mma.sync (none).\n\n")
        f.write("| kernel | instructions | " + " | ".join(ops) + " |\n|---|---:|" + "---:|" * len(ops) + "\n")
        tot = collections.Counter()
        for name, c in per.items():
            if c["_all"] == 0:
                continue
            f.write(f"| `{name[:90]}` | {c['_all']} | " + " | ".join(str(c[o]) for o in ops) + " |\n")
            tot.update(c)
        f.write(f"| **library total** | {tot['_all']} | " + " | ".join(str(tot[o]) for o in ops) + " |\n")
    return tot


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    os.makedirs(P, exist_ok=True)
    g = lambda n: os.path.join(G, f"{tag}_{n}")  # noqa: E731
    if os.path.exists(g("launches.csv")):
        sp.launches(tag, g("launches.csv"))
    traffic = {}
    for name in ("mrf_cluster", "mrf_branch", "mrf_branch_ups", "conv_tc", "enc_res_stack"):
        rep = g(name + ".ncu-rep")
        if os.path.exists(rep):
            sp.full(tag, rep)
            os.replace(os.path.join(P, f"{tag}_{tag}_{name}.md"), os.path.join(P, f"{tag}_ncu_{name}.md"))
    # DRAM traffic of the MRF launches of one hop (ncu --set full): read + write, per launch
    rows = []
    for name in ("mrf_cluster", "mrf_branch"):
        path = os.path.join(P, f"{tag}_ncu_{name}.md")
        if not os.path.exists(path):
            continue
        for s in open(path).read().split("\n## ")[1:]:
            v = lambda k: float(re.search(r"\| " + re.escape(k) + r" \| ([0-9.]+)", s).group(1))  # noqa: E731
            rows.append(dict(kernel=s.split("\n")[0][:60], read_MB=v("dram__bytes_read.sum"), write_MB=v("dram__bytes_write.sum"),
                             us=v("gpu__time_duration.sum"), tensor_pct=v("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")))
    if rows:
        tot = sum(r["read_MB"] + r["write_MB"] for r in rows)
        traffic = {"what": "DRAM read + write of the fused-MRF launches of ONE hop (256 streams, bf16x3), ncu --set full", "launches": rows,
                   "per_hop_MB": tot, "kernels": len(rows)}
        json.dump(traffic, open(os.path.join(P, f"{tag}_mrf_traffic.json"), "w"), indent=1)
    for n in ("bench_20.json", "bench_1000.json", "bench_reference.json", "bench_bf16.json", "bench_f32.json", "ops_bf16x3.txt",
              "ops_bf16x3_depth1.txt", "ablation.txt", "mrf_timeline.txt", "mrf_timeline_ups.txt", "config4_latency.json", "config5_1gpu.json", "config5_8gpu.json", "bench_2gpu.json",
              "bench_4gpu.json", "bench_8gpu.json"):
        if os.path.exists(g(n)) and os.path.getsize(g(n)) > 0:
            shutil.copy(g(n), os.path.join(P, f"{tag}_{n}"))
    tot = sass_table(tag)
    print("sass:", {k: v for k, v in tot.items() if k != "_all"})
    for n in ("bench_20.json", "bench_1000.json"):
        p = os.path.join(P, f"{tag}_{n}")
        if os.path.exists(p):
            lines = [x for x in open(p) if x.startswith("{")]
            if lines:
                b = json.loads(lines[-1])
                print(n, round(b["value"]), round(b["ms_per_step"], 4), round(b["e2e"]["value"]), b["roofline"]["frac"],
                      b.get("rms_vs_cpu_oracle"), b["latency_mode"]["value"])
    if traffic:
        print("MRF DRAM MB/hop", round(traffic["per_hop_MB"], 1), [(r["us"], round(r["tensor_pct"], 1)) for r in traffic["launches"]])


if __name__ == "__main__":
    main()
