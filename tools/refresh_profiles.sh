#!/bin/bash
# turns what tools/gpu_session_final.sh brought back in gpurun_out/ into the tracked artefacts under profiles/
#   bash tools/refresh_profiles.sh <tag>        (tag = the prefix gpu_session_final.sh used, e.g. r1g)
set -e
T=${1:-r1g}
python tools/summarize_profiles.py $T gpurun_out/${T}_launches.csv gpurun_out/${T}_mrf_cluster.ncu-rep gpurun_out/${T}_mrf_branch.ncu-rep gpurun_out/${T}_conv_tc.ncu-rep gpurun_out/${T}_enc_res_stack.ncu-rep | tail -1
for p in x3:bf16x3 bf16:bf16 f32:f32 ref:reference_cpu; do cp gpurun_out/bench_${p%%:*}.json profiles/bench_${T}_${p##*:}.json; done
cp gpurun_out/ops_x3.log profiles/${T}_ops_bf16x3.txt
python - "$T" <<'PY'
import json, re, sys
T = sys.argv[1]
def grab(path):
    out = []
    for s in open(path).read().split('\n## ')[1:]:
        g = lambda k: float(re.search(r'\| ' + re.escape(k) + r' \| ([0-9.]+)', s).group(1))
        out.append((g('dram__bytes_read.sum'), g('dram__bytes_write.sum'), g('gpu__time_duration.sum'),
                    g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')))
    return out
rows = grab(f'profiles/{T}_{T}_mrf_cluster.md') + grab(f'profiles/{T}_{T}_mrf_branch.md')
rd, wr = [r[0] for r in rows], [r[1] for r in rows]
tot = sum(rd) + sum(wr)
d = json.load(open('profiles/r1_mrf_traffic.json'))
d.update({"dram_bytes_per_launch": tot * 1e6 / len(rows), "per_hop_MB": tot, "launches": len(rows), "dram_read_MB": rd, "dram_write_MB": wr})
json.dump(d, open('profiles/r1_mrf_traffic.json', 'w'), indent=1)
print("MRF launches (us, tensor %):", [(r[2], round(r[3], 1)) for r in rows], "DRAM MB/hop", round(tot, 1))
for name in ("bf16x3", "bf16", "f32", "reference_cpu"):
    l = [x for x in open(f'profiles/bench_{T}_{name}.json') if x.startswith('{')][-1]
    b = json.loads(l)
    print(name, round(b['value']), round(b['ms_per_step'], 4), b.get('e2e', {}).get('value'), b.get('e2e', {}).get('ms_per_step'),
          (b.get('roofline') or {}).get('frac'), (b.get('roofline') or {}).get('achieved'), b.get('rms_vs_cpu_oracle'))
PY
