"""Summarises '[mrf cta]' / '[mrfc cta]' lines (BEATRICE_B200_MRF_TRACE=-1) of the LAST hop in a log:
per kernel (C) the launch window, per-branch CTA durations and start skew."""
import collections
import re
import sys

rows = []
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r"\[(mrfc?) cta\] C (\d+) bx (\d+) by (\d+) sm (\d+) start_ns (\d+) end_ns (\d+)", line)
    if m:
        rows.append((m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)), int(m.group(6)), int(m.group(7))))
byC = collections.defaultdict(list)
for r in rows:
    byC[r[1]].append(r)
for C, rs in sorted(byC.items(), reverse=True):
    # keep the last launch: rows come grouped per launch; split on large time gaps
    rs.sort(key=lambda r: r[5])
    launches, cur = [], [rs[0]]
    for r in rs[1:]:
        if r[5] - cur[-1][5] > 200000:
            launches.append(cur)
            cur = []
        cur.append(r)
    launches.append(cur)
    L = launches[-1]
    t0 = min(r[5] for r in L)
    t1 = max(r[6] for r in L)
    print(f"C={C}: {len(L)} CTAs, window {(t1 - t0) / 1e3:.1f} us, {len(set(r[4] for r in L))} SMs")
    for by in sorted(set(r[3] for r in L)):
        B = [r for r in L if r[3] == by]
        st = [(r[5] - t0) / 1e3 for r in B]
        du = [(r[6] - r[5]) / 1e3 for r in B]
        en = [(r[6] - t0) / 1e3 for r in B]
        print(f"   by={by}: {len(B)} CTAs  start {min(st):.1f}..{max(st):.1f} us  duration {min(du):.1f}..{max(du):.1f} (mean {sum(du) / len(du):.1f})  end {min(en):.1f}..{max(en):.1f}")
