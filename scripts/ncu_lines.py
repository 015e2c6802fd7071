#!/usr/bin/env python
"""Aggregate an ncu report's warp-stall samples by CUDA source line.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; lines = []
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and r and r[0] not in ('', 'Function Name') and len(r) > 7 and r[2] == '-':
        try: lines.append((cur, int(r[0]), r[1].strip(), int(r[4] or 0), int(r[7] or 0), r))
        except ValueError: pass
tot = sum(l[3] for l in lines)
idx = {n: i for i, n in enumerate(hdr)}
sc = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
print("total samples", tot)
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    r = l[5]
    st = sorted(((int(r[idx[c]] or 0), c) for c in sc), reverse=True)[:3]
    print("%s:%d %5.1f%% inst=%d  %s | %s" % (l[0][:14], l[1], 100 * l[3] / max(tot, 1), l[4], l[2][:64],
                                              ' '.join('%s=%d' % (c[6:], v) for v, c in st)))

if len(sys.argv) > 3:      # extra args: name:lo-hi line ranges of the main file to aggregate
    for spec in sys.argv[3:]:
        name, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))
        sel = [l for l in lines if lo <= l[1] <= hi and l[0].startswith('chorin_fd_slab')]
        tots = sum(l[3] for l in sel)
        agg = {}
        for l in sel:
            for c in sc:
                agg[c] = agg.get(c, 0) + int(l[5][idx[c]] or 0)
        top5 = sorted(agg.items(), key=lambda x: -x[1])[:6]
        print("%-14s %6.1f%% of samples, inst=%d | %s" % (name, 100 * tots / tot, sum(l[4] for l in sel),
              ' '.join('%s=%.1f%%' % (c[6:], 100 * v / max(tots, 1)) for c, v in top5)))
