#!/usr/bin/env python
"""Per-region instruction / stall-sample shares of one kernel from an ncu report (source page, SASS view).
usage: ncu_sass_regions.py report.ncu-rep kernel-regex [bucket]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(r for r in rows if 'Instructions Executed' in r)
ix, isamp, isrc = h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
ins = []
for r in rows[rows.index(h) + 1:]:
    try:
        ins.append((r[isrc].strip(), int(r[ix]), int(r[isamp])))
    except Exception:
        pass
tot = sum(v for _, v, _ in ins); ts = sum(s for *_, s in ins)
print(len(ins), 'SASS instructions,', tot, 'warp-instructions executed,', ts, 'samples')
for a in range(0, len(ins), bucket):
    seg = ins[a:a + bucket]
    sv = sum(v for _, v, _ in seg); ss = sum(s for *_, s in seg)
    ops = collections.Counter((s.split()[1] if s.startswith('@') else s.split()[0]) for s, _, _ in seg)
    if sv > 0.002 * tot:
        print(f"{a:5d} inst {sv / tot * 100:5.1f}%  samples {ss / ts * 100:5.1f}%  max_exec {max(v for _, v, _ in seg):9d}  " +
              ' '.join(f'{k}:{v}' for k, v in ops.most_common(5)))
