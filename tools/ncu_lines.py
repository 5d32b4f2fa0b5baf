#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel in an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [min-percent]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else ''
    minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fn, hdr, cur_file = None, None, None
    lines = []
    seen = set()
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
        elif r[0] == 'Function Name':
            fn = r[1]
        elif r[0] == 'Line No':
            hdr = r
        elif hdr and fn and kern in fn and r[0].isdigit():
            if fn not in seen:
                seen.add(fn)
            ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
            ith = hdr.index('Thread Instructions Executed')
            try:
                lines.append((fn, cur_file, int(r[0]), float(r[ie] or 0), float(r[isamp] or 0), float(r[ith] or 0), r[1].strip()))
            except (ValueError, IndexError):
                pass   # a source line with embedded quotes confused the CSV split
    # the report repeats functions per launch; keep the first launch of the kernel
    first = lines[0][0] if lines else None
    lines = [l for l in lines if l[0] == first]
    agg, order = {}, []
    for l in lines:            # the same kernel may have been captured several times: aggregate per (file, line)
        k = (l[1], l[2])
        if k not in agg:
            agg[k] = [l[0], l[1], l[2], 0.0, 0.0, 0.0, l[6]]
            order.append(k)
        agg[k][3] += l[3]; agg[k][4] += l[4]; agg[k][5] += l[5]
    lines = [tuple(agg[k]) for k in order]
    tot = sum(l[3] for l in lines) or 1
    tots = sum(l[4] for l in lines) or 1
    print('kernel:', first)
    print('warp instructions: %.0f   samples: %.0f' % (tot, tots))
    for l in lines:
        if 100 * l[3] / tot >= minpct or 100 * l[4] / tots >= minpct:
            print('%-26s %4d  inst %5.1f%%  samp %5.1f%%  thr/inst %4.1f | %s' % (l[1], l[2], 100 * l[3] / tot, 100 * l[4] / tots,
                                                                                 l[5] / max(l[3], 1), l[6][:100]))


if __name__ == '__main__':
    main()
