#!/usr/bin/env python
"""SASS lines [a, b) of one kernel with executed counts and stall samples.  usage: ncu_sass_dump.py report kernel-regex a b"""
import csv, subprocess, sys, io
rep, kre, a, b = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(r for r in rows if 'Instructions Executed' in r)
ix, isamp, isrc = h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
body = rows[rows.index(h) + 1:]
for k, r in enumerate(body[a:b]):
    print(a + k, r[ix].rjust(9), r[isamp].rjust(5), r[isrc].strip()[:90])
