#!/usr/bin/env python
"""Rank the CUDA source lines of a kernel in an .ncu-rep by executed warp instructions (needs -lineinfo and
--import-source on).  python tools/ncu_lines.py file.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
out = []; tot = 0; fname = ''
for r in rows:
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]
    if len(r) < 8 or r[0] in ('', 'Line No') or r[2] != '-': continue
    try: n = int(r[7]); st = int(r[4])
    except ValueError: continue
    out.append((n, st, fname, r[0], r[1][:105])); tot += n
print('executed warp instructions:', tot)
for n, st, f, l, s in sorted(out, reverse=True)[:top]:
    print(f'{100 * n / tot:5.1f}% stall {st:6d}  {f}:{l}: {s}')
