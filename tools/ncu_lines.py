#!/usr/bin/env python
"""Attribute an ncu SASS-level profile to CUDA source lines.

ncu's CLI prints per-instruction counters (`--page source --csv`) but not per source line; this joins them with the
line table of the same binary (`nvdisasm -gi`, needs -lineinfo) by instruction offset and prints the hottest lines
(innermost inlined location, plus the line of the kernel body they were inlined into).

  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [--id N] [--top 40] [--so path/to/lib.so]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile


def line_table(so, mangled_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '-gi', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table, chain, fresh, on = {}, [], True, False
    for ln in txt.splitlines():
        if ln.startswith('.text.'):
            on = mangled_substr in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:   # a chain of lines, innermost inlined location first, the kernel body's line last
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', ln)
        if m:
            fresh = True
            table[int(m.group(1), 16)] = (tuple(chain), m.group(2).strip())
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('kernel', help='substring of the mangled kernel name in the cubin, e.g. k_penetrationILi416')
    ap.add_argument('--ncu-kernel', default=None, help='ncu --kernel-id filter, e.g. ::regex:k_penetration:1')
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--so', default='jogramop_framework_b200/libjogramop_b200.so')
    a = ap.parse_args()
    cmd = ['ncu', '-i', a.report, '--page', 'source', '--csv']
    if a.ncu_kernel:
        cmd += ['--kernel-id', a.ncu_kernel]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    h = rows[hdr_i]
    ci = {n: h.index(n) for n in ('Address', 'Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed')}
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ('Kernel Name', 'Address'):
            break
        body.append(r)
    base = int(body[0][ci['Address']], 0)
    tab = line_table(a.so, a.kernel)
    per = collections.defaultdict(lambda: [0, 0, 0])
    outer = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in body:
        off = int(r[ci['Address']], 0) - base
        loc, _ = tab.get(off, (None, ''))
        v = [int(r[ci['# Samples']] or 0), int(r[ci['Instructions Executed']] or 0), int(r[ci['Thread Instructions Executed']] or 0)]
        key = loc[0] if loc else ('?', 0)
        for i in range(3):
            per[key][i] += v[i]
            tot[i] += v[i]
            for k in set(loc or [('?', 0)]):
                outer[k][i] += v[i]
    print(f'total samples {tot[0]}  warp-inst {tot[1]}  thread-inst {tot[2]}  lanes/inst {tot[2] / max(tot[1], 1):.1f}')
    print('--- innermost source line: samples%  inst%  lanes/inst')
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f'{k[0]}:{k[1]:<5d} {100 * v[0] / tot[0]:6.2f}% {100 * v[1] / max(tot[1], 1):6.2f}%  {v[2] / max(v[1], 1):5.1f}')
    print('--- inclusive (the line itself or anything inlined beneath it)')
    for k, v in sorted(outer.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f'{k[0]}:{k[1]:<5d} {100 * v[0] / tot[0]:6.2f}% {100 * v[1] / max(tot[1], 1):6.2f}%  {v[2] / max(v[1], 1):5.1f}')


if __name__ == '__main__':
    main()
