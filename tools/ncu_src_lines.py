"""Stall samples of a kernel by CUDA source line (needs -lineinfo and a
report captured with --import-source on).
Usage: python tools/ncu_src_lines.py rep.ncu-rep kernel_regex file_substring [top]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(rep, kernel, fname, top=30):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv',
                          '--print-source', 'cuda,sass', '--kernel-name',
                          'regex:' + kernel], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr = None, None
    agg = defaultdict(lambda: [0, 0, ''])
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1]
        elif r[0] == 'Line No':
            hdr = r
            si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
        elif cur and fname in cur and r[0].isdigit() and hdr:
            a = agg[int(r[0])]
            a[0] += int(r[si] or 0)
            a[1] += int(r[ii] or 0)
            a[2] = r[1]
    tot = sum(a[0] for a in agg.values()) or 1
    ins = sum(a[1] for a in agg.values()) or 1
    print('samples in {}: {}'.format(fname, tot))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('{:5d} {:5.1f}% samples {:5.1f}% instr  {}'.format(
            ln, 100 * a[0] / tot, 100 * a[1] / ins, a[2].strip()[:90]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3],
         int(sys.argv[4]) if len(sys.argv) > 4 else 30)
