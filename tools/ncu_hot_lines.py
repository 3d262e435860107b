"""Top stall sites of a kernel from an .ncu-rep captured with
--import-source on (SASS level, grouped by opcode and by neighbourhood).
Usage: python tools/ncu_hot_lines.py rep.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys


def main(rep, kernel, top=40):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv',
                          '--kernel-name', 'regex:' + kernel],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
    stall_cols = [c for c in rows[0] if c.startswith('stall_') and
                  'Not Issued' not in c]
    total = sum(int(r['# Samples'] or 0) for r in rows)
    inst = sum(int(r['Instructions Executed'] or 0) for r in rows)
    print('samples', total, 'warp instructions', inst)
    agg = {}
    for c in stall_cols:
        agg[c] = sum(int(r[c] or 0) for r in rows)
    print('by reason:', ', '.join('{} {:.1f}%'.format(k[6:], 100 * v / total)
                                  for k, v in sorted(agg.items(),
                                                     key=lambda kv: -kv[1])
                                  if v > 0.01 * total))
    byop = {}
    for r in rows:
        op = r['Source'].split()[0] if r['Source'].split() else '?'
        if op.startswith('@'):
            op = r['Source'].split()[1]
        op = op.split('.')[0]
        a = byop.setdefault(op, [0, 0])
        a[0] += int(r['# Samples'] or 0)
        a[1] += int(r['Instructions Executed'] or 0)
    print('by opcode (samples%, instr%):')
    for op, (s, n) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:18]:
        print('  {:12s} {:5.1f}% {:5.1f}%'.format(op, 100 * s / total,
                                                  100 * n / inst))
    print('top instructions:')
    idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]['# Samples'] or 0))
    for i in idx[:top]:
        r = rows[i]
        why = max(stall_cols, key=lambda c: int(r[c] or 0))
        print('  {:5d} {:5.1f}% {:10s} x{:<9s} {}'.format(
            i, 100 * int(r['# Samples']) / total, why[6:],
            r['Instructions Executed'], r['Source'][:90]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
