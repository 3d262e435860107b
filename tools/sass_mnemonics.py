"""Per-kernel counts of the SASS mnemonics that show which hardware paths the
library uses (tcgen05: UTCHMMA / UTCBAR / LDTM / STTM; TMA bulk copies: UBLKCP;
fp64 tensor cores: DMMA; clusters: UCGABAR; cp.async: LDGSTS; mbarriers:
SYNCS).  Usage: python tools/sass_mnemonics.py [library.so]"""
import collections
import re
import subprocess
import sys

KEYS = ('UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG',
        'SYNCS', 'UCGABAR_ARV', 'LDGSTS', 'DMMA', 'HMMA', 'MAPA', 'REDUX')


def main(path):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True,
                         text=True).stdout
    counts = collections.OrderedDict()
    fn = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], capture_output=True,
                                text=True).stdout.strip().split('(')[0]
            counts.setdefault(fn, collections.Counter())
            continue
        m = re.search(r'\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and fn and m.group(1) in KEYS:
            counts[fn][m.group(1)] += 1
    for fn, c in counts.items():
        if c:
            print('{:60s} {}'.format(fn[:60], ' '.join(
                '{}={}'.format(k, c[k]) for k in KEYS if c[k])))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1
         else 'nautilus_b200/libnautilus_b200.so')
