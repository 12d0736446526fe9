"""SASS opcode histogram of the shipped library (no GPU needed): which kernels carry tcgen05 / TMA / TMEM / reduction opcodes.

    python tools/sass_histogram.py > profiles/rNN_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'REDG', 'RED', 'ATOMG', 'ATOMS', 'HMMA', 'LDGSTS', 'UTCCP']


def main():
    lib = os.path.join(ROOT, 'evreal_b200', 'libevreal_b200.so')
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    cur, per = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            per[cur][m.group(1).split('.')[0]] += 1
    print('# SASS opcode histogram of evreal_b200/libevreal_b200.so\n')
    print('`cuobjdump -sass evreal_b200/libevreal_b200.so` through tools/sass_histogram.py (no GPU needed).  `UTCHMMA` = tcgen05.mma')
    print('(kind::f16, bf16 operands), `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor loads, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier')
    print('operations, `REDG` = red.global (voxelizer, metric sums), `ATOMS` = shared-memory atomics (radix select, histograms);')
    print('no `HMMA` (legacy mma.sync) anywhere.  Kernels without any of these opcodes are omitted from the rows (counted in the total).\n')
    print('| kernel | instructions | ' + ' | '.join(KEYS) + ' |')
    print('|---|---|' + '---|' * len(KEYS))
    tot = collections.Counter()
    for f, c in per.items():
        tot.update(c)
        if not any(c[k] for k in KEYS):
            continue
        name = subprocess.run(['c++filt', f], capture_output=True, text=True).stdout.strip().split('(')[0][:70]
        print('| `%s` | %d | ' % (name, sum(c.values())) + ' | '.join(str(c[k]) if c[k] else '' for k in KEYS) + ' |')
    print('| **all %d kernels** | %d | ' % (len(per), sum(tot.values())) + ' | '.join(str(tot[k]) if tot[k] else '' for k in KEYS) + ' |')


if __name__ == '__main__':
    main()
