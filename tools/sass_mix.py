#!/usr/bin/env python
"""Static SASS instruction mix per kernel of an object / cubin / .so: python tools/sass_mix.py <file> [name-filter]"""
import collections
import re
import subprocess
import sys

KEYS = ['FADD', 'FADD2', 'FFMA', 'FFMA2', 'FMUL', 'FMUL2', 'MOV', 'LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL', 'LDGSTS',
        'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'IMAD', 'IADD3', 'LOP3', 'SHF', 'PRMT', 'BAR', 'MUFU', 'UTCHMMA', 'LDTM']


def main():
    out = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
    filt = sys.argv[2] if len(sys.argv) > 2 else ''
    fn, c = None, collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if m and fn:
            c[fn][m.group(1).split('.')[0]] += 1
    for fn, cc in c.items():
        if filt and filt not in fn:
            continue
        name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()
        print(f"{name[:90]}: total {sum(cc.values())} " + ' '.join(f"{k}={cc[k]}" for k in KEYS if cc[k]))


if __name__ == '__main__':
    main()
