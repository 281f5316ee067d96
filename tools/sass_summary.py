"""SASS opcode summary of the compiled kernels (no GPU needed): python tools/sass_summary.py > profiles/rNN_sass_summary.txt
Counts, per kernel of the listed objects, the Blackwell-specific mnemonics (B200_PROFILING.md: UTMALDG / UTMASTG / UBLKCP = TMA,
UTMAREDG = TMA reduce-store, SYNCS = mbarrier, REDG = L2 reduction), the packed FP32 instructions and the code size."""
import collections
import os
import re
import subprocess
import sys

OBJ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'beyond_dof_b200', 'csrc', '_obj')
KEYS = ['UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKCP', 'UBLKPF', 'SYNCS', 'REDG', 'FADD2', 'FMUL2', 'FFMA2', 'FFMA', 'LDS', 'STS', 'LDG', 'STG',
        'BAR', 'SHFL', 'MUFU']


def demangle(names):
    out = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r'\(bdof::\w+Params.*', '', o).replace('bdof::', '').replace('(int)', '').replace('(bool)', '') for o in out]


def main(objs):
    for o in objs:
        path = os.path.join(OBJ, o)
        sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
        funcs, cur = collections.OrderedDict(), None
        for line in sass.splitlines():
            m = re.match(r'\s*Function : (\S+)', line)
            if m:
                cur = m.group(1)
                funcs[cur] = collections.Counter()
                continue
            m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
            if m and cur:
                funcs[cur][m.group(2)] += 1
                funcs[cur]['__last'] = int(m.group(1), 16)
        names = demangle(list(funcs))
        print('== %s' % o)
        for (mangled, c), name in zip(funcs.items(), names):
            if not any(k in name for k in ('sweep_kernel', 'resident_', 'pipe_col', 'k_halo_push', 'line_kernel')):
                continue
            if 'line_kernel' in name and o not in ('line_2048.o',):
                continue
            total = sum(v for k, v in c.items() if k != '__last')
            print('  %s' % name[:150])
            print('      %d instructions, %.1f KB; ' % (total, (c['__last'] + 16) / 1024.0) +
                  ', '.join('%s %d' % (k, c[k]) for k in KEYS if c[k]))
        print()


if __name__ == '__main__':
    main(sys.argv[1:] or ['line_2048.o', 'line_4096.o', 'line_256.o', 'resident_inst.o', 'tilehalo.o'])
