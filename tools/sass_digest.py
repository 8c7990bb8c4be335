"""Per-kernel SASS mnemonic counts of libmpe_b200.so -> profiles/r2_sass_digest.txt (no GPU needed):

    python tools/sass_digest.py > profiles/r2_sass_digest.txt

What proves a Blackwell-native kernel (B200 profiling guide): tcgen05.mma shows up as UTC*MMA, tcgen05.ld/st as
LDTM/STTM, TMA bulk copies as UBLKCP / UTMA*, tcgen05.commit / mbarrier traffic as UTCBAR / SYNCS; MUFU.* are the
transcendental (XU pipe) instructions and FFMA2 / FADD2 / FMUL2 the packed fp32x2 forms."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'multiagent_rl_b200', 'libmpe_b200.so')
WATCH = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'MUFU.EX2', 'MUFU.LG2', 'MUFU.RCP',
         'MUFU.RSQ', 'MUFU.SQRT', 'MUFU.TANH', 'FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'HMMA', 'LDG', 'STG', 'LDS', 'STS',
         'SHFL', 'BAR', 'ATOM', 'RED', 'DFMA', 'DADD', 'DMUL']


def demangle(names):
    out = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts, total, order = {}, {}, []
    cur = None
    for line in sass.split('\n'):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            total[cur] = 0
            order.append(cur)
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur is not None:
            op = m.group(1)
            total[cur] += 1
            for w in WATCH:
                if op == w or op.startswith(w + '.') or (w.startswith('MUFU') and op.startswith(w)):
                    counts[cur][w] += 1
                    break
    names = demangle(order)
    print('SASS digest of %s (cuobjdump -sass, sm_100a); columns = static instruction counts per kernel' % os.path.basename(LIB))
    print('kernels: %d' % len(order))
    agg = collections.Counter()
    for k in order:
        agg.update(counts[k])
    print('whole library: ' + ', '.join('%s %d' % (w, agg[w]) for w in WATCH if agg[w]))
    print()
    for k in sorted(order, key=lambda k: -total[k]):
        c = counts[k]
        name = re.sub(r'\(.*', '', names.get(k, k))
        name = name.replace('void mpe::', '').replace('mpe::', '')
        print('%-64s %6d instr | %s' % (name[:64], total[k], ', '.join('%s %d' % (w, c[w]) for w in WATCH if c[w])))


if __name__ == '__main__':
    main()
