"""Where does a kernel wait?  Reads the SASS-level warp-state samples of an `ncu --set full` report and prints
(1) samples per opcode, (2) samples per spin-wait site grouped by the mbarrier operand, (3) the hottest instructions.

    python tools/dev/ncu_stalls.py gpurun_out/prof.ncu-rep conv3d_tc_kernel [launch_index]

This is how round 1 found that the conv kernel's single MMA-issuing thread (not TMA, not the tensor pipe) was the
limiter, and that the gather kernels wait on their loads (long scoreboard) rather than on issue slots."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def sections(rep, kernel):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name',
                          f'regex:{kernel}'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    idx = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    for a, b in zip(idx, idx[1:]):
        hdr = rows[a + 1]
        data = [r for r in rows[a + 2:b] if len(r) > 6 and r[hdr.index('# Samples')].isdigit()]
        yield rows[a][1], hdr, data


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
    secs = list(sections(rep, kernel))
    if not secs:
        sys.exit('no such kernel in the report')
    name, h, data = secs[which]
    ns, ne, src = h.index('# Samples'), h.index('Instructions Executed'), h.index('Source')
    seen, uniq = set(), []
    for r in data:                                              # the CSV repeats every row once per view
        if r[0] not in seen:
            seen.add(r[0]); uniq.append(r)
    total = sum(int(r[ns]) for r in uniq)
    print(name[:110]); print('samples', total, 'instructions', len(uniq))
    by_op, ex = Counter(), Counter()
    for r in uniq:
        t = r[src].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        by_op[op] += int(r[ns]); ex[op] += int(r[ne])
    print('\n-- samples per opcode')
    for op, v in by_op.most_common(12):
        print(f'  {op:<14}{v:>8}{100 * v / max(total, 1):>7.1f} %   executed {ex[op]}')
    spin = Counter()
    for i, r in enumerate(uniq):
        if 'TRYWAIT' in r[src]:
            m = re.search(r'\[(.*?)\]', r[src])
            s = int(r[ns])
            for j in range(i + 1, min(i + 4, len(uniq))):
                if 'BRA' in uniq[j][src]:
                    s += int(uniq[j][ns]); break
            spin[m.group(1)] += s
    if spin:
        print('\n-- samples in mbarrier spin waits, by barrier operand')
        for k, v in spin.most_common(10):
            print(f'  {k:<28}{v:>8}{100 * v / max(total, 1):>7.1f} %')
    stall_cols = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    print('\n-- hottest instructions')
    for r in sorted(uniq, key=lambda r: -int(r[ns]))[:15]:
        st = {c[6:]: int(r[h.index(c)]) for c in stall_cols if r[h.index(c)].isdigit() and int(r[h.index(c)]) > 0}
        print(f'  {r[ns]:>6} {r[ne]:>9}  {r[src][:64]:<64} {st}')


if __name__ == '__main__':
    main()
