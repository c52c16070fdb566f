"""Source-level stall attribution of one kernel of an ncu report, by warp role.

    ncu -i gpurun_out/r2d_kernels.ncu-rep --page source --csv > /tmp/src.csv
    python profiles/stall_roles.py /tmp/src.csv <kernel index in the csv> name:first_instr ...

Each `name:first_instr` starts a region of the SASS listing (the roles of a warp-specialised kernel are contiguous
branches).  Prints per region: share of the warp-stall samples, the split into issuing / stall reasons, the samples
spent in mbarrier polls (NANOSLEEP.SYNCS after a TRYWAIT: the role is waiting for another role), and the shared-memory
wavefronts / global requests per opcode.  Warp-stall samples are per warp: a role with w of the CTA's W warps owns
about w / W of them while it lives.
"""
import collections
import csv
import re
import sys


def load(path, which):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    s, e = starts[which], starts[which + 1]
    hdr = rows[s + 1]
    data = [r for r in rows[s + 2:e] if r and r[0].startswith('0x')]
    return rows[s][1], {h: i for i, h in enumerate(hdr)}, data


def main():
    path, which = sys.argv[1], int(sys.argv[2])
    regions = [(a.split(':')[0], int(a.split(':')[1])) for a in sys.argv[3:]]
    name, ix, data = load(path, which)
    stalls = [h for h in ix if h.startswith('stall_') and 'Not Issued' not in h]
    S, X = ix['# Samples'], ix['Instructions Executed']
    tot = sum(int(r[S]) for r in data)
    print(f'== {name}\n   {len(data)} SASS instructions, {tot} warp-stall samples')
    bounds = regions + [('end', len(data))]
    for (rn, lo), (_, hi) in zip(bounds, bounds[1:]):
        seg = data[lo:hi]
        n = sum(int(r[S]) for r in seg)
        ex = sum(int(r[X]) for r in seg)
        st = collections.Counter()
        poll = 0
        polls = []
        mem = collections.defaultdict(lambda: [0, 0, 0])
        for i, r in enumerate(seg):
            for k in stalls:
                st[k[6:]] += int(r[ix[k]])
            src = r[1].strip()
            if 'NANOSLEEP' in src or 'UCGABAR_WAIT' in src:
                poll += int(r[S])
                if int(r[S]) >= 10:
                    polls.append((lo + i, int(r[S])))
            toks = src.split()
            op = toks[1] if toks[0].startswith('@') else toks[0]
            if re.match(r'LDS|STS|LDG|STG|RED|LDTM|STTM|ATOM', op):
                m = mem['.'.join(op.split('.')[:2]) if re.match(r'LDS|STS', op) else op.split('.')[0]]
                m[0] += int(r[X]); m[1] += int(r[ix['L1 Wavefronts Shared']] or 0); m[2] += int(r[ix['L1 Tag Requests Global']] or 0)
        top = ' '.join(f'{k}:{100 * v / max(n, 1):.0f}%' for k, v in st.most_common(7))
        print(f'-- {rn:10s} instr {lo:5d}..{hi - 1:5d}  samples {n:5d} ({100 * n / tot:4.1f} %)  warp-instr executed {ex:10d}')
        print(f'   of its samples: waiting in a barrier poll {100 * poll / max(n, 1):4.1f} %  {polls}')
        print(f'   reasons: {top}')
        for k, (e_, w, g) in sorted(mem.items()):
            if e_:
                print(f'   {k:10s} executed {e_:9d}  shared wavefronts {w:9d}  global L1 requests {g:9d}')


if __name__ == '__main__':
    main()
