#!/usr/bin/env python
"""Static SASS census of one kernel of libsnb.so, grouped by source line.

No GPU needed: `cuobjdump -xelf` + `nvdisasm -g` on the sm_100a cubin built with
-lineinfo.  Prints instructions per source line (or per line range given as
`--regions a-b:name,...`) and the opcode mix, which is how the unrolled hot
loop of fused_features_512_kernel is budgeted before spending GPU time.

usage: python tools/sass_regions.py features 'fused_features_512_kernelILi4' [--top 40]
"""
import argparse
import collections
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '..', 'shennong_b200', '_build', 'libsnb.so')


def disassemble(unit):
    tmp = tempfile.mkdtemp(prefix='snb_sass_')
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(LIB)], cwd=tmp,
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = os.path.join(tmp, '%s.sm_100a.cubin' % unit)
    return subprocess.run(['nvdisasm', '-g', cubin], check=True, capture_output=True,
                          text=True).stdout


def census(text, kernel):
    per_line = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    inside = False
    cur = None
    for line in text.splitlines():
        if line.startswith('.text.'):
            inside = kernel in line
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            per_line[cur] += 1
            ops[cur][m.group(2)] += 1
    return per_line, ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('unit', help='features | post | pitch')
    ap.add_argument('kernel', help='substring of the mangled kernel name')
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--file', default=None, help='restrict the line table to this source file')
    ap.add_argument('--regions', default='', help='a-b:name,... line ranges of --file to aggregate')
    args = ap.parse_args()
    per_line, ops = census(disassemble(args.unit), args.kernel)
    total = sum(per_line.values())
    print('# %d SASS instructions in %s' % (total, args.kernel))
    if args.regions:
        regions = []
        for item in args.regions.split(','):
            rng, name = item.split(':')
            a, b = rng.split('-')
            regions.append((int(a), int(b), name))
        agg = collections.Counter()
        mix = collections.defaultdict(collections.Counter)
        for (f, ln), c in per_line.items():
            key = f
            if args.file is None or f == args.file:
                key = '%s:other' % f
                for a, b, name in regions:
                    if a <= ln <= b:
                        key = name
                        break
            agg[key] += c
            mix[key].update(ops[(f, ln)])
        for key, c in agg.most_common():
            top = ' '.join('%s=%d' % kv for kv in mix[key].most_common(8))
            print('%6d  %-28s %s' % (c, key, top))
        return
    rows = [(c, f, ln) for (f, ln), c in per_line.items() if args.file is None or f == args.file]
    rows.sort(reverse=True)
    for c, f, ln in rows[:args.top]:
        top = ' '.join('%s=%d' % kv for kv in ops[(f, ln)].most_common(6))
        print('%6d  %s:%d  %s' % (c, f, ln, top))


if __name__ == '__main__':
    main()
