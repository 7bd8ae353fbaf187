#!/usr/bin/env python
"""Warp-stall samples per CUDA source line, from an `ncu --set full --import-source on` report and the object file
the kernel came from (built with -lineinfo).

    python tools/ncu_hot_lines.py gpurun_out/big3.ncu-rep doubly-stochastic-dgp_b200/lib/layer_tc_bwd.o k_layer_bwd_tc [top_n]

ncu's CSV export carries per-SASS-instruction samples but no line numbers; `nvdisasm -g -c` carries the line of every
instruction; the two are joined by instruction index (the counts must match, i.e. the report and the object must come from
the same build).  Output: share of samples, file:line, warp-instructions executed, dominant stall reason, source text."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_selected', 'stall_branch_resolving',
        'stall_no_inst', 'stall_math', 'stall_mio', 'stall_lg', 'stall_membar']


def disasm(obj):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith('.cubin')][0]
        txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    funs, cur, name = {}, None, None
    for line in txt.splitlines():
        m = re.match(r'//-+ \.text\.(\S+) -+', line)
        if m:
            name, cur = m.group(1), None
            funs[name] = []
            continue
        if name is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
            funs[name].append(cur)
    return funs


def kernels(report, regex):
    out = subprocess.run(['ncu', '-i', report, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name',
                          'regex:' + regex], capture_output=True, text=True).stdout
    rows, res, i = list(csv.reader(out.splitlines())), [], 0
    while i < len(rows):
        if rows[i] and rows[i][0] == 'Kernel Name':
            name, hdr, j, data = rows[i][1], rows[i + 1], i + 2, []
            while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
                if len(rows[j]) == len(hdr):
                    data.append(rows[j])
                j += 1
            res.append((name, hdr, data))
            i = j
        else:
            i += 1
    return res


def mangled_matches(name, funs, n_inst):
    demangle = lambda s: subprocess.run(['c++filt', s], capture_output=True, text=True).stdout.strip()
    want = re.sub(r'\((int|bool)\)', '', name).replace(' ', '')
    for k, seq in funs.items():
        if len(seq) == n_inst and demangle(k).replace(' ', '').replace('false', '0').replace('true', '1') == want:
            return seq
    cands = [seq for seq in funs.values() if len(seq) == n_inst]
    return cands[0] if len(cands) == 1 else None


def main():
    report, obj, regex = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    funs = disasm(obj)
    src = {}
    for name, hdr, data in kernels(report, regex):
        seq = mangled_matches(name, funs, len(data))
        print(f"\n##### {name}: {len(data)} SASS instructions")
        if seq is None:
            print("   no function with the same instruction count in the object: report and build differ")
            continue
        si, ie, ki = hdr.index('# Samples'), hdr.index('Instructions Executed'), [hdr.index(k) for k in KEYS]
        agg, tot = collections.defaultdict(lambda: [0, 0] + [0] * len(KEYS)), 0
        for loc, r in zip(seq, data):
            n = int(r[si] or 0)
            tot += n
            a = agg[loc]
            a[0] += n
            a[1] += int(r[ie] or 0)
            for j, i in enumerate(ki):
                a[2 + j] += int(r[i] or 0)
        for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            if not loc:
                continue
            fn, ln = loc
            if fn not in src:
                p = os.path.join(ROOT, 'doubly-stochastic-dgp_b200', 'csrc', fn)
                src[fn] = open(p).read().splitlines() if os.path.exists(p) else []
            text = src[fn][ln - 1].strip()[:100] if ln - 1 < len(src[fn]) else ''
            dom = KEYS[max(range(len(KEYS)), key=lambda j: a[2 + j])][6:]
            print(f"{100 * a[0] / max(tot, 1):5.1f}%  {fn}:{ln:<4d} inst={a[1]:>9d}  top={dom:<16s} {text}")


if __name__ == '__main__':
    main()
