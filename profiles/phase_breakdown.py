"""Attribute ncu source-page counters of k_seed_middles to kernel phases.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > cs.csv
    ncu -i X.ncu-rep --page source --csv --print-source sass      > sass.csv
    python profiles/phase_breakdown.py cs.csv sass.csv

Instructions of inlined seed_math.h functions are attributed to the phase of the
closest preceding seeding_kernels.cuh line in address order.
"""
import collections
import csv
import os
import sys

cs, sass_csv = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(cs)))
cur = None
addr2line = {}
line = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 8 and r[0].isdigit():
        line = (cur, int(r[0]))
        continue
    if len(r) > 8 and r[0] == '' and r[2].startswith('0x'):
        addr2line[int(r[2], 16)] = line
sass = list(csv.reader(open(sass_csv)))
hdr = sass[1]
idx = {h: i for i, h in enumerate(hdr)}
ins = []
for r in sass[2:]:
    try:
        ins.append((int(r[0], 16), int(r[idx['Instructions Executed']]), int(r[idx['Thread Instructions Executed']]),
                    int(r[idx['# Samples']]), int(r[idx['stall_barrier']])))
    except Exception:
        pass
ins.sort()
src = open(os.path.join(root, 'acts_b200/csrc/seeding_kernels.cuh')).read().split('\n')


def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    return None


marks = [(n, find(t)) for n, t in [
    ("phase0 windows", "---- phase 0: middle"), ("phase1 doublets", "---- phase 1: doublets"),
    ("phase2 sort", "---- phase 2: order both"), ("phase2 top records", "// tops: full records in sorted order"),
    ("phase3a scans", "---- phase 3a"), ("phase3b prefix max", "---- phase 3b"), ("phase3c gap", "---- phase 3c"),
    ("phase3d group", "---- phase 3d"), ("phase3e weights", "---- phase 3e"), ("phase3f heap", "---- phase 3f"),
    ("phase4 select", "---- phase 4")]]
helpers = [("find_doublets", find("__device__ __forceinline__ void find_doublets"), find("// Keys of the in-block bucket sort")),
           ("bucket sort / array scan", find("// Keys of the in-block bucket sort"), find("// true (block-uniform) when two neighbours")),
           ("tie detect / replay", find("// true (block-uniform) when two neighbours"), find("template <int CAPB, int CAPT, int CAPPOOL, int NBK>\n__global__") or find("k_seed_middles(const __grid_constant__")),
           ("block_scan_exclusive", find("__device__ __forceinline__ uint32_t block_scan_exclusive"), find("// Ordered warp-aggregated append")),
           ("warp_first_true", find("__device__ __forceinline__ uint32_t warp_first_true"), find("// Doublet search for one side"))]
kstart = find("k_seed_middles(const __grid_constant__ SeedParams p)")


def phase_of(ln):
    f, l = ln
    if f != 'seeding_kernels.cuh':
        return None
    if l >= kstart:
        p = None
        for n, m in marks:
            if m and l >= m:
                p = n
        return p or "kernel prologue / work fetch"
    for n, a, b in helpers:
        if a and b and a <= l < b:
            return n
    return "other helpers"


agg = collections.defaultdict(lambda: [0, 0, 0, 0])
cur = "?"
for a, i, t, s, b in ins:
    ln = addr2line.get(a)
    if ln:
        p = phase_of(ln)
        if p:
            cur = p
    agg[cur][0] += i
    agg[cur][1] += t
    agg[cur][2] += s
    agg[cur][3] += b
ti = sum(v[0] for v in agg.values())
ts = sum(v[2] for v in agg.values())
print("total warp instructions %.3g, thread instructions %.3g, samples %d" % (ti, sum(v[1] for v in agg.values()), ts))
print("%-30s %7s %7s %7s %16s" % ("phase", "inst%", "smp%", "lanes", "barrier% of smp"))
for k, v in sorted(agg.items(), key=lambda x: -x[1][2]):
    print("%-30s %6.1f%% %6.1f%% %6.1f %15.1f%%" % (k, 100 * v[0] / ti, 100 * v[2] / ts, v[1] / max(v[0], 1), 100 * v[3] / max(v[2], 1)))
