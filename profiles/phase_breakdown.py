"""Attribute the ncu source-page counters of k_seed_middles to kernel phases and to functions.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > cs.csv
    ncu -i X.ncu-rep --page source --csv --print-source sass      > sass.csv
    python profiles/phase_breakdown.py cs.csv sass.csv

The cuda+sass view lists a SASS instruction under EVERY source line of its inline chain (the line inside the
inlined function and the call sites up to the kernel body).  Table 1 attributes every instruction to the kernel
phase of its outermost line inside the kernel body (the "---- phase" comments of seeding_kernels.cuh; instructions
of inlined functions without a call-site frame inherit the phase of the preceding instruction); table 2 to its
innermost function (helpers of seeding_kernels.cuh, functions of seed_math.h).
"""
import collections
import csv
import os
import re
import sys

cs, sass_csv = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(cs)))
cur = None
line = None
addr2lines = collections.defaultdict(list)
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 8 and r[0].isdigit():
        line = (cur, int(r[0]))
        continue
    if len(r) > 8 and r[0] == '' and r[2].startswith('0x'):
        addr2lines[int(r[2], 16)].append(line)
sass = list(csv.reader(open(sass_csv)))
hdr = sass[1]
idx = {h: i for i, h in enumerate(hdr)}
STALLS = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ins = []
stall_of = {}
for r in sass[2:]:
    try:
        a = int(r[0], 16)
        ins.append((a, int(r[idx['Instructions Executed']]), int(r[idx['Thread Instructions Executed']]),
                    int(r[idx['# Samples']]), int(r[idx['stall_barrier']])))
        stall_of[a] = [int(r[idx[h]]) for h in STALLS]
    except Exception:
        pass
ins.sort()
ksrc = open(os.path.join(root, 'acts_b200/csrc/seeding_kernels.cuh')).read().split('\n')
msrc = open(os.path.join(root, 'acts_b200/csrc/seed_math.h')).read().split('\n')


def find(s):
    for i, l in enumerate(ksrc):
        if s in l:
            return i + 1
    return None


marks = [(n, find(t)) for n, t in [
    ("phase 0 header / carve-up", "---- phase 0: header"), ("phase 1 keys", "---- phase 1: keys"),
    ("phase 2 sort both lists", "---- phase 2: order both"), ("phase 2 top records", "// tops: full records in sorted order"),
    ("phase 3a scans", "---- phase 3a"), ("phase 3b prefix max", "---- phase 3b"), ("phase 3c gap", "---- phase 3c"),
    ("phase 3d group candidates", "---- phase 3d"), ("seedConfirmation records", "---- seedConfirmation: weights"),
    ("phase 3e weights", "---- phase 3e"), ("phase 3f selection + phase 4 seeds", "---- phase 3f")]]
kstart = find("k_seed_middles(const __grid_constant__ SeedParams p)")
kend = find("// Seed compaction (ordered): tiled exclusive scan of the per-middle counts")


def functions(src, pattern):
    out = []
    for i, l in enumerate(src):
        m = re.match(pattern, l)
        if m:
            out.append((i + 1, m.group(1)))
    return out


kfns = functions(ksrc, r'^(?:__device__|__global__).*?\b(\w+)\s*\(')
mfns = functions(msrc, r'^(?:template.*)?B2S_HD\s+[\w:<>\s\*&]+?\b(\w+)\s*\(')


def fn_of(table, l):
    name = "?"
    for first, n in table:
        if first <= l:
            name = n
    return name


def phase_of(lines):
    body = [l for f, l in lines if f == 'seeding_kernels.cuh' and kstart <= l < kend]
    if not body:
        return None
    l = max(body)  # outermost call site (lambdas are defined before their uses)
    p = "kernel prologue / work fetch"
    for n, m in marks:
        if m and l >= m:
            p = n
    return p


def innermost(lines):
    for f, l in lines:
        if f == 'seed_math.h':
            return "seed_math.h: " + fn_of(mfns, l)
    ks = [l for f, l in lines if f == 'seeding_kernels.cuh']
    helpers = [l for l in ks if not (kstart <= l < kend)]
    if helpers:
        return "kernels.cuh: " + fn_of(kfns, min(helpers))
    if ks:
        return "kernel body"
    return "cuda headers"


# helpers that are called from exactly one phase: used when an instruction has no call-site frame
HELPER_PHASE = {
    "warp_append": "phase 3c gap",
    "block_sort_both": "phase 2 sort both lists", "block_bucket_sort": "phase 2 sort both lists",
    "block_fix_ties": "phase 2 sort both lists", "block_has_ties": "phase 2 sort both lists",
    "warp_sort_replay_ties": "phase 2 sort both lists", "tie_flagged": "phase 2 sort both lists",
    "tie_less": "phase 2 sort both lists", "ordered_to_float": "phase 2 sort both lists",
    "warp_first_true": "phase 0 windows",
}


def helper_phase(lines):
    for f, l in lines:
        if f == 'seeding_kernels.cuh' and not (kstart <= l < kend):
            p = HELPER_PHASE.get(fn_of(kfns, l))
            if p:
                return p
    return None


phase = collections.defaultdict(lambda: [0, 0, 0, 0])
func = collections.defaultdict(lambda: [0, 0, 0, 0])
phase_stalls = collections.defaultdict(lambda: [0] * len(STALLS))
curp = "?"
for a, i, t, s, b in ins:
    lines = addr2lines.get(a, [])
    p = phase_of(lines)
    # An inlined helper / math instruction whose only kernel-body frame is the kernel's entry line carries no call
    # site: like instructions without line info it inherits the phase of the preceding instruction (address order).
    if p and not (p == "kernel prologue / work fetch" and innermost(lines) != "kernel body"):
        curp = p
    elif helper_phase(lines):
        curp = helper_phase(lines)
    for q, v in enumerate(stall_of[a]):
        phase_stalls[curp][q] += v
    for agg, key in ((phase, curp), (func, innermost(lines) if lines else "no line info")):
        agg[key][0] += i
        agg[key][1] += t
        agg[key][2] += s
        agg[key][3] += b
ti = sum(v[0] for v in phase.values())
ts = sum(v[2] for v in phase.values())
print("total warp instructions %.3g, thread instructions %.3g, samples %d" % (ti, sum(v[1] for v in phase.values()), ts))
for title, agg in (("kernel phase (outermost line in the kernel body)", phase), ("innermost function", func)):
    print()
    print("%-48s %7s %7s %7s %16s" % (title, "inst%", "smp%", "lanes", "barrier% of smp"))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][2]):
        if v[0] == 0 and v[2] == 0:
            continue
        print("%-48s %6.1f%% %6.1f%% %6.1f %15.1f%%" % (k[:48], 100 * v[0] / ti, 100 * v[2] / ts, v[1] / max(v[0], 1), 100 * v[3] / max(v[2], 1)))

print()
print("stall reasons per phase (share of the phase's samples; 'selected' = issuing)")
for k, v in sorted(phase.items(), key=lambda x: -x[1][2]):
    st = phase_stalls[k]
    tot = max(sum(st), 1)
    top = sorted(zip(st, STALLS), reverse=True)[:5]
    print("%-28s %s" % (k[:28], ", ".join("%s %.0f%%" % (n.replace('stall_', ''), 100 * c / tot) for c, n in top)))
