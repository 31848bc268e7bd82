#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell-native path (tcgen05 MMAs, TMEM loads/stores, TMA
tensor and bulk copies, cluster barriers, NVSwitch multimem), from `cuobjdump -sass libosb.so`.

    python tools/sass_summary.py > profiles/rNN_sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "openset_imagenet_b200", "csrc", "libosb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
GROUPS = collections.OrderedDict([
    ("UTCHMMA (tcgen05.mma, bf16)", r"\bUTCHMMA"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("LDTM (tcgen05.ld)", r"\bLDTM"),
    ("STTM (tcgen05.st)", r"\bSTTM"), ("UTMALDG (TMA tensor load)", r"\bUTMALDG"), ("UTMASTG (TMA tensor store)", r"\bUTMASTG"), ("UTMAREDG (TMA tensor reduce-add)", r"\bUTMAREDG"),
    ("UTMAPF (TMA L2 prefetch)", r"\bUTMAPF"), ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"),
    ("UCGABAR (cluster barrier)", r"\bUCGABAR"), ("multimem (LDGMC / STG.MC / REDG.MC)", r"MULTIMEM|LDGMC|\.MC\b|\bSTGMC|\bREDGMC"),
    ("HMMA (mma.sync)", r"\bHMMA"), ("ACQBULK / fence.proxy", r"\bACQBULK|FENCE\.VIEW\.ASYNC"),
])
kern, counts, total = None, {}, collections.Counter()
arch = set(re.findall(r"arch = (sm_\w+)", out))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None or "/*" not in line:
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m:
        continue
    op = m.group(1)
    counts[kern]["__all__"] += 1
    for g, pat in GROUPS.items():
        if re.search(pat, op):
            counts[kern][g] += 1
            total[g] += 1
def demangle(n):
    try:
        d = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().replace("void ", "")
        d = re.sub(r"\((?:int|bool)\)", "", d)                   # template arguments print as (int)256, (bool)1
        depth = 0
        for i, ch in enumerate(d):                                # cut the function parameter list, keep the template arguments
            depth += ch == "<"
            depth -= ch == ">"
            if ch == "(" and depth == 0:
                return d[:i].replace("osb::head::", "").replace("osb::", "")
        return d
    except Exception:
        return n
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: architectures {sorted(arch)}, {len(counts)} kernels")
print("# totals: " + ", ".join(f"{g.split(' ')[0]} {total[g]}" for g in GROUPS))
cols = list(GROUPS)
print(f"{'kernel':<78} {'instr':>7} " + " ".join(f"{c.split(' ')[0][:8]:>8}" for c in cols))
for k, c in sorted(counts.items(), key=lambda kv: -sum(v for g, v in kv[1].items() if g != '__all__')):
    if not any(c[g] for g in cols):
        continue
    name = demangle(k)
    name = name if len(name) <= 77 else name[:74] + "..."
    print(f"{name:<78} {c['__all__']:>7} " + " ".join(f"{c[g]:>8}" for g in cols))
