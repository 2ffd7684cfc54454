"""SASS evidence: per kernel of ssc_b200/libssc_b200.so, how often the instructions DESIGN.md claims appear.

    python tools/sass_census.py > profiles/r2_sass_census.txt      (needs only cuobjdump, no GPU)
"""
import collections
import re
import subprocess
import sys

LIB = "ssc_b200/libssc_b200.so"
WATCH = ["DMMA.8x8x4", "UBLKCP", "SYNCS", "REDG.E.ADD.F64", "ATOM.E.ADD.F64", "ATOMS", "LDG.E.64", "LDG.E.128", "LDG.E.EF.64", "STG.E.EF", "STG.E.128",
         "LDS.128", "STS.128", "CREDUX", "REDUX", "BAR.SYNC", "DFMA", "MUFU.RCP64H", "UTCMMA", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fn, counts, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)
        counts[fn] = collections.Counter()
        sizes[fn] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Za-z0-9_.]+)", line)
    if m and fn:
        sizes[fn] += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                counts[fn][w] += 1
print("SASS census of %s (sm_100a), `cuobjdump -sass`: instruction counts per kernel for the mnemonics DESIGN.md refers to." % LIB)
print("DMMA.8x8x4 = mma.sync.m8n8k4.f64 (fp64 tensor cores); UBLKCP = cp.async.bulk (TMA bulk copy) with SYNCS = mbarrier; REDG.E.ADD.F64 = fp64 reduction at L2")
print("(atomicAdd whose result is unused); LDG.E.EF / STG.E.EF = evict-first loads / stores; CREDUX / REDUX = redux.sync; no UTCMMA / HMMA: fp64 has no tcgen05 kind.\n")
for fn, c in counts.items():
    if not fn.startswith("void ssc_") and not fn.startswith("ssc_"):
        continue
    print("%-95s %5d instr  %s" % (fn[:95], sizes[fn], "  ".join("%s x%d" % (k, v) for k, v in c.items()) or "-"))
