#!/usr/bin/env python
"""Per-kernel census of the Blackwell-specific SASS in the built library (runs without a GPU: cuobjdump only reads the
cubin): tcgen05 MMAs (UTCHMMA / UTCQMMA ...), TMA loads / stores (UTMALDG / UTMASTG / UTMAPF), TMEM traffic
(LDTM / STTM), tcgen05.commit (UTCBAR), mbarrier waits (SYNCS), cluster / DSMEM ops, packed fp32 (FFMA2 / FADD2 / FMUL2),
MUFU.EX2, plus registers / shared memory / spills from `cuobjdump -res-usage`.

    python tools/sass_census.py [superslam_b200/lib/libsuperslam_b200.so] > profiles/sass_census_r01.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "superslam_b200", "lib", "libsuperslam_b200.so")
GROUPS = [("tcgen05.mma", r"\bUTC[A-Z]*MMA\b"), ("tma.load", r"\bUTMALDG\b"), ("tma.store", r"\bUTMASTG\b"),
          ("tma.prefetch", r"\bUTMAPF\b|\bUTMACCTL\b"), ("tmem.ld", r"\bLDTM\b"), ("tmem.st", r"\bSTTM\b"),
          ("tcgen05.commit", r"\bUTCBAR\b"), ("mbarrier", r"\bSYNCS\b"), ("tmem.alloc", r"\bUTCATOMSWS\b|\bUTCALLOC\b"),
          ("elect", r"\bELECT\b"), ("cluster", r"\bUCGABAR|\bMEMBAR\.ALL\.CLUSTER|\bMAPA\b|\bST\.ASYNC|\bSTAS\b"), ("ffma2/fadd2/fmul2", r"\bF(FMA|ADD|MUL)2\b"),
          ("hfma2", r"\bHFMA2\b"), ("mufu.ex2", r"\bMUFU\.EX2\b"), ("mufu.other", r"\bMUFU\.(?!EX2)")]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
usage = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in ln:
        usage[cur] = ln.strip()
        cur = None
kernels = collections.OrderedDict()
cur = None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in ln:
        continue
    body = ln.split("*/", 1)[-1]
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln) is None:
        continue
    kernels[cur]["instructions"] += 1
    for name, pat in GROUPS:
        if re.search(pat, body):
            kernels[cur][name] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


print(f"# SASS census of {os.path.relpath(lib, ROOT)} (sm_100a), made by tools/sass_census.py - static instruction counts per kernel")
for k, c in kernels.items():
    name = demangle(k)
    name = re.sub(r"\(.*", "", name)[:110]
    cols = "  ".join(f"{g}={c[g]}" for g, _ in GROUPS if c[g])
    print(f"\n{name}\n    instructions={c['instructions']}  {cols}\n    {usage.get(k, '')}")
