"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library (evidence for profiles/):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st, tensor memory), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce-add),
UTCBAR (tcgen05.commit -> mbarrier), SYNCS (mbarrier ops), UBLKCP (cp.async.bulk, non-tensor bulk copy), FHFMA (mixed-precision
fma.rn.f32.f16 / .bf16), plus the legacy tensor-core HMMA as a negative check.

    python scripts/sass_summary.py [path/to/lib.so] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "countr_b200", "lib", "libcountr_sm100.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "ELECT", "FFMA2", "MUFU", "UBLKCP", "FHFMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        if op in MNEMONICS:
            counts[cur][op] += 1
demangled = {}
try:
    names = list(counts)
    dm = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangled = dict(zip(names, dm))
except Exception:
    pass
print(f"# {os.path.relpath(lib, ROOT)}: cuobjdump -sass, arch = {', '.join(arch)}; {len(counts)} kernels")
print("# columns: " + " ".join(MNEMONICS) + " | total instructions | kernel")
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    name = demangled.get(k, k)
    name = re.sub(r"countr::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)[:90]
    print(" ".join(f"{c.get(m, 0):6d}" for m in MNEMONICS) + f" | {c['_total']:7d} | {name}")
print(" ".join(f"{tot.get(m, 0):6d}" for m in MNEMONICS) + f" | {tot['_total']:7d} | TOTAL")
