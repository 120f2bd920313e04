"""Per-kernel SASS instruction census of libnerf_b200.so (cuobjdump -sass): the mnemonics that prove what a kernel is made of.
  UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTMALDG = tensor-map TMA,
  SYNCS = mbarrier ops, REDG / ATOMG = global reductions, LDL / STL = local-memory (spill) traffic, MUFU = SFU.
usage: python tools/sass_census.py [path/to/lib.so] > profiles/rNN_sass_census.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "cse-573-minimal-nerf_b200" / "libnerf_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "REDG", "ATOMG", "LDL", "STL", "MUFU", "LDG", "STG", "LDS", "STS"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if cur and m:
        op = m.group(1)
        kernels[cur]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                kernels[cur][k] += 1
                break
print(f"# SASS census of {Path(lib).name} (sm_100a), one row per kernel; columns = static instruction counts")
print("| kernel | total | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for name, c in kernels.items():
    short = re.sub(r"\(.*", "", demangle(name)).replace("nerf::", "").replace("void ", "")
    print(f"| `{short}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "." for k in KEYS) + " |")
