"""Per-kernel count of the Blackwell-specific SASS mnemonics in the shipped library (cuobjdump -sass), the evidence table
of B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP,
tcgen05.commit -> UTCBAR, legacy tensor path -> HMMA, packed fp32 -> FFMA2/FADD2, 3-input max -> FMNMX3.
usage: python tools/sass_summary.py [lib.so] > profiles/<tag>_sass_summary.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "reverb_b200/librvb_b200.so"
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "FFMA2", "FADD2", "FMNMX3",
        "MUFU.EX2", "SYNCS", "DFMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void rvb::", "").replace("rvb::", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    for k in KEYS:
        if op.startswith(k):
            counts[cur][k] += 1
            break
print(f"# SASS evidence per kernel (`cuobjdump -sass {lib}`)\n")
print("UTCHMMA = tcgen05.mma kind::f16 (`.2CTA` = cta_group::2), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load;")
print("HMMA = legacy mma.sync (only the cross-check kernel `attention_kernel`); FFMA2 / FADD2 = packed fp32; DFMA = fp64 (prefix beam scores).\n")
used = [k for k in KEYS if any(c[k] for c in counts.values())]
print("| kernel | " + " | ".join(used) + " |")
print("|---|" + "---:|" * len(used))
for name, c in counts.items():
    if not any(c[k] for k in used):
        continue
    print(f"| `{name}` | " + " | ".join(str(c[k]) if c[k] else "" for k in used) + " |")
