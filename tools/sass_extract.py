"""Per-kernel SASS summary of the built library (cuobjdump -sass), so that the Blackwell-native claims do not rest
on a binary that is git-ignored: instruction count, registers, and the mnemonics that matter -- UTMALDG (TMA tensor
loads), SYNCS (mbarrier), LDGSTS (cp.async), STG.E.128 / .64, LDG / LD.E (peer loads in the strip variants), the
packed-halfword and byte-SIMD arithmetic, shuffles.  usage: python tools/sass_extract.py [library] > profiles/rNN_sass_extract.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "gpu_video_codec_b200", "libhevc_deblock_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(5)))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "STG.E.128", "STG.E.64", "STG.E ", "LDG.E", "LD.E", "LDS", "STS", "SHFL", "VABSDIFF4",
         "VIADDMNMX", "VIMNMX", "VIADD", "IDP", "PRMT", "LOP3", "IMAD", "IADD3", "SHF", "ATOMG", "HMMA", "UTC"]
cur, ops = None, collections.OrderedDict()
for ln in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        ops[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        ops[cur]["_total"] += 1
        ops[cur][m.group(1)] += 1
print("library: %s" % os.path.basename(lib))
print("compiled for: %s" % ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))))
for fn, c in ops.items():
    name = demangle(fn)
    if "hdb::" not in name:
        continue
    r = regs.get(fn, (0, 0, 0))
    print("\n%s\n  %d instructions, %d registers, %d bytes stack, %d bytes local" % (name[:150], c["_total"], r[0], r[1], r[2]))
    row = []
    for w in WATCH:
        n = sum(v for k, v in c.items() if k.startswith(w.strip()) and (w != "STG.E " or k in ("STG.E", "STG.E.U8")))
        if n:
            row.append("%s %d" % (w.strip(), n))
    print("  " + " | ".join(row))
    peer = [k for k in c if k.startswith("LD.E") or "STRONG" in k]
    if peer:
        print("  strong / generic loads: " + ", ".join("%s %d" % (k, c[k]) for k in peer))
