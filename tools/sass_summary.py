"""Opcode census of libkws.so per kernel (cuobjdump -sass): the Blackwell-specific instructions that prove which
hardware path a kernel takes -- UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FHFMA (mixed fp16 x fp16 + fp32 FMA) -- next to the
CUDA-core mix.  usage: python tools/sass_summary.py [libkws.so] > profiles/sass_r02.md"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "speech_recognition_b200", "libkws.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
funcs = re.split(r"\n\s*Function : ", txt)[1:]
KEY = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FHFMA", "HFMA2", "FFMA", "F2FP", "STS", "LDS", "LDG", "STG",
       "STL", "LDL"]
print("# SASS opcode census of `speech_recognition_b200/libkws.so` (r02)\n")
print(f"`cuobjdump -sass` of the tracked-source build; cubin architectures: {', '.join(arch)}.  Columns are static instruction")
print("counts per kernel (not executed counts).  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor")
print("load / store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FHFMA = fma.rn.f32.f16 (the fp32-")
print("accumulated depthwise FIR), STL / LDL = local-memory spills.\n")
print("| kernel | instr | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for f in funcs:
    name = f.split("\n")[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    m = re.search(r"(\w+_kernel(?:<[^(]*>)?)\(", dem.replace("(anonymous namespace)::", "").replace("kws::", ""))
    short = m.group(1) if m else dem[:58]
    short = re.sub(r"\(bool\)|\(int\)", "", short)
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", f, flags=re.M)
    c = collections.Counter(ops)
    if len(ops) < 50:
        continue
    print(f"| `{short[:58]}` | {len(ops)} | " + " | ".join(str(c.get(k, 0)) for k in KEY) + " |")
