"""Static view of a kernel's SASS (no GPU needed): loops (backward branches) with their instruction counts, spill traffic
(STL/LDL) and where it sits.  usage: sass_regions.py <object.o> <mangled kernel name substring>"""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, body = None, {}
for line in names.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); body[cur] = []
    elif cur:
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
for fn, ins in body.items():
    if pat not in fn:
        continue
    print("==", fn, len(ins), "instructions")
    loops = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a:
            loops.append((int(m.group(1), 16), a))
    for lo, hi in sorted(loops):
        seg = [t for a, t in ins if lo <= a <= hi]
        def cnt(k): return sum(1 for t in seg if re.search(r"(^|\s)" + k, t))
        print("  loop 0x%04x-0x%04x: %4d instr | FFMA2/FMUL2/FADD2 %3d  FFMA/FMUL/FADD %3d  MUFU %2d  LDG %2d  LDL %2d STL %2d  LDS %d STS %d SHFL %d  ISETP/SEL/FSEL/VIMNMX %d" % (
            lo, hi, len(seg), cnt("F(FMA|MUL|ADD)2"), cnt("F(FMA|MUL|ADD)(\\s|\\.)"), cnt("MUFU"), cnt("LDG"), cnt("LDL"), cnt("STL"), cnt("LDS"), cnt("STS"), cnt("SHFL"),
            cnt("ISETP") + cnt("SEL") + cnt("VIMNMX")))
    print("  total LDL %d STL %d" % (sum(1 for a, t in ins if "LDL" in t), sum(1 for a, t in ins if "STL" in t)))
