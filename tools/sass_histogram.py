"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass), the evidence that the kernels are
sm_100a code using packed fp32x2 math, 128-bit loads and vector reductions.

    python tools/sass_histogram.py [lara_b200/libsurfel_b200.so] > profiles/sass_histogram_r02.md
"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "lara_b200/libsurfel_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist, arch = None, collections.OrderedDict(), set()
for line in out.split("\n"):
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        kern = re.sub(r"\((srf::)?\w*Args\)$|\(int, float const\*, float const\*, unsigned char\*\)$", "", kern).replace("srf::", "")
        kern = kern.replace("(int)", "").replace("(bool)", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
WATCH = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU.RCP", "MUFU.EX2", "MUFU.RSQ", "MUFU.SQRT", "MUFU.LG2", "DFMA", "DMUL",
         "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "STG.E", "LDS.128", "LDS.64", "LDS", "STS.128", "STS", "REDG.E.ADD.F32x4",
         "RED.E.ADD.F32", "ATOMG", "ATOMS", "SHFL", "VOTE", "REDUX", "BAR.SYNC", "LDL", "STL", "UTMALDG", "UBLKCP", "LDGSTS", "HMMA", "UTCMMA"]
print(f"# SASS opcode histogram of `{lib}` (cuobjdump -sass), architectures: {sorted(arch)}\n")
print("Static instruction counts per kernel (not execution counts).  Columns: opcodes worth watching -- packed fp32x2 math (FFMA2/FMUL2/FADD2), "
      "MUFU, fp64, 128-bit global/shared accesses, vector reductions (REDG...F32x4), shuffles/votes, barriers, local-memory spills (LDL/STL), "
      "TMA / tensor-core opcodes (none: the path is gather/sort/blend).\n")
for k, h in hist.items():
    tot = sum(h.values())
    sel = []
    for w in WATCH:
        n = sum(v for op, v in h.items() if op == w or op.startswith(w + "."))
        if n:
            sel.append(f"{w} {n}")
    print(f"## {k}  ({tot} instructions)\n")
    print(", ".join(sel) + "\n")
    print("top opcodes: " + ", ".join(f"{op} {n}" for op, n in h.most_common(14)) + "\n")
