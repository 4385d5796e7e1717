"""Per-source-line instruction / stall-sample shares from an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [min_pct]
"""
import collections, csv, io, os, subprocess, sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, cur_line, hdr = None, None, None
per = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix_inst, ix_samp, ix_thr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if r[0] != "":
        try:
            cur_line = int(r[0])
        except ValueError:
            pass
        continue
    if hdr is None or len(r) <= ix_thr or r[2] in ("...", ""):
        continue
    try:
        n, s, t = int(r[ix_inst]), int(r[ix_samp]), int(r[ix_thr])
    except ValueError:
        continue
    d = per.setdefault((cur_file, cur_line), [0, 0, 0])
    d[0] += n; d[1] += s; d[2] += t
tot = sum(d[0] for d in per.values()); tots = sum(d[1] for d in per.values())
print(f"total warp instructions {tot/1e6:.1f} M, samples {tots}")
cache = {}
for (f, l), d in sorted(per.items(), key=lambda kv: (kv[0][0] or "", kv[0][1] or 0)):
    if d[0] >= tot * min_pct / 100 or d[1] >= tots * min_pct / 100:
        if f not in cache:
            p = f if f and os.path.isfile(f) else os.path.join("lara_b200/csrc", os.path.basename(f or ""))
            cache[f] = open(p).read().split("\n") if os.path.isfile(p) else []
        src = cache[f][l - 1].strip()[:90] if l and l <= len(cache[f]) else ""
        print(f"{os.path.basename(f or '?'):18s}{l:5d} inst {d[0]/1e6:7.2f}M {100*d[0]/tot:5.1f}%  samp {100*d[1]/max(tots,1):5.1f}%  lanes {d[2]/max(d[0],1):4.1f} | {src}")
