"""Instruction / stall-sample share of the warp kernel by source region.
usage: python tools/ncu_regions.py report.ncu-rep [kernel_sub]"""
import csv, io, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines
rep = sys.argv[1]; ksub = sys.argv[2] if len(sys.argv) > 2 else "k_traverse_wILb0"
src = open(os.path.join(ncu_lines.ROOT, "rle-based-voxel-raycasting_b200", "csrc", "traverse_warp.cu")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r"// ---- ([A-Z0-9]+)\. ", l)
    if m: marks.append((i, m.group(1)))
    if "---- owner lane" in l: marks.append((i, "B.owner"))
    if "---- long column" in l: marks.append((i, "B.long"))
    if "RLERC_DDA_STEP(j) RLERC_DDA_STEP" in l: marks.append((i - 25, "A.dda")); marks.append((i + 8, "A.geom"))
    if l.startswith("k_traverse_w("): marks.append((i, "setup"))
    if "__noinline__ int coop_span" in l: marks.append((i, "coop_span"))
    if "__noinline__ void long_column" in l: marks.append((i, "long_column"))
    if "sky sentinel on every pixel" in l: marks.append((i, "epilogue"))
marks.sort()
def region(f, n):
    if f != "traverse_warp.cu": return f
    r = "head"
    for ln, name in marks:
        if n >= ln: r = name
    return r
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]; ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
inst = [(int(r[ci]), int(r[cs])) for r in rows[h + 1:] if len(r) > ci and r[ci].isdigit()]
sass = ncu_lines.sass_lines(ksub)
acc = {}
for (n, s), (_, key, _) in zip(inst, sass):
    name = region(*key) if key else "?"
    a = acc.setdefault(name, [0, 0]); a[0] += n; a[1] += s
tot = sum(a[0] for a in acc.values()); ts = sum(a[1] for a in acc.values()) or 1
print("total %d warp instructions" % tot)
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print("%-22s %5.1f%% inst %5.1f%% stall samples" % (k, 100.0 * v[0] / tot, 100.0 * v[1] / ts))
