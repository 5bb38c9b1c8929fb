"""Attribute the per-SASS-instruction counters of an ncu report to CUDA source lines.

ncu's CSV export of the source page carries metrics only in the SASS view, so this joins it
with `nvdisasm -g` line info of the cubin inside librlerc.so (same instruction order).

usage: python tools/ncu_lines.py report.ncu-rep kernel_mangled_substring [top_n]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rle-based-voxel-raycasting_b200", "librlerc.so")


def sass_lines(kernel_sub, cubin_sub="traverse_warp", src_holder=None):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(d) if cubin_sub in f][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], check=True, capture_output=True, text=True).stdout
    out, cur, on = [], None, False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = kernel_sub in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cubin_sub = sys.argv[4] if len(sys.argv) > 4 else "traverse_warp"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    inst = [(int(r[ci]), int(r[cs]), r[1]) for r in rows[h + 1:] if len(r) > ci and r[ci].isdigit()]
    sass = sass_lines(ksub, cubin_sub)
    if len(sass) != len(inst):
        print("warning: %d SASS instructions in cubin vs %d in report" % (len(sass), len(inst)))
    per = {}
    for (n, smp, _), (_, line, _) in zip(inst, sass):
        a = per.setdefault(line, [0, 0, 0])
        a[0] += n
        a[1] += smp
        a[2] += 1
    tot = sum(a[0] for a in per.values())
    ts = sum(a[1] for a in per.values()) or 1
    srcs = {}
    def text_of(key):
        if not key:
            return "?"
        f, n = key
        if f not in srcs:
            pth = os.path.join(ROOT, "rle-based-voxel-raycasting_b200", "csrc", f)
            srcs[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
        return srcs[f][n - 1].strip() if 0 < n <= len(srcs[f]) else "?"
    print("total warp instructions %d, samples %d" % (tot, ts))
    for key, a in sorted(per.items(), key=lambda kv: -kv[1][1 if os.environ.get("BY_SAMPLES") else 0])[:topn]:
        loc = "%s:%d" % (key[0][:14], key[1]) if key else "?"
        print("%5.1f%% inst %5.1f%% stall  %-20s (%3d sass) %s" % (100.0 * a[0] / tot, 100.0 * a[1] / ts, loc, a[2], text_of(key)[:95]))


if __name__ == "__main__":
    main()
