"""Instruction / stall-sample share of k_traverse_f by kernel stage: buckets the per-SASS counters of an ncu
report (tools/ncu_sass.py output on stdin) by the stage markers found in the CURRENT sources.
usage: python tools/ncu_sass.py rep.ncu-rep k_traverse_fILb0ELb0 traverse_filter | python tools/ncu_stages.py"""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "rle-based-voxel-raycasting_b200", "csrc")


def marks(fname, table):
    out = []
    for i, l in enumerate(open(os.path.join(CS, fname)).read().splitlines(), 1):
        for pat, name in table:
            if pat in l:
                out.append((i, name))
    return sorted(out)


M = {
    "traverse_filte": marks("traverse_filter.cu", [
        ("struct DdaQ {", "DDA (lane-parallel chunk re-expansion)"), ("// ---- pre-pass:", "DDA (lane-parallel chunk re-expansion)"),
        ("// ---- FILTER, geometry half", "F geometry + pointer-map gather"), ("// ---- FILTER, test half", "F first-run test + queue"),
        ("// entry of the live-column queue", "F first-run test + queue"),
        ("// ---- C1: take a live column", "C1 run-word loads"), ("// ---- C2: project the runs", "C2 run projection"),
        ("__device__ __forceinline__ void filter_ray_init", "set-up / loop / epilogue")]),
    "traverse_commo": marks("traverse_common.cuh", [
        ("__device__ __noinline__ int coop_span(", "cooperative span shaders"), ("void long_column(", "long_column (lane <-> run)"),
        ("// ---- B0. rising-horizon", "B0 rising-horizon path"),
        ("// ---- B1. ownership-resolved", "B1 ownership-resolved path"), ("// ---- E. one event-loop iteration", "event loop (owner lane)"),
        ("// SHADE (S): the short spans", "S deferred shading")]),
}
acc, T, C = {}, 0, 0
for ln in sys.stdin:
    p = ln.split()
    if len(p) < 4 or not p[1].isdigit():
        continue
    n, s, key = int(p[1]), int(p[2]), p[3]
    f, _, l = key.partition(":")
    name = {"device_common.": "f2i / first_clear / ray set-up"}.get(f, "warp intrinsics (shuffle, ballot, ldg)")
    if f in M and l.isdigit():
        name = "set-up / loop / epilogue" if f == "traverse_filte" else "helpers"
        for ml, mn in M[f]:
            if int(l) >= ml:
                name = mn
    a = acc.setdefault(name, [0, 0]); a[0] += n; a[1] += s
    C += n; T += s
print("total %d warp instructions, %d stall samples" % (C, T))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print("%-44s %5.1f%% inst %5.1f%% stall samples" % (k, 100.0 * v[0] / C, 100.0 * v[1] / max(T, 1)))
