"""The frame's LOD / z schedule (csrc/kernels.cuh LodSched, filled by capi.cu fill_lod_sched) against a per-crossing
simulation of the integer half of the reference's traversal loop (R/src/Cuda_Render.h:343-367): z += dz per crossing,
dz / mapswitch / mip double whenever z > mapswitch, the ray plane ends when z + dz would pass z_far.  The traversal
kernels restart the DDA of a ray plane at every 32nd crossing from this table, so it has to be exact."""
import ctypes as C
import random

import pytest


def simulate(mapswitch0, z_far, mountain):
    """Per-crossing reference loop; returns [(z before the crossing, nsw)] for every crossing made."""
    nsw = 0
    yms = mountain
    while yms > 512.0:
        nsw += 1
        yms *= 0.5
    z, dz, ms = 0, 1 << nsw, mapswitch0 << nsw
    out = []
    while True:
        while z > ms:                       # Cuda_Render.h:343-365
            nsw += 1; dz *= 2; ms *= 2
        if z + dz > z_far:                  # Cuda_Render.h:366-367: z += dz; if (z > z_far) return
            break
        out.append((z, nsw))
        z += dz
    return out


@pytest.mark.parametrize("case", [(1728, 80000, -150.0), (921, 80000, -818.0), (3456, 80000, 900.0), (1, 100, 0.0),
                                  (7, 5000, 2000.0), (6900, 80000, -50.0), (100000, 80000, -1.0), (230, 1 << 30, 0.0)])
def test_lod_sched_matches_per_crossing_loop(R, case):
    ms0, zfar, mountain = case
    out = (C.c_int * 99)()
    assert R.lib().rlerc_debug_lod_sched(ms0, zfar, C.c_float(mountain), out) == 0
    nphase, k_total = out[0], out[1]
    ph_k, ph_z, ph_nsw = out[2:35], out[35:67], out[67:99]
    want = simulate(ms0, zfar, mountain)
    assert k_total == len(want)
    assert ph_k[nphase] == k_total
    # every crossing: z and nsw from the table as the kernel derives them (dda_chunk)
    idx = sorted(set(list(range(0, k_total, 32)) + [k for p in range(nphase) for k in (ph_k[p] - 1, ph_k[p], ph_k[p] + 1) if 0 <= k < k_total]
                     + random.Random(1).sample(range(k_total), min(k_total, 500))))
    for k in idx:
        p = 0
        while p + 1 < nphase and ph_k[p + 1] <= k:
            p += 1
        nsw = ph_nsw[p]
        z = ph_z[p] + ((k - ph_k[p]) << nsw)
        # the kernel re-enters the loop at crossing k with this state and lets `while (z > mapswitch)` run: the state
        # BEFORE that adjustment may lag the simulated one by the switches made exactly at crossing k
        ms = ms0 << nsw
        while z > ms:
            nsw += 1; ms *= 2
        assert (z, nsw) == want[k], (k, p)


def test_lod_sched_rejects_what_would_never_terminate(R):
    out = (C.c_int * 99)()
    assert R.lib().rlerc_debug_lod_sched(1728, 80000, C.c_float(float("inf")), out) < 0
    assert R.lib().rlerc_debug_lod_sched(1728, 80000, C.c_float(1e30), out) < 0
