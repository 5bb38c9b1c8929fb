"""Host side of the path: .rle4 loader/writer, pointer map, compressor, tiler."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from util import levels_of


def test_save_load_round_trip(R, scene_small, tmp_path):
    f = str(tmp_path / "a.rle4")
    scene_small.save(f)
    back = R.RLE4.load(f)
    assert back.nummaps == scene_small.nummaps
    for m in range(back.nummaps):
        a, b = scene_small.level(m), back.level(m)
        assert a[:3] == b[:3]
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    # the file is the reference layout: int32 nummaps, then per level 4 x int32 + ushort slabs (Rle4.cpp:227-238)
    raw = open(f, "rb").read()
    assert struct.unpack_from("<i", raw, 0)[0] == scene_small.nummaps
    sx, sy, sz, n = struct.unpack_from("<4i", raw, 4)
    assert (sx, sy, sz) == (64, 64, 64) and n == len(scene_small.level(0)[4])


def test_pointer_map_matches_oracle(rb, scene_mid):
    for sx, sy, sz, mp, sl in levels_of(scene_mid):
        assert np.array_equal(rb.orc_build_map(sl, sx, sz), mp)


def test_pointer_map_semantics(scene_small):
    sx, sy, sz, mp, sl = scene_small.level(0)
    ofs = mp[0::2].astype(np.int64)
    cnt = (mp[1::2] & 0xffff).astype(np.int64)
    first = mp[1::2] >> 16
    assert np.array_equal(sl[ofs], cnt)                       # header word 0 = run count
    nvox = sl[ofs + 1].astype(np.int64)
    assert np.array_equal(ofs[1:], (ofs + cnt + nvox + 2)[:-1])  # columns are packed back to back, x fastest
    assert np.array_equal(first[:-1], sl[ofs[:-1] + 2])         # first run rides in the map entry
    # every column's runs cover at most sy voxels; attributes = sum of solid counts
    c = int(np.argmax(cnt))
    runs = sl[ofs[c] + 2: ofs[c] + 2 + cnt[c]]
    assert int((runs & 1023).sum() + (runs >> 10).sum()) <= sy
    assert int((runs >> 10).sum()) == nvox[c]


def test_loader_rejects_bad_files(R, scene_small, tmp_path):
    with pytest.raises(R.RlercError):
        R.RLE4.load(str(tmp_path / "missing.rle4"))
    f = str(tmp_path / "a.rle4")
    scene_small.save(f)
    raw = open(f, "rb").read()
    open(str(tmp_path / "trunc.rle4"), "wb").write(raw[: len(raw) // 2])
    with pytest.raises(R.RlercError):
        R.RLE4.load(str(tmp_path / "trunc.rle4"))
    open(str(tmp_path / "levels.rle4"), "wb").write(struct.pack("<i", 99) + raw[4:])
    with pytest.raises(R.RlercError):
        R.RLE4.load(str(tmp_path / "levels.rle4"))
    open(str(tmp_path / "empty.rle4"), "wb").write(b"")
    with pytest.raises(R.RlercError):
        R.RLE4.load(str(tmp_path / "empty.rle4"))
    # a column header that runs past the end of the stream
    bad = bytearray(raw)
    struct.pack_into("<H", bad, 4 + 16, 60000)
    open(str(tmp_path / "overrun.rle4"), "wb").write(bytes(bad))
    with pytest.raises(R.RlercError):
        R.RLE4.load(str(tmp_path / "overrun.rle4"))


def test_from_maps_copies(R, scene_small):
    maps = [scene_small.map4(m)[0] for m in range(scene_small.nummaps)]
    cp = R.RLE4.from_maps(maps)
    for m in range(cp.nummaps):
        a, b = scene_small.level(m), cp.level(m)
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
        assert a[4].ctypes.data != b[4].ctypes.data


def test_compress_run_splitting_and_empty_columns(R):
    """Long air gaps and long solid runs are split at the 10-bit / 6-bit limits (Rle4.cpp:166-182)."""
    sx, sy, sz = 8, 2048, 8
    vol = np.zeros((sz, sy, sx), np.uint8)
    vol[2, 1500:1700, 3] = 1          # gap of 1500 (> 1023) then 200 surface voxels (> 63): a 1-voxel-wide pole
    vol[5, 10, 5] = 1
    bits = np.packbits(vol.reshape(-1, 8)[:, ::-1], axis=1).reshape(-1)   # bit (x&7) of byte lin>>3
    s = R.RLE4.compress_all(bits, sx, sy, sz)
    _, _, _, mp, sl = s.level(0)
    c = 3 + 2 * sx
    o, n = int(mp[c * 2]), int(mp[c * 2 + 1] & 0xffff)
    runs = sl[o + 2: o + 2 + n]
    assert list(runs[:2]) == [1023, 63 * 1024 + (1500 - 1023)]
    assert int((runs >> 10).sum()) == 200 and int(sl[o + 1]) == 200
    assert int((runs & 1023).sum()) == 1500
    e = 0                                                      # an empty column: header only
    assert int(mp[e * 2 + 1] & 0xffff) == 0 and int(sl[int(mp[e * 2]) + 1]) == 0


def test_compress_is_deterministic_and_thread_count_independent(R):
    a = R.RLE4.synth(0, 64, 64, 64, seed=7)
    os.environ["OMP_NUM_THREADS"] = "1"
    b = R.RLE4.synth(0, 64, 64, 64, seed=7)
    for m in range(a.nummaps):
        assert np.array_equal(a.level(m)[4], b.level(m)[4])


def test_compress_wide_path_equals_generic_path(R):
    """sx % 64 == 0 uses the 64-voxels-at-a-time interior test; narrower volumes the 27-tap loop."""
    n = 64 * 64 * 64 // 8
    v = np.zeros(n, np.uint8); c1 = np.zeros(n, np.uint8); c2 = np.zeros(n, np.uint8)
    R._check(R.lib().rlerc_synth_volume(0, 64, 64, 64, 3, v.ctypes.data, c1.ctypes.data, c2.ctypes.data))
    wide = R.RLE4.compress_all(v, 64, 64, 64, c1, c2)
    # the same voxels inside a 72-wide volume (sx % 64 != 0): its level-0 columns x < 63 must agree
    vol = np.unpackbits(v.reshape(-1, 1), axis=1)[:, ::-1].reshape(64, 64, 64)
    big = np.zeros((64, 64, 72), np.uint8); big[:, :, :64] = vol
    bits = np.packbits(big.reshape(-1, 8)[:, ::-1], axis=1).reshape(-1)
    def expand(c):
        cc = np.unpackbits(c.reshape(-1, 1), axis=1)[:, ::-1].reshape(64, 64, 64)
        b = np.zeros((64, 64, 72), np.uint8); b[:, :, :64] = cc
        return np.packbits(b.reshape(-1, 8)[:, ::-1], axis=1).reshape(-1)
    narrow = R.RLE4.compress_all(bits, 72, 64, 64, expand(c1), expand(c2))
    _, _, _, mw, sw = wide.level(0)
    _, _, _, mn, sn = narrow.level(0)
    for z in (0, 17, 63):
        for x in (0, 1, 31, 62):
            cw, cn = x + z * 64, x + z * 72
            ow, on = int(mw[cw * 2]), int(mn[cn * 2])
            lw = 2 + int(sw[ow]) + int(sw[ow + 1])
            assert np.array_equal(sw[ow: ow + lw], sn[on: on + lw]), (x, z)


def test_tile_layout(R, scene_small):
    t = scene_small.tile(2, 3)
    for m in range(scene_small.nummaps):
        sx, sy, sz, mp, sl = scene_small.level(m)
        tx, ty, tz, tmp, tsl = t.level(m)
        assert (tx, ty, tz) == (sx * 2, sy, sz * 3)
        assert len(tsl) == len(sl) * 6
        # every tiled column decodes to the source column it wraps to
        for (x, z) in ((0, 0), (sx + 1 if sx > 1 else 0, 0), (2 * sx - 1, 3 * sz - 1), (sx // 2, sz + sz // 2)):
            cs = (x % sx) + (z % sz) * sx
            ct = x + z * tx
            os_, ot = int(mp[cs * 2]), int(tmp[ct * 2])
            ln = 2 + int(sl[os_]) + int(sl[os_ + 1])
            assert np.array_equal(sl[os_: os_ + ln], tsl[ot: ot + ln])
            assert (mp[cs * 2 + 1] & 0xffff) == (tmp[ct * 2 + 1] & 0xffff)
    with pytest.raises(R.RlercError):
        scene_small.tile(0, 1)


def test_synth_scene_shape(scene_mid):
    assert scene_mid.nummaps == 7
    dims = [scene_mid.level(m)[:3] for m in range(7)]
    assert dims == [(128 >> m,) * 3 for m in range(7)]
    sx, sy, sz, mp, sl = scene_mid.level(0)
    cnt = mp[1::2] & 0xffff
    assert cnt.min() >= 1                      # the terrain covers every column


def test_direct_rle_scene(R, rb, scene_rle, tmp_path):
    """rlerc_synth_rle writes columns straight into the .rle4 layout (BASELINE config 4 without a 32 GiB volume):
    valid format (decodes column by column, run-splitting rules of Rle4.cpp:166-182), halved mip chain, the
    short-run band is there, the file round-trips through the reference's own loader layout, and the budget of the
    full-size scene fits the format's int32 slab count."""
    assert scene_rle.nummaps == 6
    for m in range(scene_rle.nummaps):
        sx, sy, sz, mp, sl = scene_rle.level(m)
        assert (sx, sy, sz) == (256 >> m,) * 3
        assert np.array_equal(rb.orc_build_map(sl, sx, sz), mp)
        # walk every column: header, runs, attributes; heights stay inside the volume
        ofs = 0
        for _ in range(sx * sz):
            n_runs, n_vox = int(sl[ofs]), int(sl[ofs + 1])
            runs = sl[ofs + 2: ofs + 2 + n_runs].astype(np.int64)
            assert int((runs >> 10).sum()) == n_vox
            assert int(((runs & 1023) + (runs >> 10)).sum()) <= sy
            ofs += 2 + n_runs + n_vox
        assert ofs == len(sl)
    cnt = scene_rle.level(0)[3][1::2] & 0xffff
    assert cnt.max() >= 33 and 1.0 < cnt.mean() < 8.0           # band columns have 1 + 32 runs, most have one or two
    f = str(tmp_path / "rle.rle4")
    scene_rle.save(f)
    back = R.RLE4.load(f)
    for m in range(back.nummaps):
        a, b = scene_rle.level(m), back.level(m)
        assert a[:3] == b[:3] and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    # same statistics at 1024^2 columns: ushorts per column x 16384^2 must stay below 2^31 (Rle4.cpp:237)
    big = R.RLE4.synth_rle(1024, 1024, 1024, seed=42, band_every=48)
    per_col = len(big.level(0)[4]) / float(1024 * 1024)
    assert per_col * 16384.0 * 16384.0 < 2.0 ** 31
    with pytest.raises(R.RlercError):
        R.RLE4.synth_rle(100, 256, 256)


def test_untrusted_scenes_are_validated(R, scene_small, tmp_path):
    """a header may claim any size, a pointer map any offset.  The loader compares the claim with the
    file before allocating, from_maps checks every column against the stream, and a column whose runs claim more solid
    voxels than it has attributes (the reference's undefined tiny mip levels) loads, with the device copy padded."""
    import struct
    p = str(tmp_path / "a.rle4")
    scene_small.save(p)
    raw = bytearray(open(p, "rb").read())
    # level 0 claims 2^31 - 1 slabs: refused from the file length, no 4 GiB allocation
    huge = bytearray(raw)
    struct.pack_into("<i", huge, 4 + 12, 0x7fffffff)
    q = str(tmp_path / "huge.rle4")
    open(q, "wb").write(huge)
    with pytest.raises(R.RlercError, match="truncated"):
        R.RLE4.load(q)
    # a pointer map that points outside the stream
    m4, n = scene_small.map4(0)
    sx, sy, sz, mp, sl = scene_small.level(0)
    bad_map = mp.copy()
    bad_map[2 * 7] = n + 5
    maps = [scene_small.map4(m)[0] for m in range(scene_small.nummaps)]
    maps[0].map = bad_map.ctypes.data
    with pytest.raises(R.RlercError, match="outside the slab stream"):
        R.RLE4.from_maps(maps)
    # a column whose voxel count is smaller than its runs claim: loads (as in the reference's own files)
    sx, sy, sz, mp, sl = scene_small.level(scene_small.nummaps - 1)
    ofs = int(mp[0])
    assert sl[ofs] >= 1
    body = bytearray(raw)
    # find that level in the file: walk the headers
    pos = 4
    for m in range(scene_small.nummaps - 1):
        nsl = struct.unpack_from("<I", body, pos + 12)[0]
        pos += 16 + 2 * nsl
    first_run = struct.unpack_from("<H", body, pos + 16 + 2 * (ofs + 2))[0]
    struct.pack_into("<H", body, pos + 16 + 2 * (ofs + 2), (first_run & 1023) | (63 << 10))     # 63 solid voxels claimed
    z = str(tmp_path / "inconsistent.rle4")
    open(z, "wb").write(body)
    s = R.RLE4.load(z)
    assert s.nummaps == scene_small.nummaps
