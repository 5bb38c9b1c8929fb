// Host-side scene: .rle4 reader/writer, pointer-map build, RLE compressor + mip pyramid,
// physical tiling.  Mirrors the surface of the reference's RLE4 (R/src/Rle4.h:25-52):
//   rlerc_scene_load      <-> RLE4::load          R/src/Rle4.cpp:244-384
//   rlerc_scene_save      <-> RLE4::save          R/src/Rle4.cpp:220-242
//   rlerc_scene_compress  <-> RLE4::compress_all  R/src/Rle4.cpp:16-198 + Tree::get_mipmap tree.h:23-85
// Written from the format description in SURVEY.md §8a; 64-bit clean (the reference is ILP32).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "rlerc_internal.h"

namespace rlerc {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}


// Walk the slab stream column by column (x fastest, then z) and emit, per column,
// {offset of the column header, n_runs | first_run << 16}.  The "first run" word is
// whatever ushort sits behind the two header words, even for an empty column — the
// reference stores it unconditionally and the map must be byte-identical.
int build_pointer_map(Level& lv)
{
	const uint64_t ncol = (uint64_t)lv.sx * (uint64_t)lv.sz;
	const uint64_t n = lv.slabs.size();
	lv.map.assign(ncol * 2, 0u);
	uint64_t ofs = 0;
	const uint16_t* s = lv.slabs.data();
	for (uint64_t c = 0; c < ncol; c++)
	{
		if (ofs + 2 > n) { set_error("slab stream ends inside column %llu", (unsigned long long)c); return RLERC_ERR_FORMAT; }
		if (ofs > 0xffffffffull) { set_error("slab offset exceeds 32 bits"); return RLERC_ERR_FORMAT; }
		const uint32_t n_runs = s[ofs];
		const uint32_t n_vox = s[ofs + 1];
		const uint32_t first = (ofs + 2 < n) ? s[ofs + 2] : 0u;
		lv.map[c * 2 + 0] = (uint32_t)ofs;
		lv.map[c * 2 + 1] = n_runs + (first << 16);
		ofs += (uint64_t)n_runs + n_vox + 2;
	}
	if (ofs > n) { set_error("slab stream shorter than its columns claim"); return RLERC_ERR_FORMAT; }
	return RLERC_OK;
}

// Scenes that come from outside (a file, caller-provided Map4 arrays) are not trusted: every column's header, runs and
// attributes have to lie inside the slab stream, and its map entry has to agree with the stream.  A column whose runs
// claim more solid voxels than its attribute count (n_vox) is NOT an error — the reference's compressor leaves the mip
// levels narrower than 8 voxels undefined and its own files have to load — but the traversal's
// attribute gather send[u], u < sum of the solid counts (Cuda_Render.h:498-499,707-713), would then read past the column;
// gather_pad is the number of ushorts by which such a read can leave the STREAM, and the device copy is padded by it.
int validate_columns(Level& lv)
{
	const uint64_t ncol = (uint64_t)lv.sx * (uint64_t)lv.sz;
	const uint64_t n = lv.slabs.size();
	if (lv.map.size() != ncol * 2) { set_error("pointer map has %llu words for %llu columns", (unsigned long long)lv.map.size(), (unsigned long long)ncol); return RLERC_ERR_FORMAT; }
	const uint16_t* s = lv.slabs.data();
	uint64_t pad = 0;
	int bad = 0;
	#pragma omp parallel for schedule(static) reduction(max : pad) reduction(+ : bad)
	for (long long c = 0; c < (long long)ncol; c++)
	{
		const uint64_t ofs = lv.map[c * 2];
		if (ofs + 2 > n) { bad++; continue; }
		const uint64_t n_runs = s[ofs], n_vox = s[ofs + 1];
		if ((lv.map[c * 2 + 1] & 0xffffu) != n_runs || ofs + 2 + n_runs + n_vox > n) { bad++; continue; }
		if (n_runs > 0 && (lv.map[c * 2 + 1] >> 16) != s[ofs + 2]) { bad++; continue; }
		uint64_t solid = 0;
		for (uint64_t r = 0; r < n_runs; r++) solid += s[ofs + 2 + r] >> 10;
		if (solid > n_vox)
		{
			const uint64_t end = ofs + 2 + n_runs + solid;               // one past the farthest attribute the runs can address
			if (end > n && end - n > pad) pad = end - n;
		}
	}
	if (bad) { set_error("%d columns of a %d x %d level point outside the slab stream or disagree with it", bad, lv.sx, lv.sz); return RLERC_ERR_FORMAT; }
	lv.gather_pad = pad;
	return RLERC_OK;
}

// ---------------------------------------------------------------------------------------
// Bit volumes.  Voxel (x,y,z) lives at bit (x&7) of byte (x + y*sx + z*sx*sy)>>3, the
// addressing the reference's Tree uses (tree.h:60-80, Rle4.cpp:86).  For sx%8==0 this is
// plain bit-linear addressing; for the tiny tail levels (sx<8) distinct voxels alias, and
// the reference's pyramid is built on exactly that aliasing, so it is kept.
struct BitVol {
	int sx = 0, sy = 0, sz = 0;
	std::vector<uint8_t> v, c1, c2;
	bool color = false;
	void init(int x, int y, int z, bool col)
	{
		sx = x; sy = y; sz = z; color = col;
		size_t bytes = ((size_t)x * y * z + 7) / 8 + 8;
		v.assign(bytes, 0);
		if (col) { c1.assign(bytes, 0); c2.assign(bytes, 0); }
	}
};

static inline int vbit(const uint8_t* m, size_t lin, int x) { return (m[lin >> 3] >> (x & 7)) & 1; }

// 2x reduction: a coarse voxel is set when at least three of its eight children are; its
// two material bits come from the third set child in x-major, then y, then z order.
static void build_mip(const BitVol& p, BitVol& o)
{
	o.init(p.sx / 2, p.sy / 2, p.sz / 2, p.color);
	const int sx = o.sx, sy = o.sy, sz = o.sz;
	const size_t sxy = (size_t)sx * sy;
	const bool par = (sx % 8 == 0);
	#pragma omp parallel for schedule(static) if (par)
	for (int k = 0; k < sz; k++)
	for (int j = 0; j < sy; j++)
	for (int i = 0; i < sx; i++)
	{
		int need = 3;
		bool hit = false, b1 = false, b2 = false;
		for (int a = 0; a < 2 && !hit; a++)
		for (int b = 0; b < 2 && !hit; b++)
		for (int c = 0; c < 2 && !hit; c++)
		{
			const int ii = i * 2 + a, jj = j * 2 + b, kk = k * 2 + c;
			const size_t lin = (size_t)ii + (size_t)jj * sx * 2 + (size_t)kk * 4 * sxy;
			if (vbit(p.v.data(), lin, ii) && --need == 0)
			{
				hit = true;
				if (p.color) { b1 = vbit(p.c1.data(), lin, ii); b2 = vbit(p.c2.data(), lin, ii); }
			}
		}
		if (!hit) continue;
		const size_t lin = (size_t)i + (size_t)j * sx + (size_t)k * sxy;
		const uint8_t bit = (uint8_t)(1u << (i & 7));
		o.v[lin >> 3] |= bit;
		if (p.color)
		{
			if (b1) o.c1[lin >> 3] |= bit;
			if (b2) o.c2[lin >> 3] |= bit;
		}
	}
}

// Attribute of a surface voxel: brightness from the y component of the (negated-x/z)
// neighbourhood centroid direction, modulated by a position hash for materials 0 and 3,
// OR-ed with the two material bits (Rle4.cpp:122-163).
static inline uint16_t surface_attr(int mx, int my, int mz, int bit1, int bit2, int i, int k, int scale)
{
	float cx = (float)-mx, cy = (float)my, cz = (float)-mz;
	const float square = cx * cx + cy * cy + cz * cz;
	if (square <= 0.00001f) { cy = 0.0f; }
	else { const float len = 1.0f / sqrtf(square); cy *= len; }
	int ny = (int)(float)(127.0f * cy + 128.0f);
	if (ny > 255) ny = 255;
	if (ny < 0) ny = 0;
	const int mat = (bit1 | bit2) >> 8;
	if (mat == 0) ny = (ny * (((((i * scale) ^ (k * scale)) >> 4) & 15) + 15)) >> 5;
	if (mat == 3) ny = (ny * (((((i * scale) ^ (k * scale)) >> 3) & 15) + 25)) >> 5;
	if (ny > 255) ny = 255;
	if (ny < 0) ny = 0;
	return (uint16_t)(ny | bit1 | bit2);
}

// Emit one closed run (air gap `skip` followed by `solid` stored voxels), splitting at the
// 10-bit / 6-bit field limits the way the reference does (Rle4.cpp:166-182).
static inline void flush_run(std::vector<uint16_t>& out, int& skip, int& solid)
{
	while (skip > 1023) { out.push_back(1023); skip -= 1023; }
	while (solid > 63) { out.push_back((uint16_t)(63 * 1024 + (skip & 1023))); solid -= 63; skip = 0; }
	out.push_back((uint16_t)((solid & 63) * 1024 + (skip & 1023)));
	solid = 0;
	skip = 0;
}

// Compress the z-slice k of a bit volume into `out` (columns x = 0..sx-1 of that slice).
// `col_ofs[i]` receives the offset of column i relative to the start of `out`.
static void compress_slice(const BitVol& t, int k, int scale, std::vector<uint16_t>& out, std::vector<uint32_t>& col_ofs)
{
	const int sx = t.sx, sy = t.sy, sz = t.sz;
	const size_t sxy = (size_t)sx * sy;
	const uint8_t* mem = t.v.data();
	std::vector<uint16_t> tex;
	col_ofs.resize(sx);
	for (int i = 0; i < sx; i++)
	{
		col_ofs[i] = (uint32_t)out.size();
		const size_t head = out.size();
		out.push_back(0);
		out.push_back(0);
		tex.clear();
		int skip = 0, solid = 0;
		for (int j = 0; j <= sy; j++)
		{
			bool store = false;
			const size_t lin = (size_t)i + (size_t)j * sx + (size_t)k * sxy;
			if (j < sy && vbit(mem, lin, i))
			{
				int cnt = 0, mx = 0, my = 0, mz = 0;
				for (int a = -1; a < 2; a++)
				{
					const int x = i + a;
					if (x < 0 || x >= sx) continue;
					for (int b = -1; b < 2; b++)
					{
						const int y = j + b;
						if (y < 0 || y >= sy) continue;
						for (int c = -1; c < 2; c++)
						{
							const int z = k + c;
							if (z < 0 || z >= sz) continue;
							if (vbit(mem, (size_t)x + (size_t)y * sx + (size_t)z * sxy, x)) { mx += a; my += b; mz += c; cnt++; }
						}
					}
				}
				if (cnt < 27)
				{
					int bit1 = 0, bit2 = 0;
					if (t.color)
					{
						bit1 = vbit(t.c1.data(), lin, i) ? 256 : 0;
						bit2 = vbit(t.c2.data(), lin, i) ? 512 : 0;
					}
					tex.push_back(surface_attr(mx, my, mz, bit1, bit2, i, k, scale));
					store = true;
				}
			}
			if (solid > 0 && !store) flush_run(out, skip, solid);
			if (store) solid++; else skip++;
		}
		out[head] = (uint16_t)(out.size() - (head + 2));
		out[head + 1] = (uint16_t)tex.size();
		out.insert(out.end(), tex.begin(), tex.end());
	}
}

// Fast interior test for wide volumes: a voxel is interior when its whole 3x3x3
// neighbourhood is set; only non-interior set voxels are stored.  For sx%64==0 the
// neighbourhood AND is done 64 voxels at a time; the per-voxel path above then runs the
// 27-tap loop for surface voxels only.  Produces the same stream as compress_slice.
static void compress_slice_wide(const BitVol& t, int k, int scale, std::vector<uint16_t>& out, std::vector<uint32_t>& col_ofs)
{
	const int sx = t.sx, sy = t.sy, sz = t.sz;
	const int wx = sx / 64;
	const size_t sxy = (size_t)sx * sy;
	const uint8_t* mem = t.v.data();
	// interior[j*wx + w] : bit x set when voxel (x,j,k) has all 27 neighbours set
	std::vector<uint64_t> solidw((size_t)sy * wx), interior((size_t)sy * wx, 0);
	auto row = [&](int y, int z, int w) -> uint64_t {
		if (y < 0 || y >= sy || z < 0 || z >= sz) return 0;
		uint64_t r;
		memcpy(&r, mem + (((size_t)y * sx + (size_t)z * sxy) >> 3) + (size_t)w * 8, 8);
		return r;
	};
	for (int j = 0; j < sy; j++)
	for (int w = 0; w < wx; w++)
	{
		solidw[(size_t)j * wx + w] = row(j, k, w);
		uint64_t acc = ~0ull;
		for (int b = -1; b < 2 && acc; b++)
		for (int c = -1; c < 2 && acc; c++)
		{
			const uint64_t r = row(j + b, k + c, w);
			const uint64_t lo = (w > 0) ? row(j + b, k + c, w - 1) >> 63 : 0;
			const uint64_t hi = (w + 1 < wx) ? row(j + b, k + c, w + 1) << 63 : 0;
			acc &= r & ((r << 1) | lo) & ((r >> 1) | hi);
		}
		interior[(size_t)j * wx + w] = acc;
	}
	std::vector<uint16_t> tex;
	col_ofs.resize(sx);
	for (int i = 0; i < sx; i++)
	{
		col_ofs[i] = (uint32_t)out.size();
		const size_t head = out.size();
		out.push_back(0);
		out.push_back(0);
		tex.clear();
		int skip = 0, solid = 0;
		const int w = i >> 6;
		const uint64_t m = 1ull << (i & 63);
		for (int j = 0; j <= sy; j++)
		{
			bool store = false;
			if (j < sy && (solidw[(size_t)j * wx + w] & m) && !(interior[(size_t)j * wx + w] & m))
			{
				int mx = 0, my = 0, mz = 0;
				for (int a = -1; a < 2; a++)
				{
					const int x = i + a;
					if (x < 0 || x >= sx) continue;
					for (int b = -1; b < 2; b++)
					{
						const int y = j + b;
						if (y < 0 || y >= sy) continue;
						for (int c = -1; c < 2; c++)
						{
							const int z = k + c;
							if (z < 0 || z >= sz) continue;
							if (vbit(mem, (size_t)x + (size_t)y * sx + (size_t)z * sxy, x)) { mx += a; my += b; mz += c; }
						}
					}
				}
				int bit1 = 0, bit2 = 0;
				if (t.color)
				{
					const size_t lin = (size_t)i + (size_t)j * sx + (size_t)k * sxy;
					bit1 = vbit(t.c1.data(), lin, i) ? 256 : 0;
					bit2 = vbit(t.c2.data(), lin, i) ? 512 : 0;
				}
				tex.push_back(surface_attr(mx, my, mz, bit1, bit2, i, k, scale));
				store = true;
			}
			if (solid > 0 && !store) flush_run(out, skip, solid);
			if (store) solid++; else skip++;
		}
		out[head] = (uint16_t)(out.size() - (head + 2));
		out[head + 1] = (uint16_t)tex.size();
		out.insert(out.end(), tex.begin(), tex.end());
	}
}

static int compress_level(const BitVol& t, int mip, Level& lv)
{
	lv.sx = t.sx; lv.sy = t.sy; lv.sz = t.sz;
	const int scale = 1 << mip;
	std::vector<std::vector<uint16_t>> part(t.sz);
	std::vector<std::vector<uint32_t>> pofs(t.sz);
	const bool wide = (t.sx % 64 == 0);
	#pragma omp parallel for schedule(dynamic, 1)
	for (int k = 0; k < t.sz; k++)
	{
		if (wide) compress_slice_wide(t, k, scale, part[k], pofs[k]);
		else compress_slice(t, k, scale, part[k], pofs[k]);
	}
	uint64_t total = 0;
	for (int k = 0; k < t.sz; k++) total += part[k].size();
	if (total > 0xffffffffull) { set_error("level %d: slab stream exceeds 32-bit offsets", mip); return RLERC_ERR_FORMAT; }
	lv.slabs.resize(total);
	lv.map.assign((size_t)t.sx * t.sz * 2, 0);
	uint64_t base = 0;
	for (int k = 0; k < t.sz; k++)
	{
		if (!part[k].empty()) memcpy(lv.slabs.data() + base, part[k].data(), part[k].size() * 2);
		base += part[k].size();
		std::vector<uint16_t>().swap(part[k]);
	}
	return build_pointer_map(lv);
}

} // namespace rlerc

using namespace rlerc;

extern "C" {

const char* rlerc_last_error(void) { return g_err; }
const char* rlerc_version(void) { return "rlerc 0.1 (sm_100a)"; }

int rlerc_set_host_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
	return omp_get_max_threads();
#else
	(void)n;
	return 1;
#endif
}

int rlerc_scene_nummaps(const rlerc_scene* s) { return s ? (int)s->levels.size() : RLERC_ERR_ARG; }

void rlerc_scene_free(rlerc_scene* s) { delete s; }

int rlerc_scene_level(const rlerc_scene* s, int m, rlerc_map4* out, uint64_t* slabs_size64)
{
	if (!s || !out || m < 0 || m >= (int)s->levels.size()) { set_error("rlerc_scene_level: bad argument"); return RLERC_ERR_ARG; }
	const Level& lv = s->levels[m];
	out->sx = lv.sx; out->sy = lv.sy; out->sz = lv.sz;
	out->slabs_size = (int32_t)std::min<uint64_t>(lv.slabs.size(), 0x7fffffffull);
	out->map = const_cast<uint32_t*>(lv.map.data());
	out->slabs = const_cast<uint16_t*>(lv.slabs.data());
	if (slabs_size64) *slabs_size64 = lv.slabs.size();
	return RLERC_OK;
}

int rlerc_scene_load(const char* path, rlerc_scene** out)
{
	if (!path || !out) { set_error("rlerc_scene_load: null argument"); return RLERC_ERR_ARG; }
	*out = nullptr;
	FILE* f = fopen(path, "rb");
	if (!f) { set_error("cannot open %s", path); return RLERC_ERR_IO; }
	int32_t nummaps = 0;
	if (fread(&nummaps, 4, 1, f) != 1 || nummaps < 1 || nummaps > RLERC_MAX_MAPS)
	{
		fclose(f);
		set_error("%s: bad level count %d", path, nummaps);
		return RLERC_ERR_FORMAT;
	}
	// the header's sizes are claims: compare them with what is left of the file before allocating anything
	fseeko(f, 0, SEEK_END);
	const uint64_t file_bytes = (uint64_t)ftello(f);
	fseeko(f, 4, SEEK_SET);
	rlerc_scene* s = nullptr;
	try {
	s = new rlerc_scene();
	s->levels.resize(nummaps);
	for (int m = 0; m < nummaps; m++)
	{
		Level& lv = s->levels[m];
		int32_t hdr[4];
		if (fread(hdr, 4, 4, f) != 4) { fclose(f); delete s; set_error("%s: truncated header of level %d", path, m); return RLERC_ERR_IO; }
		lv.sx = hdr[0]; lv.sy = hdr[1]; lv.sz = hdr[2];
		// slabs_size is a signed int32 in the reference (Rle4.cpp:237); sizes in
		// [2^31, 2^32) written by our own tiler are read back as unsigned.
		const uint64_t n = (uint32_t)hdr[3];
		if (lv.sx <= 0 || lv.sy <= 0 || lv.sz <= 0 || n < 2ull * lv.sx * lv.sz)
		{
			fclose(f); delete s;
			set_error("%s: level %d has implausible header %d x %d x %d, %llu slabs", path, m, hdr[0], hdr[1], hdr[2], (unsigned long long)n);
			return RLERC_ERR_FORMAT;
		}
		const uint64_t at = (uint64_t)ftello(f);
		if (at + 2 * n > file_bytes)
		{
			fclose(f); delete s;
			set_error("%s: truncated slabs of level %d (header claims %llu, the file has %llu bytes left)", path, m, (unsigned long long)(2 * n), (unsigned long long)(file_bytes - at));
			return RLERC_ERR_IO;
		}
		lv.slabs.resize(n);
		if (fread(lv.slabs.data(), 2, n, f) != n) { fclose(f); delete s; set_error("%s: truncated slabs of level %d", path, m); return RLERC_ERR_IO; }
		int rc = build_pointer_map(lv);
		if (rc == RLERC_OK) rc = validate_columns(lv);
		if (rc != RLERC_OK) { fclose(f); delete s; return rc; }
	}
	} catch (const std::bad_alloc&) {
		fclose(f); delete s;
		set_error("%s: out of host memory", path);
		return RLERC_ERR_NOMEM;
	}
	fclose(f);
	*out = s;
	return RLERC_OK;
}

int rlerc_scene_save(const rlerc_scene* s, const char* path)
{
	if (!s || !path || s->levels.empty()) { set_error("rlerc_scene_save: bad argument"); return RLERC_ERR_ARG; }
	for (const Level& lv : s->levels)
		if (lv.slabs.size() > 0xffffffffull) { set_error("level too large for the .rle4 header"); return RLERC_ERR_FORMAT; }
	FILE* f = fopen(path, "wb");
	if (!f) { set_error("cannot create %s", path); return RLERC_ERR_IO; }
	int32_t nummaps = (int32_t)s->levels.size();
	bool ok = fwrite(&nummaps, 4, 1, f) == 1;
	for (const Level& lv : s->levels)
	{
		int32_t hdr[4] = { lv.sx, lv.sy, lv.sz, (int32_t)(uint32_t)lv.slabs.size() };
		ok = ok && fwrite(hdr, 4, 4, f) == 4;
		ok = ok && fwrite(lv.slabs.data(), 2, lv.slabs.size(), f) == lv.slabs.size();
	}
	fclose(f);
	if (!ok) { set_error("short write to %s", path); return RLERC_ERR_IO; }
	return RLERC_OK;
}

int rlerc_scene_from_maps(const rlerc_map4* maps, int nummaps, rlerc_scene** out)
{
	if (!maps || !out || nummaps < 1 || nummaps > RLERC_MAX_MAPS) { set_error("rlerc_scene_from_maps: bad argument"); return RLERC_ERR_ARG; }
	*out = nullptr;
	rlerc_scene* s = nullptr;
	try {
	s = new rlerc_scene();
	s->levels.resize(nummaps);
	for (int m = 0; m < nummaps; m++)
	{
		const rlerc_map4& src = maps[m];
		Level& lv = s->levels[m];
		if (!src.map || !src.slabs || src.sx <= 0 || src.sz <= 0 || src.slabs_size <= 0) { delete s; set_error("level %d is empty", m); return RLERC_ERR_ARG; }
		lv.sx = src.sx; lv.sy = src.sy; lv.sz = src.sz;
		lv.map.assign(src.map, src.map + (size_t)src.sx * src.sz * 2);
		lv.slabs.assign(src.slabs, src.slabs + (size_t)(uint32_t)src.slabs_size);
		const int rc = validate_columns(lv);                             // the caller's pointer map is a claim too
		if (rc != RLERC_OK) { delete s; return rc; }
	}
	} catch (const std::bad_alloc&) {
		delete s;
		set_error("rlerc_scene_from_maps: out of host memory");
		return RLERC_ERR_NOMEM;
	}
	*out = s;
	return RLERC_OK;
}

int rlerc_scene_compress(const uint8_t* voxel, const uint8_t* col1, const uint8_t* col2, int sx, int sy, int sz, rlerc_scene** out)
{
	if (!voxel || !out || sx < 1 || sy < 1 || sz < 1 || ((col1 == nullptr) != (col2 == nullptr)))
	{
		set_error("rlerc_scene_compress: bad argument");
		return RLERC_ERR_ARG;
	}
	*out = nullptr;
	// pyramid: keep halving while every dimension is >= 4 (at most 15 reductions)
	std::vector<BitVol> t(1);
	t[0].init(sx, sy, sz, col1 != nullptr);
	const size_t bytes = ((size_t)sx * sy * sz + 7) / 8;
	memcpy(t[0].v.data(), voxel, bytes);
	if (col1) { memcpy(t[0].c1.data(), col1, bytes); memcpy(t[0].c2.data(), col2, bytes); }
	for (int m = 0; m < 15; m++)
	{
		if (t[m].sx < 4 || t[m].sy < 4 || t[m].sz < 4) break;
		t.emplace_back();
		build_mip(t[m], t[m + 1]);
	}
	rlerc_scene* s = new rlerc_scene();
	s->levels.resize(t.size());
	for (size_t m = 0; m < t.size(); m++)
	{
		int rc = compress_level(t[m], (int)m, s->levels[m]);
		if (rc != RLERC_OK) { delete s; return rc; }
		BitVol().v.swap(t[m].v);
	}
	*out = s;
	return RLERC_OK;
}

int rlerc_scene_tile(const rlerc_scene* src, int nx, int nz, rlerc_scene** out)
{
	if (!src || !out || nx < 1 || nz < 1) { set_error("rlerc_scene_tile: bad argument"); return RLERC_ERR_ARG; }
	*out = nullptr;
	rlerc_scene* s = new rlerc_scene();
	s->levels.resize(src->levels.size());
	for (size_t m = 0; m < src->levels.size(); m++)
	{
		const Level& a = src->levels[m];
		Level& b = s->levels[m];
		const uint64_t total = (uint64_t)a.slabs.size() * nx * nz;
		if (total > 0xffffffffull) { delete s; set_error("tiled level %zu needs %llu slabs (> 32-bit offsets)", m, (unsigned long long)total); return RLERC_ERR_FORMAT; }
		b.sx = a.sx * nx; b.sy = a.sy; b.sz = a.sz * nz;
		b.slabs.resize(total);
		b.map.resize((size_t)b.sx * b.sz * 2);
		// Row z of the source level occupies one contiguous slab range; a tiled row is that
		// range repeated nx times, and the tiled level is the row sequence repeated nz times.
		std::vector<uint64_t> row_begin(a.sz + 1);
		for (int z = 0; z < a.sz; z++) row_begin[z] = a.map[(size_t)z * a.sx * 2];
		row_begin[a.sz] = a.slabs.size();
		// the stream may carry unreferenced tail words; rows are taken up to the next row start
		std::vector<uint64_t> dst_row((size_t)b.sz + 1);
		uint64_t ofs = 0;
		for (int tz = 0; tz < nz; tz++)
		for (int z = 0; z < a.sz; z++)
		{
			dst_row[(size_t)tz * a.sz + z] = ofs;
			ofs += (row_begin[z + 1] - row_begin[z]) * nx;
		}
		b.slabs.resize(ofs);
		#pragma omp parallel for schedule(static)
		for (int zz = 0; zz < b.sz; zz++)
		{
			const int z = zz % a.sz;
			const uint64_t len = row_begin[z + 1] - row_begin[z];
			for (int tx = 0; tx < nx; tx++)
				memcpy(b.slabs.data() + dst_row[zz] + len * tx, a.slabs.data() + row_begin[z], len * 2);
		}
		// pointer map: every tiled row starts at a known offset, so rows are scanned in parallel
		const uint64_t n = b.slabs.size();
		const uint16_t* sl = b.slabs.data();
		#pragma omp parallel for schedule(static)
		for (int zz = 0; zz < b.sz; zz++)
		{
			uint64_t o = dst_row[zz];
			for (int x = 0; x < b.sx; x++)
			{
				const uint32_t n_runs = sl[o], n_vox = sl[o + 1];
				const uint32_t first = (o + 2 < n) ? sl[o + 2] : 0u;
				const size_t dc = ((size_t)zz * b.sx + x) * 2;
				b.map[dc] = (uint32_t)o;
				b.map[dc + 1] = n_runs + (first << 16);
				o += (uint64_t)n_runs + n_vox + 2;
			}
		}
	}
	*out = s;
	return RLERC_OK;
}

void rlerc_frame_config_default(int width, int height, rlerc_frame_config* out)
{
	if (!out) return;
	out->width = width;
	out->height = height;
	out->render_size = width;
	out->rays_casted = width * 4;
	out->rays_casted_res = width * 4;
	out->z_far = 80000;
	out->mip_distance = width;
	out->border = (1.0f - (float)height / (float)width) * 0.5f;
	out->flags = 0;
}

int rlerc_frame_setup(const float pos[3], const float rot[3], const rlerc_frame_config* cfg, rlerc_raymap* out)
{
	if (!pos || !rot || !cfg || !out) { set_error("rlerc_frame_setup: null argument"); return RLERC_ERR_ARG; }
	if (cfg->rays_casted_res < 4 || cfg->rays_casted_res % 4) { set_error("rays_casted_res must be a positive multiple of 4"); return RLERC_ERR_ARG; }
	rlerc::get_ray_map(pos, rot, cfg->border, cfg->rays_casted_res, out);
	return RLERC_OK;
}

} // extern "C"
