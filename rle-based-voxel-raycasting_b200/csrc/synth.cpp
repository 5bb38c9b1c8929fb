// Procedural bit volumes for the synthetic benchmark scenes (Imrodh.rle4 is
// not in the reference checkout; DESIGN.md §6 describes the substitutes).  Pure integer
// hashing + fixed-point value noise so that a (kind, size, seed) triple names exactly one
// volume on every machine.  Tooling for benchmarks/tests; not on the frame loop.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "rlerc_internal.h"

namespace {

inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
	uint32_t h = seed * 0x9E3779B1u ^ (x * 0x85EBCA77u) ^ (y * 0xC2B2AE3Du) ^ (z * 0x27D4EB2Fu);
	h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
	return h;
}

// 2-D value noise on a lattice of `cell` voxels, periodic in `period` lattice points,
// 16.16 fixed point in [0, 65535]; smoothstep-free bilinear blend keeps it integer.
inline uint32_t vnoise2(int x, int z, int cell, int period, uint32_t seed)
{
	const int cx = x / cell, cz = z / cell;
	const uint32_t fx = (uint32_t)(x % cell) * 65536u / (uint32_t)cell;
	const uint32_t fz = (uint32_t)(z % cell) * 65536u / (uint32_t)cell;
	const int x0 = cx % period, x1 = (cx + 1) % period, z0 = cz % period, z1 = (cz + 1) % period;
	const uint64_t a = hash3(x0, 0, z0, seed) & 0xffff, b = hash3(x1, 0, z0, seed) & 0xffff;
	const uint64_t c = hash3(x0, 0, z1, seed) & 0xffff, d = hash3(x1, 0, z1, seed) & 0xffff;
	const uint64_t top = a * (65536 - fx) + b * fx, bot = c * (65536 - fx) + d * fx;
	return (uint32_t)((top * (65536 - fz) + bot * fz) >> 32);
}

inline uint32_t vnoise3(int x, int y, int z, int cell, int period, uint32_t seed)
{
	const int cx = x / cell, cy = y / cell, cz = z / cell;
	const uint64_t fx = (uint32_t)(x % cell) * 256u / (uint32_t)cell;
	const uint64_t fy = (uint32_t)(y % cell) * 256u / (uint32_t)cell;
	const uint64_t fz = (uint32_t)(z % cell) * 256u / (uint32_t)cell;
	uint64_t acc = 0;
	for (int k = 0; k < 8; k++)
	{
		const int ix = (cx + (k & 1)) % period, iy = cy + ((k >> 1) & 1), iz = (cz + (k >> 2)) % period;
		const uint64_t w = ((k & 1) ? fx : 256 - fx) * ((k & 2) ? fy : 256 - fy) * ((k & 4) ? fz : 256 - fz);
		acc += w * (hash3(ix, iy, iz, seed) & 0xffff);
	}
	return (uint32_t)(acc >> 24);
}

inline void setbit(uint8_t* m, size_t lin, int x) { m[lin >> 3] |= (uint8_t)(1u << (x & 7)); }

} // namespace

extern "C" int rlerc_synth_volume(int kind, int sx, int sy, int sz, uint32_t seed, uint8_t* voxel, uint8_t* col1, uint8_t* col2)
{
	if (!voxel || sx < 8 || sy < 8 || sz < 8 || (sx % 8) || (kind != 0 && kind != 1) || ((col1 == nullptr) != (col2 == nullptr)))
	{
		rlerc::set_error("rlerc_synth_volume: bad argument");
		return RLERC_ERR_ARG;
	}
	const size_t sxy = (size_t)sx * sy;
	const size_t bytes = sxy * sz / 8;
	memset(voxel, 0, bytes);
	if (col1) { memset(col1, 0, bytes); memset(col2, 0, bytes); }
	// World y points down: y = 0 is the top of the volume, the ground fills y >= h(x,z).
	const int cell0 = sx / 4 > 8 ? sx / 4 : 8;
	#pragma omp parallel for schedule(dynamic, 4)
	for (int z = 0; z < sz; z++)
	for (int x = 0; x < sx; x++)
	{
		// 4-octave heightfield in [sy/4, 3*sy/4], periodic so that the infinite tiling is seamless
		uint64_t acc = 0, wsum = 0;
		for (int o = 0; o < 4; o++)
		{
			const int cell = (cell0 >> o) > 2 ? (cell0 >> o) : 2;
			const int period = sx / cell > 1 ? sx / cell : 1;
			const uint32_t w = 8u >> o;
			acc += (uint64_t)vnoise2(x, z, cell, period, seed + 17 * o) * w;
			wsum += 65535ull * w;
		}
		const int h = sy / 4 + (int)(acc * (uint64_t)(sy / 2) / wsum);
		// sparse pillars / boulders above the ground: one candidate per 64x64 block
		const int bx = x / 64, bz = z / 64;
		const uint32_t hb = hash3(bx, 7, bz, seed ^ 0xabcdu);
		const int px = bx * 64 + 16 + (int)(hb & 31), pz = bz * 64 + 16 + (int)((hb >> 5) & 31);
		const int pr = 4 + (int)((hb >> 10) & 7);
		const int ptop = h - 8 - (int)((hb >> 13) & 63);
		const bool pillar = ((hb >> 20) & 3) == 0 && (x - px) * (x - px) + (z - pz) * (z - pz) <= pr * pr;
		for (int y = 0; y < sy; y++)
		{
			bool solid = y >= h;
			int mat = 1;
			if (solid)
			{
				if (kind == 1)
				{
					// worst-case short-run band: alternate solid/air every voxel for 128 voxels under the surface
					if (y - h < 128 && y - h >= 2) solid = ((y - h) & 1) == 0;
				}
				else
				{
					// caves: 3-D noise threshold below a crust of 6 voxels
					if (y > h + 6 && y < sy - 4 && vnoise3(x, y, z, 32, sx / 32 > 1 ? sx / 32 : 1, seed + 99) > 47000u) solid = false;
					if (y > h + 3) mat = 0;
				}
			}
			else if (pillar && y >= ptop && y < h) { solid = true; mat = 3; }
			if (!solid) continue;
			const size_t lin = (size_t)x + (size_t)y * sx + (size_t)z * sxy;
			setbit(voxel, lin, x);
			if (col1)
			{
				if (mat & 1) setbit(col1, lin, x);
				if (mat & 2) setbit(col2, lin, x);
			}
		}
	}
	return RLERC_OK;
}


// ---- direct-to-RLE scene: BASELINE config 4 at its full size ---------------------------------------------------------
// A 16384 x 1024 x 16384 bit volume would be 32 GiB, so this scene is written column by column straight into the
// .rle4 layout ([n_runs][n_vox][runs][attr16 x n_vox], run = solid << 10 | skip, split as R/src/Rle4.cpp:166-182),
// every mip level from the same functions of the level-0 coordinates:
//   * periodic 4-octave heightfield in [sy/4, 3*sy/4], a crust of 2 surface voxels;
//   * one column in `band_every`: the worst-case short-run band, 32 runs of one solid voxel / one air voxel under
//     the crust (halved per mip level) — as many runs per column as the per-level ushort budget of the format allows
//     on average (slabs_size is an int32 count of ushorts, R/src/Rle4.cpp:237);
//   * a cave floor two voxels thick under a noise threshold;
//   * attribute = hash(x, y, z) & 0x3ff.
namespace {

inline int height_at(int X, int Z, int sx, int sy, uint32_t seed)
{
	const int cell0 = sx / 4 > 8 ? sx / 4 : 8;
	uint64_t acc = 0, wsum = 0;
	for (int o = 0; o < 4; o++)
	{
		const int cell = (cell0 >> (2 * o)) > 2 ? (cell0 >> (2 * o)) : 2;
		const int period = sx / cell > 1 ? sx / cell : 1;
		const uint32_t w = 8u >> o;
		acc += (uint64_t)vnoise2(X, Z, cell, period, seed + 17 * o) * w;
		wsum += 65535ull * w;
	}
	return sy / 4 + (int)(acc * (uint64_t)(sy / 2) / wsum);
}

inline void push_run(std::vector<uint16_t>& runs, int& skip, int solid)
{
	while (skip > 1023) { runs.push_back(1023); skip -= 1023; }                        // Rle4.cpp:168-172
	while (solid > 63) { runs.push_back((uint16_t)(63 * 1024 + (skip & 1023))); solid -= 63; skip = 0; }
	runs.push_back((uint16_t)((solid & 63) * 1024 + (skip & 1023)));
	skip = 0;
}

// one column of level m at level coordinates (x, z), appended to `out`
void gen_column(int m, int x, int z, int sx0, int sy0, uint32_t seed, int band_every, std::vector<uint16_t>& out,
                std::vector<uint16_t>& runs, std::vector<uint16_t>& attrs)
{
	const int X = x << m, Z = z << m, sym = sy0 >> m;
	const int hm = height_at(X, Z, sx0, sy0, seed) >> m;
	runs.clear(); attrs.clear();
	int y = 0, skip = 0;                       // y: next voxel not yet encoded
	auto solid_span = [&](int a, int b)       // [a, b) solid, a >= y
	{
		if (b > sym) b = sym;
		if (a < y) a = y;
		if (b <= a) return;
		skip += a - y;
		push_run(runs, skip, b - a);
		for (int v = a; v < b; v++) attrs.push_back((uint16_t)(hash3((uint32_t)X, (uint32_t)(v << m), (uint32_t)Z, seed) & 0x3ffu));
		y = b;
	};
	solid_span(hm, hm + 2);
	if (band_every > 0 && (hash3((uint32_t)X, 3u, (uint32_t)Z, seed ^ 0x5bd1u) % (uint32_t)band_every) == 0)
	{
		const int nb = 32 >> m;
		for (int k = 0; k < nb; k++) solid_span(hm + 3 + 2 * k, hm + 4 + 2 * k);
	}
	if (vnoise2(X, Z, 64, sx0 / 64 > 1 ? sx0 / 64 : 1, seed + 99) > 40000u)
	{
		const int cf = hm + (96 >> m) + 4;
		solid_span(cf, cf + 2);
	}
	out.push_back((uint16_t)runs.size());
	out.push_back((uint16_t)attrs.size());
	out.insert(out.end(), runs.begin(), runs.end());
	out.insert(out.end(), attrs.begin(), attrs.end());
}

} // namespace

extern "C" int rlerc_synth_rle(int sx, int sy, int sz, uint32_t seed, int band_every, rlerc_scene** out)
{
	auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
	if (!out || !pow2(sx) || !pow2(sz) || !pow2(sy) || sx < 64 || sz < 64 || sy < 64 || sy > 1024 || band_every < 0)
	{
		rlerc::set_error("rlerc_synth_rle: sizes must be powers of two, sx, sz >= 64, 64 <= sy <= 1024");
		return RLERC_ERR_ARG;
	}
	*out = nullptr;
	rlerc_scene* sc = new rlerc_scene();
	for (int m = 0; m < 16 && (sx >> m) >= 8 && (sz >> m) >= 8 && (sy >> m) >= 8; m++)
	{
		const int lx = sx >> m, ly = sy >> m, lz = sz >> m;
		std::vector<std::vector<uint16_t>> rows((size_t)lz);
		#pragma omp parallel
		{
			std::vector<uint16_t> runs, attrs;
			#pragma omp for schedule(dynamic, 8)
			for (int z = 0; z < lz; z++)
			{
				std::vector<uint16_t>& r = rows[(size_t)z];
				r.reserve((size_t)lx * 8);
				for (int x = 0; x < lx; x++) gen_column(m, x, z, sx, sy, seed, band_every, r, runs, attrs);
			}
		}
		uint64_t total = 0;
		for (const auto& r : rows) total += r.size();
		if (total >= 0x7fffffffull)
		{
			delete sc;
			rlerc::set_error("rlerc_synth_rle: level %d needs %llu ushorts, the format holds 2^31-1 per level (R/src/Rle4.cpp:237)", m, (unsigned long long)total);
			return RLERC_ERR_ARG;
		}
		sc->levels.emplace_back();
		rlerc::Level& lv = sc->levels.back();
		lv.sx = lx; lv.sy = ly; lv.sz = lz;
		lv.slabs.resize((size_t)total);
		std::vector<uint64_t> base((size_t)lz + 1, 0);
		for (int z = 0; z < lz; z++) base[(size_t)z + 1] = base[(size_t)z] + rows[(size_t)z].size();
		#pragma omp parallel for schedule(dynamic, 8)
		for (int z = 0; z < lz; z++)
		{
			memcpy(lv.slabs.data() + base[(size_t)z], rows[(size_t)z].data(), rows[(size_t)z].size() * sizeof(uint16_t));
			std::vector<uint16_t>().swap(rows[(size_t)z]);
		}
		const int rc = rlerc::build_pointer_map(lv);
		if (rc) { delete sc; return rc; }
	}
	*out = sc;
	return RLERC_OK;
}
