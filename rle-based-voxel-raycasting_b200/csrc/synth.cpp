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
