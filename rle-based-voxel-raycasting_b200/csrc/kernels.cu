// Hand-written CUDA for sm_100a: the two kernels of the frame loop.
//
//   k_traverse  one GROUP of G lanes per ray plane (G = 1..32): replaces
//               cudaRender + Render::render_line (R/src/Cuda_Main.cu:150-181,
//               R/src/Cuda_Render.h:96-737).
//   k_unwarp    one thread per 4 output pixels: replaces GLSL pass 1
//               (R/bin/shader/colorize_buddha_soft.frag, uniforms R/src/main.cpp:578-603).
//
// Arithmetic contract: this file is compiled with -fmad=false, IEEE
// division and square root, no flush-to-zero; every float expression keeps the operand
// order and the float/int typing of the reference statement it cites; float->int goes
// through f2i() which reproduces x86 cvttss2si (the oracle is the reference compiled for
// the host).  Results are bit-identical for every G; G only changes how the work of one
// ray plane is spread over lanes.
//
// How one ray plane is spread over G lanes (nothing in the reference corresponds to this):
//   1. DDA batch: all G lanes step the (inherently serial, float-accumulating) 2-D DDA
//      together for G cell crossings; lane s keeps the state of crossing s.
//   2. Each lane turns its crossing into a column address + the projected cell geometry
//      and, if the top-clip test passes, issues its pointer-map gather right away: up to G
//      independent 8-byte gathers in flight per ray instead of one.
//   3. Crossings are consumed front to back: ballot finds the next lane whose column
//      survives the top-clip test under the CURRENT floating-horizon bounds.
//   4. A column's runs are loaded G at a time (coalesced 2-byte loads), their y extents and
//      attribute offsets come from a shuffle prefix sum, all G runs are projected in
//      parallel, and ballots find the runs that actually change state (draw / break) in
//      order, so the order-dependent horizon + bitmask updates happen exactly as in the
//      serial loop.
//   5. The pixels of a run are shaded G at a time, stores coalesced along the ray row.
#include <stdint.h>
#include <limits.h>
#include "kernels.cuh"

namespace rlerc {

#define RLERC_BLOCK 128
#define RLERC_SKY 0xff8844u

// x86 cvttss2si: truncation, and the "integer indefinite" 0x80000000 for NaN / out of range.
__device__ __forceinline__ int f2i(float f)
{
	return (fabsf(f) < 2147483648.0f) ? __float2int_rz(f) : INT_MIN;
}

// while (y < b && bit(y)) ++y, starting at y = a (Cuda_Render.h:577,639): the first row in
// [a, b) whose occlusion bit is clear, else b; a itself when the range is empty.
__device__ __forceinline__ int first_clear(const uint32_t* ymask, int a, int b)
{
	if (a >= b) return a;
	int y = a;
	while (y < b)
	{
		const int w = y >> 5;
		const uint32_t inv = ~ymask[w] & (0xffffffffu << (y & 31));
		if (inv) { y = (w << 5) + __ffs(inv) - 1; break; }
		y = (w + 1) << 5;
	}
	return y < b ? y : b;
}

// Cuda_Render.h:39-52
__device__ __forceinline__ float line_scale(float ix, float iy, float cx, float cy, float clip_max, float clip_min)
{
	float sx = 1, sy = 1;
	if (cx > 1) sx = (1 - ix) / (cx - ix);
	if (cx < 0) sx = ix / (ix - cx);
	if (cy > clip_max) sy = (clip_max - iy) / (cy - iy);
	if (cy < clip_min) sy = (-clip_min + iy) / (iy - cy);
	return (sx < sy) ? sx : sy;
}

// i-th ray plane of this launch: contiguous from ray_begin, or, for interleaved multi-GPU
// slices, the i-th ray r >= ray_begin... with (r / slice_block) % slice_n == slice_rank.
__device__ __forceinline__ int owned_ray(const TraverseParams& P, int i)
{
	if (P.slice_n <= 1) return P.ray_begin + i;
	const int blk = i / P.slice_block, off = i - blk * P.slice_block;
	return (blk * P.slice_n + P.slice_rank) * P.slice_block + off;
}

// host: how many rays of [0, count) an interleaved slice owns
int owned_count(int count, int block, int n, int rank)
{
	if (n <= 1) return count;
	const int cyc = block * n;
	int owned = (count / cyc) * block;
	int rem = count % cyc - rank * block;
	if (rem > block) rem = block;
	if (rem > 0) owned += rem;
	return owned;
}

template <int G, bool IDS>
__global__ void __launch_bounds__(RLERC_BLOCK)
k_traverse(const __grid_constant__ TraverseParams P)
{
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int GPB = RLERC_BLOCK / G;                 // ray planes per block
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int gl = tid & (G - 1);                        // lane within the group
	const int grp = tid / G;
	const int gshift = lane & ~(G - 1);                  // warp lane of the group's lane 0
	const unsigned gbits = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
	const unsigned gmask = gbits << gshift;

	const int x = owned_ray(P, blockIdx.x * GPB + grp);
	if (x >= P.ray_end) return;

	// shared: [GPB][G] crossing records (2 x float4), then [GPB][mask_words] occlusion bits
	float4* rec = reinterpret_cast<float4*>(smem) + (size_t)grp * G * 2;
	uint32_t* ymask = smem + (size_t)GPB * G * 8 + (size_t)grp * P.mask_words;

	const int res_x = P.res_x, res_y = P.res_y;
	const float res_x2 = (float)(res_x / 2);             // Cuda_Render.h:107-108 (integer division)
	const float res_y2 = (float)(res_y / 2);
	uint32_t* row = P.warp + (size_t)x * res_y;

	// ---- A. ray set-up (Cuda_Render.h:128-172) -------------------------------------------
	float ray_x, ray_z, s2x, s2y, e2x, e2y;
	bool vertical;
	{
		const int r0 = P.res[0], r1 = P.res[1] + r0, r2 = P.res[2] + r1;
		int q = 0;
		if (x >= r2) q = 3; else if (x >= r1) q = 2; else if (x >= r0) q = 1;
		float qofs = (float)x;
		if (q >= 1) qofs -= (float)(q == 1 ? r0 : (q == 2 ? r1 : r2));
		const float a = qofs / (float)P.res[q];
		float p1x = P.vp[0], p1y = P.vp[1], p1z = P.vp[2];
		const float ax = P.p_no[q * 2][0], ay = P.p_no[q * 2][1], az = P.p_no[q * 2][2];
		float p2x = ax + (P.p_no[q * 2 + 1][0] - ax) * a;
		float p2y = ay + (P.p_no[q * 2 + 1][1] - ay) * a;
		float p2z = az + (P.p_no[q * 2 + 1][2] - az) * a;
		{	// ClipLine (Cuda_Render.h:54-65)
			float sc = line_scale(p1x, p1y, p2x, p2y, P.clip_max, P.clip_min);
			const float c2x = p1x + (p2x - p1x) * sc, c2y = p1y + (p2y - p1y) * sc, c2z = p1z + (p2z - p1z) * sc;
			sc = line_scale(p2x, p2y, p1x, p1y, P.clip_max, P.clip_min);
			const float c1x = p2x + (p1x - p2x) * sc, c1y = p2y + (p1y - p2y) * sc, c1z = p2z + (p1z - p2z) * sc;
			p1x = c1x; p1y = c1y; p1z = c1z;
			p2x = c2x; p2y = c2y; p2z = c2z;
		}
		const float a1x = p1x * 4.0f, a1y = p1y * 4.0f, a1z = p1z * 4.0f;
		const float a2x = p2x * 4.0f, a2y = p2y * 4.0f, a2z = p2z * 4.0f;
		// MatMul (Cuda_Render.h:67-73); only x and z of the sum survive delta.y = 0
		const float b1x = P.to3d[0][0] * a1x + P.to3d[1][0] * a1y + P.to3d[2][0] * a1z + P.to3d[3][0];
		const float b1z = P.to3d[0][2] * a1x + P.to3d[1][2] * a1y + P.to3d[2][2] * a1z + P.to3d[3][2];
		const float b2x = P.to3d[0][0] * a2x + P.to3d[1][0] * a2y + P.to3d[2][0] * a2z + P.to3d[3][0];
		const float b2z = P.to3d[0][2] * a2x + P.to3d[1][2] * a2y + P.to3d[2][2] * a2z + P.to3d[3][2];
		float dx = (b1x + b2x) * 0.5f - P.p4[0];
		float dz = (b1z + b2z) * 0.5f - P.p4[2];
		const float dy = 0.0f;
		// normalize = v * rsqrtf(dot), host fallback rsqrtf = 1.0f/sqrtf (R/inc/cutil_math.h:58-61,1184-1188)
		const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
		dx = dx * inv; dz = dz * inv;
		// vec3f_rot_y(viewrot.y) (Cuda_Render.h:75-80)
		ray_x = P.cos_my * dx + P.sin_my * dz;
		ray_z = P.cos_my * dz - P.sin_my * dx;
		s2x = p1x; s2y = p1y; e2x = p2x; e2y = p2y;
		vertical = (q < 2);
	}

	// ---- B. screen-space clip of the ray row (Cuda_Render.h:203-250) ----------------------
	const float sin_x = P.sin_x, cos_x = P.cos_x;
	bool reverse = false;
	if (vertical) { if (ray_z <= 0) reverse = true; }
	else
	{
		if (ray_x <= 0) { if (sin_x > 0) reverse = true; }
		if (ray_x > 0) { if (sin_x < 0) reverse = true; }
	}
	float rx2mr = reverse ? -res_x2 : res_x2;
	if (vertical) rx2mr = -rx2mr;

	int ycmin, ycmax;
	{
		const int p_add = reverse ? 1 : -2;
		int q1x = f2i((float)res_x * s2x) + p_add;
		int q1y = f2i((float)res_y * s2y) + p_add;
		int q2x = f2i((float)res_x * e2x) - p_add;
		int q2y = f2i((float)res_y * e2y) - p_add;
		if (q1x < 0) q1x = 0; if (q1x >= res_x) q1x = res_x - 1;
		if (q1y < 0) q1y = 0; if (q1y >= res_y) q1y = res_y - 1;
		if (q2x < 0) q2x = 0; if (q2x >= res_x) q2x = res_x - 1;
		if (q2y < 0) q2y = 0; if (q2y >= res_y) q2y = res_y - 1;
		bool skip_ray = (q1y == q2y);                   // Cuda_Render.h:226
		ycmin = res_x - 1 - q1x;
		ycmax = res_x - 1 - q2x;
		if (vertical) { ycmin = res_y - 1 - q1y; ycmax = res_y - 1 - q2y; }
		if (reverse) { ycmin = res_y - 1 - ycmin; ycmax = res_y - 1 - ycmax; }
		if (ycmin > ycmax) { const int t = ycmin; ycmin = ycmax; ycmax = t; }
		if (ycmin >= ycmax) skip_ray = true;            // Cuda_Render.h:250
		// Texels outside the clip range are never written by the reference (stale from the previous
		// frame) yet the unwarp samples a few of them; they are defined as 0 here so
		// that a frame does not depend on history.
		if (skip_ray) { ycmin = res_y; ycmax = res_y - 1; }
		for (int y = gl; y < ycmin; y += G) row[y] = 0;
		for (int y = ycmax + 1 + gl; y < res_y; y += G) row[y] = 0;
		if (skip_ray) return;
	}
	const int ymin0 = ycmin, ymax0 = ycmax;

	// ---- C. occlusion mask clear; the sky sentinel is written at the end to the pixels
	//         that stayed open (same final row as clear-then-overwrite, Cuda_Render.h:255-264)
	for (int w = gl; w < P.mask_words; w += G) ymask[w] = 0;
	__syncwarp(gmask);

	// ---- D. DDA initialisation (Cuda_Render.h:270-305) ------------------------------------
	const float vpx = P.viewpos[0], mountain = P.viewpos[1], vpz = P.viewpos[2];
	int fixx = -1, fixz = -1;
	float g0x, g0y, g1x, g1y, i0x, i0y, i1x, i1y, gd0, gd1, d0, d1;
	{
		const float drx = ray_x * P.cos_y + ray_z * P.sin_y;
		const float dry = ray_x * P.sin_y - ray_z * P.cos_y;
		float fx = vpx - (float)f2i(vpx);
		float fy = vpz - (float)f2i(vpz);
		float sgx = -1, sgy = -1;
		if (drx >= 0) { fixx = 0; sgx = 1; fx = 1 - fx; }
		if (dry >= 0) { fixz = 0; sgy = 1; fy = 1 - fy; }
		g0y = dry / fabsf(drx); g0x = sgx;
		g1x = drx / fabsf(dry); g1y = sgy;
		i0x = g0x * fx; i0y = g0y * fx;
		i1x = g1x * fy; i1y = g1y * fy;
		gd0 = sqrtf(g0x * g0x + g0y * g0y);
		gd1 = sqrtf(g1x * g1x + g1y * g1y);
		d0 = sqrtf(i0x * i0x + i0y * i0y);
		d1 = sqrtf(i1x * i1x + i1y * i1y);
	}
	float posx = 0, posy = 0, dist_now = 0;
	int index = 0;
	int mip = 0;
	int gridx = P.level[0].sx, gridz = P.level[0].sz;
	const float pz_add = sin_x;                                  // pos3d_z_add
	const float py_add = (vertical ? cos_x : 0.0f) * rx2mr;      // pos3d_y_add
	float z = 0, dz = 1.0f;
	int mapswitch = P.mapswitch0;
	float yms = mountain;                                        // y_map_switch
	const float zfar = (float)P.z_far;
	const int last_map = P.nummaps - 1;

	unsigned long long c_total = 0, c_proc = 0, c_vox = 0, c_rend = 0, c_pix = 0, c_cols = 0, c_iter = 0, c_cols1 = 0, c_steps = 0;

	// ---- E. main loop ------------------------------------------------------------------------
	bool alive = true;
	while (alive)
	{
		if (ycmin >= ycmax) break;                               // Cuda_Render.h:370

		// E1-E3 for G consecutive cell crossings; lane s keeps crossing s
		int nvalid = G;
		#pragma unroll 4
		for (int s = 0; s < G; s++)
		{
			while (z > (float)mapswitch || yms > 512.0f)         // Cuda_Render.h:343-365
			{
				yms = yms * 0.5f;
				if (mip < last_map) { mip++; gridx >>= 1; gridz >>= 1; }
				g0x *= 2; g0y *= 2; g1x *= 2; g1y *= 2;
				gd0 *= 2; gd1 *= 2;
				mapswitch *= 2;
				dz *= 2;
			}
			z += dz;
			if (z > zfar) { nvalid = s; break; }                 // Cuda_Render.h:367
			const float db = dist_now, pbx = posx, pby = posy;
			const int ib = index;
			if (d1 < d0)                                         // Cuda_Render.h:398-414
			{
				dist_now = d1; index = 1; d1 += gd1;
				posx = i1x; posy = i1y;
				i1x += g1x; i1y += g1y;
			}
			else
			{
				dist_now = d0; index = 0;
				posx = i0x; posy = i0y;
				d0 += gd0;
				i0x += g0x; i0y += g0y;
			}
			if (gl == s)
			{
				rec[gl * 2 + 0] = make_float4(db, dist_now, pbx, pby);
				rec[gl * 2 + 1] = make_float4(__int_as_float(ib), __int_as_float(mip), 0.0f, 0.0f);
			}
		}
		if (IDS) c_steps += nvalid;

		// E4-E6 per lane: column address, projected cell, top-clip test, map gather
		float pz = 0, py = 0, czz = 0, cyy = 0;
		int cmip = 0, cidx = 0;
		uint2 ent = make_uint2(0, 0);
		bool loaded = false;
		if (gl < nvalid)
		{
			const float4 ra = rec[gl * 2 + 0], rb = rec[gl * 2 + 1];
			const float db = ra.x, dn = ra.y;
			const int ib = __float_as_int(rb.x);
			cmip = __float_as_int(rb.y);
			const int fix_x = (1 - ib) * fixx, fix_z = ib * fixz;    // Cuda_Render.h:418-419
			const float ddelta = dn - db;
			const float vsx = ray_x * db, vsz = ray_z * db;
			const int voxel_x = f2i(vpx + ra.z) + fix_x;
			const int voxel_z = f2i(vpz + ra.w) + fix_z;
			const int gx = P.level[cmip].sx, gz = P.level[cmip].sz;
			const int vx = (voxel_x >> cmip) & (gx - 1);             // Cuda_Render.h:441-442
			const int vz = (voxel_z >> cmip) & (gz - 1);
			cidx = vx + vz * gx;
			const float corx = ray_x * ddelta, corz = ray_z * ddelta;
			pz = cos_x * vsz + sin_x * mountain;                    // Cuda_Render.h:459-464
			py = vertical ? (cos_x * mountain - sin_x * vsz) : vsx;
			py *= rx2mr;
			czz = cos_x * corz;                                     // Cuda_Render.h:483-486
			cyy = vertical ? (-sin_x * corz) : corx;
			cyy *= rx2mr;
			if (!(pz * res_y2 + py <= pz * (float)ycmin))
			{
				ent = __ldg(P.level[cmip].map + cidx);
				loaded = true;
			}
		}

		// consume crossings front to back
		unsigned todo = (nvalid >= 32) ? 0xffffffffu : ((1u << nvalid) - 1u);
		while (true)
		{
			if (ycmin >= ycmax) { alive = false; break; }
			const bool pass = ((todo >> gl) & 1u) && !(pz * res_y2 + py <= pz * (float)ycmin);   // Cuda_Render.h:467
			const unsigned pb = (__ballot_sync(gmask, pass) >> gshift) & gbits;
			if (!pb) break;
			const int L = __ffs(pb) - 1;
			todo &= ~((2u << L) - 1u);
			if (gl == L && !loaded) ent = __ldg(P.level[cmip].map + cidx);
			// broadcast the column of lane L
			const float cpz = __shfl_sync(gmask, pz, L, G);
			const float cpy = __shfl_sync(gmask, py, L, G);
			const float cczz = __shfl_sync(gmask, czz, L, G);
			const float ccyy = __shfl_sync(gmask, cyy, L, G);
			const int m = __shfl_sync(gmask, cmip, L, G);
			const unsigned e0 = __shfl_sync(gmask, ent.x, L, G);
			const unsigned e1 = __shfl_sync(gmask, ent.y, L, G);
			const int colid = IDS ? __shfl_sync(gmask, cidx, L, G) : 0;

			const int slen = (int)(e1 & 0xffffu);
			const unsigned first = e1 >> 16;
			const uint16_t* runs = P.level[m].slabs + 2 + (size_t)e0;     // Cuda_Render.h:498-499
			const uint16_t* send = runs + slen;
			if (IDS) { c_cols++; c_total += slen; if (slen) c_cols1++; }

			// E7. the column's runs, G at a time
			int base_len = 0, base_tex = 0;
			bool done = false;
			for (int c = 0; c < slen && !done; c += G)
			{
				const int j = c + gl;
				unsigned r = 0;
				if (j < slen) r = (j == 0) ? first : (unsigned)__ldg(runs + j);
				const int skip = (int)(r & 1023u), solid = (int)(r >> 10);
				// inclusive prefix sum of {skip+solid, solid}, packed 16:16
				const unsigned v = ((unsigned)(skip + solid) << 16) | (unsigned)solid;
				unsigned inc = v;
				#pragma unroll
				for (int d = 1; d < G; d <<= 1)
				{
					const unsigned t = __shfl_up_sync(gmask, inc, d, G);
					if (gl >= d) inc += t;
				}
				const unsigned exc = inc - v;
				const int top = (base_len + (int)(exc >> 16) + skip) << m;     // sti_general_sti_skip
				const int bot = top + (solid << m);                             // sti_general
				const int texture = base_tex + (int)(exc & 0xffffu);
				const int texn = texture + solid;                               // tex
				const unsigned tot = __shfl_sync(gmask, inc, G - 1, G);
				base_len += (int)(tot >> 16);
				base_tex += (int)(tot & 0xffffu);

				// project top and bottom of every run (Cuda_Render.h:529-560)
				bool v1 = false, v2 = false;
				int sy2 = 0, sy1 = 0;
				if (solid > 0)
				{
					const float ft = (float)top, fb = (float)bot;
					float zz1 = cpz, yy1 = cpy;
					if (mountain + ft >= 0) { zz1 += cczz; yy1 += ccyy; }
					const float z1 = zz1 + pz_add * ft;
					if (!(z1 <= 0))
					{
						v1 = true;
						const float y1 = yy1 + py_add * ft;
						sy2 = f2i(res_y2 + y1 / z1);
						float zz2 = cpz, yy2 = cpy;
						if (mountain + fb < 0) { zz2 += cczz; yy2 += ccyy; }
						const float z2 = zz2 + pz_add * fb;
						if (!(z2 <= 0))
						{
							v2 = true;
							const float y2 = yy2 + py_add * fb;
							sy1 = f2i(res_y2 + y2 / z2 - 1);
						}
					}
				}

				// resolve the runs in order: only "break" and "draw" events change state
				unsigned rem = (slen - c >= 32) ? 0xffffffffu : ((1u << (slen - c)) - 1u);
				rem &= gbits;
				int limit = G - 1;          // last run of this chunk the serial loop reaches
				while (true)
				{
					const bool inrem = (rem >> gl) & 1u;
					const bool brk = inrem && v1 && (sy2 <= ycmin);
					const bool drw = inrem && v1 && v2 && !brk && !(sy1 >= ycmax);
					const unsigned bb = (__ballot_sync(gmask, brk) >> gshift) & gbits;
					const unsigned bd = (__ballot_sync(gmask, drw) >> gshift) & gbits;
					if (!(bb | bd)) break;
					const int fb = bb ? (__ffs(bb) - 1) : 64;
					const int fd = bd ? (__ffs(bd) - 1) : 64;
					if (fb < fd) { done = true; limit = fb; break; }   // Cuda_Render.h:543
					rem &= ~((2u << fd) - 1u);

					int s2 = __shfl_sync(gmask, sy2, fd, G);
					int s1 = __shfl_sync(gmask, sy1, fd, G);
					const int rtop = __shfl_sync(gmask, top, fd, G);
					const int rbot = __shfl_sync(gmask, bot, fd, G);
					const int rtex = __shfl_sync(gmask, texture, fd, G);
					const int rtexn = __shfl_sync(gmask, texn, fd, G);

					// floating horizon (Cuda_Render.h:564-580)
					if (s2 >= ycmax) { s2 = ycmax; ycmax = s1; }
					if (s1 <= ycmin)
					{
						s1 = ycmin;
						ycmin = s2;
						ycmin = first_clear(ymask, ycmin, ycmax);
					}
					int y = first_clear(ymask, s1, s2);                // Cuda_Render.h:639-640
					if (y >= s2) continue;

					// interpolants (Cuda_Render.h:645-680)
					const float ft = (float)rtop, fb2 = (float)rbot;
					const float z1r = cpz + pz_add * ft, y1r = cpy + py_add * ft;
					const float z2r = cpz + pz_add * fb2, y2r = cpy + py_add * fb2;
					const float s2r = res_y2 + y1r / z1r;
					const float s1r = res_y2 + y2r / z2r;
					const float u1z = (float)rtexn / z2r;
					float u2dz = (float)rtex / z1r - u1z;
					const float onez1 = 1.0f / z2r;
					float onedz2 = 1.0f / z1r - onez1;
					u2dz /= s2r - s1r;
					onedz2 /= s2r - s1r;
					if (IDS) c_rend++;
					const float mult = (float)(y + 1) - s1r;
					float uz = u1z + u2dz * mult;
					float onez = onez1 + onedz2 * mult;
					const int tex_hi = rtexn - 1;                      // int(float(tex-1.0))

					// E8. pixels, G at a time (Cuda_Render.h:687-733)
					const int n = s2 - y;
					for (int c0 = 0; c0 < n; c0 += G)
					{
						const int steps = (n - c0 < G) ? (n - c0) : G;
						float muz = uz, monez = onez;
						for (int t = 0; t < steps; t++)
						{
							if (gl == t) { muz = uz; monez = onez; }
							uz += u2dz; onez += onedz2;
						}
						const int yy = y + c0 + gl;
						bool wr = false;
						if (gl < steps && !((ymask[yy >> 5] >> (yy & 31)) & 1u))
						{
							wr = true;
							int ui = f2i(muz / monez);
							ui = (ui > rtex) ? ui : rtex;
							ui = (ui < tex_hi) ? ui : tex_hi;
							const unsigned real_z = (unsigned)f2i(1.0f / monez) & 0xfffeu;
							const unsigned color16 = __ldg(send + ui);
							row[yy] = color16 + (real_z << 16);
							if (IDS)
							{
								uint32_t* id = P.ids + ((size_t)x * res_y + yy) * 2;
								id[0] = (uint32_t)colid;
								id[1] = ((uint32_t)m << 16) | (uint32_t)ui;
							}
						}
						const unsigned wb = (__ballot_sync(gmask, wr) >> gshift) & gbits;
						if (wb)
						{
							if (IDS) c_pix += __popc(wb);
							if (gl == 0)
							{
								const int y0 = y + c0, wi = y0 >> 5, sh = y0 & 31;
								ymask[wi] |= wb << sh;
								if (sh && (wb >> (32 - sh))) ymask[wi + 1] |= wb >> (32 - sh);
							}
						}
						__syncwarp(gmask);
					}
				}
				if (IDS)
				{
					const unsigned reach = (limit >= 31) ? 0xffffffffu : ((2u << limit) - 1u);
					const bool cnt = (j < slen) && ((reach >> gl) & 1u) && solid > 0;
					const unsigned cb = (__ballot_sync(gmask, cnt) >> gshift) & gbits;
					c_proc += __popc(cb);
					// voxels_processed: sum of solid<<mip over the counted runs
					int vsum = cnt ? (solid << m) : 0;
					#pragma unroll
					for (int d = G / 2; d > 0; d >>= 1) vsum += __shfl_xor_sync(gmask, vsum, d, G);
					c_vox += vsum;
					if (done) c_iter += c + limit + 1;
				}
			}
			if (IDS && !done) c_iter += slen;
		}
		if (nvalid < G) alive = false;
	}

	// sky sentinel on every pixel of the clip range that no run covered
	for (int y = ymin0 + gl; y <= ymax0; y += G)
		if (!((ymask[y >> 5] >> (y & 31)) & 1u)) row[y] = RLERC_SKY;

	if (IDS && gl == 0 && P.counters)
	{
		atomicAdd(P.counters + 0, c_total);
		atomicAdd(P.counters + 1, c_proc);
		atomicAdd(P.counters + 2, c_vox);
		atomicAdd(P.counters + 3, c_rend);
		atomicAdd(P.counters + 4, c_pix);
		atomicAdd(P.counters + 5, c_cols);
		atomicAdd(P.counters + 6, c_iter);
		atomicAdd(P.counters + 7, c_cols1);
		atomicAdd(P.counters + 8, (unsigned long long)(ymax0 - ymin0 + 1));
		atomicAdd(P.counters + 9, c_steps);
	}
}


// ------------------------------------------------------------------------------------------
// k_traverse_w: one WARP per ray plane, lane <-> COLUMN.  Same results as k_traverse<G>.
//
// The serial kernel (and k_traverse<G>) pays one dependent memory round trip per visited
// column (pointer-map entry -> run words -> attribute word).  Here a batch of 32 cell
// crossings is turned into 32 columns at once:
//   1. DDA for 32 crossings (uniform, serial float recurrence; lane s keeps crossing s).
//   2. Every lane whose column may survive the top-clip test gathers its map entry, then the
//      next RW-1 run words (the first rides in the entry), and projects those runs to screen
//      rows.  None of this depends on the occlusion state, so it is 32-wide and all loads
//      of a batch are in flight together.
//   3. Only columns that would DRAW under the current floating-horizon bounds change state.
//      Each lane checks that for its own pre-projected runs; a ballot finds the first such
//      column in front-to-back order; every column in between is a provable no-op and is
//      skipped.  The state is then advanced by that column's owner lane alone, with exactly
//      the serial statement order, and the ballot repeats under the new state.
//   4. Short pixel spans are shaded by the owner lane (attribute gathers left in flight,
//      stores flushed at the end of the batch); long spans are shaded 32 pixels at a time
//      by the whole warp with coalesced stores.  Columns with more than RW runs fall back
//      to the lane <-> run scheme of k_traverse<32>.
#define RLERC_RW 4          // runs pre-projected per column
#define RLERC_PEND 4        // deferred pixel stores per lane
#define RLERC_COOP_MIN 12   // pixel spans at least this long are shaded by the whole warp

struct DrawJob {            // owner lane -> warp hand-off for a long pixel span (in shared memory)
	float cpz, cpy;
	int y, s2, rtop, rbot, rtex, rtexn;
	int m, colid;
	unsigned e0, slen;
};

template <bool IDS>
__global__ void __launch_bounds__(RLERC_BLOCK, 4)
k_traverse_w(const __grid_constant__ TraverseParams P)
{
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int G = 32;
	constexpr int WPB = RLERC_BLOCK / 32;
	const int gl = threadIdx.x & 31;
	const int wid = threadIdx.x >> 5;
	const unsigned FULL = 0xffffffffu;

	const int x = owned_ray(P, blockIdx.x * WPB + wid);
	if (x >= P.ray_end) return;

	// shared per warp: 33 crossing records (float4) | DrawJob (16 words) | occlusion bits
	const int per_warp = ((G + 1) * 4 + 16 + P.mask_words + 3) & ~3;   // keep the float4 records 16-byte aligned
	uint32_t* wbase = smem + (size_t)wid * per_warp;
	float4* rec = reinterpret_cast<float4*>(wbase);
	DrawJob* job = reinterpret_cast<DrawJob*>(wbase + (G + 1) * 4);
	uint32_t* ymask = wbase + (G + 1) * 4 + 16;

	const int res_x = P.res_x, res_y = P.res_y;
	const float res_x2 = (float)(res_x / 2);
	const float res_y2 = (float)(res_y / 2);
	uint32_t* row = P.warp + (size_t)x * res_y;

	// ---- A. ray set-up (Cuda_Render.h:128-172) -------------------------------------------
	float ray_x, ray_z, s2x, s2y, e2x, e2y;
	bool vertical;
	{
		const int r0 = P.res[0], r1 = P.res[1] + r0, r2 = P.res[2] + r1;
		int q = 0;
		if (x >= r2) q = 3; else if (x >= r1) q = 2; else if (x >= r0) q = 1;
		float qofs = (float)x;
		if (q >= 1) qofs -= (float)(q == 1 ? r0 : (q == 2 ? r1 : r2));
		const float a = qofs / (float)P.res[q];
		float p1x = P.vp[0], p1y = P.vp[1], p1z = P.vp[2];
		const float ax = P.p_no[q * 2][0], ay = P.p_no[q * 2][1], az = P.p_no[q * 2][2];
		float p2x = ax + (P.p_no[q * 2 + 1][0] - ax) * a;
		float p2y = ay + (P.p_no[q * 2 + 1][1] - ay) * a;
		float p2z = az + (P.p_no[q * 2 + 1][2] - az) * a;
		{
			float sc = line_scale(p1x, p1y, p2x, p2y, P.clip_max, P.clip_min);
			const float c2x = p1x + (p2x - p1x) * sc, c2y = p1y + (p2y - p1y) * sc, c2z = p1z + (p2z - p1z) * sc;
			sc = line_scale(p2x, p2y, p1x, p1y, P.clip_max, P.clip_min);
			const float c1x = p2x + (p1x - p2x) * sc, c1y = p2y + (p1y - p2y) * sc, c1z = p2z + (p1z - p2z) * sc;
			p1x = c1x; p1y = c1y; p1z = c1z;
			p2x = c2x; p2y = c2y; p2z = c2z;
		}
		const float a1x = p1x * 4.0f, a1y = p1y * 4.0f, a1z = p1z * 4.0f;
		const float a2x = p2x * 4.0f, a2y = p2y * 4.0f, a2z = p2z * 4.0f;
		const float b1x = P.to3d[0][0] * a1x + P.to3d[1][0] * a1y + P.to3d[2][0] * a1z + P.to3d[3][0];
		const float b1z = P.to3d[0][2] * a1x + P.to3d[1][2] * a1y + P.to3d[2][2] * a1z + P.to3d[3][2];
		const float b2x = P.to3d[0][0] * a2x + P.to3d[1][0] * a2y + P.to3d[2][0] * a2z + P.to3d[3][0];
		const float b2z = P.to3d[0][2] * a2x + P.to3d[1][2] * a2y + P.to3d[2][2] * a2z + P.to3d[3][2];
		float dx = (b1x + b2x) * 0.5f - P.p4[0];
		float dz = (b1z + b2z) * 0.5f - P.p4[2];
		const float dy = 0.0f;
		const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
		dx = dx * inv; dz = dz * inv;
		ray_x = P.cos_my * dx + P.sin_my * dz;
		ray_z = P.cos_my * dz - P.sin_my * dx;
		s2x = p1x; s2y = p1y; e2x = p2x; e2y = p2y;
		vertical = (q < 2);
	}

	// ---- B. screen-space clip (Cuda_Render.h:203-250) --------------------------------------
	const float sin_x = P.sin_x, cos_x = P.cos_x;
	bool reverse = false;
	if (vertical) { if (ray_z <= 0) reverse = true; }
	else
	{
		if (ray_x <= 0) { if (sin_x > 0) reverse = true; }
		if (ray_x > 0) { if (sin_x < 0) reverse = true; }
	}
	float rx2mr = reverse ? -res_x2 : res_x2;
	if (vertical) rx2mr = -rx2mr;

	int ycmin, ycmax;
	{
		const int p_add = reverse ? 1 : -2;
		int q1x = f2i((float)res_x * s2x) + p_add;
		int q1y = f2i((float)res_y * s2y) + p_add;
		int q2x = f2i((float)res_x * e2x) - p_add;
		int q2y = f2i((float)res_y * e2y) - p_add;
		if (q1x < 0) q1x = 0; if (q1x >= res_x) q1x = res_x - 1;
		if (q1y < 0) q1y = 0; if (q1y >= res_y) q1y = res_y - 1;
		if (q2x < 0) q2x = 0; if (q2x >= res_x) q2x = res_x - 1;
		if (q2y < 0) q2y = 0; if (q2y >= res_y) q2y = res_y - 1;
		bool skip_ray = (q1y == q2y);                   // Cuda_Render.h:226
		ycmin = res_x - 1 - q1x;
		ycmax = res_x - 1 - q2x;
		if (vertical) { ycmin = res_y - 1 - q1y; ycmax = res_y - 1 - q2y; }
		if (reverse) { ycmin = res_y - 1 - ycmin; ycmax = res_y - 1 - ycmax; }
		if (ycmin > ycmax) { const int t = ycmin; ycmin = ycmax; ycmax = t; }
		if (ycmin >= ycmax) skip_ray = true;            // Cuda_Render.h:250
		// texels outside the clip range: defined as 0 (see k_traverse / DESIGN.md §4)
		if (skip_ray) { ycmin = res_y; ycmax = res_y - 1; }
		for (int y = gl; y < ycmin; y += G) row[y] = 0;
		for (int y = ycmax + 1 + gl; y < res_y; y += G) row[y] = 0;
		if (skip_ray) return;
	}
	const int ymin0 = ycmin, ymax0 = ycmax;

	for (int w = gl; w < P.mask_words; w += G) ymask[w] = 0;
	__syncwarp();

	// ---- D. DDA initialisation (Cuda_Render.h:270-305) ------------------------------------
	const float vpx = P.viewpos[0], mountain = P.viewpos[1], vpz = P.viewpos[2];
	int fixx = -1, fixz = -1;
	float g0x, g0y, g1x, g1y, i0x, i0y, i1x, i1y, gd0, gd1, d0, d1;
	{
		const float drx = ray_x * P.cos_y + ray_z * P.sin_y;
		const float dry = ray_x * P.sin_y - ray_z * P.cos_y;
		float fx = vpx - (float)f2i(vpx);
		float fy = vpz - (float)f2i(vpz);
		float sgx = -1, sgy = -1;
		if (drx >= 0) { fixx = 0; sgx = 1; fx = 1 - fx; }
		if (dry >= 0) { fixz = 0; sgy = 1; fy = 1 - fy; }
		g0y = dry / fabsf(drx); g0x = sgx;
		g1x = drx / fabsf(dry); g1y = sgy;
		i0x = g0x * fx; i0y = g0y * fx;
		i1x = g1x * fy; i1y = g1y * fy;
		gd0 = sqrtf(g0x * g0x + g0y * g0y);
		gd1 = sqrtf(g1x * g1x + g1y * g1y);
		d0 = sqrtf(i0x * i0x + i0y * i0y);
		d1 = sqrtf(i1x * i1x + i1y * i1y);
	}
	float posx = 0, posy = 0, dist_now = 0;
	int index = 0;
	int mip = 0;
	const float pz_add = sin_x;
	const float py_add = (vertical ? cos_x : 0.0f) * rx2mr;
	int zi = 0, dzi = 1;                                        // z and dz (Cuda_Render.h:181,325), integer valued
	int mapswitch = P.mapswitch0;
	const int zfar_i = P.z_far;
	const int last_map = P.nummaps - 1;
	// The y_map_switch half of the LOD loop condition (Cuda_Render.h:343) can only be true on
	// the first crossing (it halves until <= 512 and never grows), where z = 0 < mapswitch.
	for (float yms = mountain; yms > 512.0f; yms = yms * 0.5f)
	{
		if (mip < last_map) mip++;
		g0x *= 2; g0y *= 2; g1x *= 2; g1y *= 2;
		gd0 *= 2; gd1 *= 2;
		mapswitch *= 2;
		dzi *= 2;
	}

	// per-lane statistics (IDS build only)
	unsigned long long c_total = 0, c_proc = 0, c_vox = 0, c_rend = 0, c_pix = 0, c_cols = 0, c_iter = 0, c_cols1 = 0, c_steps = 0;

	// deferred pixel stores of this lane
	int pend_n = 0;
	int pend_y[RLERC_PEND];
	unsigned pend_z[RLERC_PEND];
	unsigned pend_c[RLERC_PEND];
	#define RLERC_FLUSH_PENDING()                                                   \
		do {                                                                        \
			_Pragma("unroll")                                                       \
			for (int k_ = 0; k_ < RLERC_PEND; k_++)                                 \
				if (k_ < pend_n) row[pend_y[k_]] = pend_c[k_] + (pend_z[k_] << 16); \
			pend_n = 0;                                                             \
		} while (0)

	bool alive = true;
	while (alive)
	{
		if (ycmin >= ycmax) break;

		// ---- 1. DDA batch -------------------------------------------------------------------
		// z, dz, mapswitch and z_far are integer valued (z counts steps of 2^k), so the number of
		// crossings before the next LOD switch / before z_far is known up front and the inner
		// loop runs without per-step tests.  Slot s+1 receives the state after crossing s; slot
		// 0 carries the state before the batch.  All lanes store the same words (uniform address).
		int nvalid = G;
		rec[0] = make_float4(dist_now, posx, posy, __int_as_float(index));
		for (int s = 0; s < G;)
		{
			while (zi > mapswitch)                               // Cuda_Render.h:343-365
			{
				if (mip < last_map) mip++;
				g0x *= 2; g0y *= 2; g1x *= 2; g1y *= 2;
				gd0 *= 2; gd1 *= 2;
				mapswitch *= 2;
				dzi *= 2;
			}
			const int lod_free = (mapswitch - zi) / dzi + 1;     // crossings before z > mapswitch
			const int far_free = (zfar_i - zi) / dzi;            // crossings with z + dz <= z_far (Cuda_Render.h:366-367)
			if (far_free <= 0) { nvalid = s; break; }
			int n = G - s;
			n = n < lod_free ? n : lod_free;
			n = n < far_free ? n : far_free;
			const float mipbits = __int_as_float(mip << 1);
			float4* out = rec + s + 1;
			#define RLERC_DDA_STEP(K)                                                         \
				{                                                                             \
					const bool t1 = d1 < d0;                      /* Cuda_Render.h:398-414 */ \
					dist_now = t1 ? d1 : d0;                                                  \
					posx = t1 ? i1x : i0x;                                                    \
					posy = t1 ? i1y : i0y;                                                    \
					if (t1) { d1 += gd1; i1x += g1x; i1y += g1y; }                            \
					else    { d0 += gd0; i0x += g0x; i0y += g0y; }                            \
					out[K] = make_float4(dist_now, posx, posy, __int_as_float(__float_as_int(mipbits) | (t1 ? 1 : 0))); \
				}
			int j = 0;
			for (; j + 4 <= n; j += 4)
			{
				RLERC_DDA_STEP(j) RLERC_DDA_STEP(j + 1) RLERC_DDA_STEP(j + 2) RLERC_DDA_STEP(j + 3)
			}
			for (; j < n; j++) RLERC_DDA_STEP(j)
			#undef RLERC_DDA_STEP
			index = __float_as_int(out[n - 1].w) & 1;
			zi += n * dzi;
			s += n;
		}
		if (IDS && gl == 0) c_steps += nvalid;

		// ---- 2. per-lane column: address, projected cell, gathers, run pre-projection ---------
		float pz = 0, py = 0, czz = 0, cyy = 0;
		int cmip = 0, cidx = 0;
		unsigned e0 = 0, e1 = 0, rw12 = 0, rw3 = 0;
		int slen = 0;
		int nr = 0;                       // runs of this column that were pre-projected (<= RW)
		bool longcol = false;             // column has more than RW runs that may matter
		int sy1[RLERC_RW], sy2[RLERC_RW];
		unsigned flags = 0;               // bit r: v1 of run r, bit 8+r: v2 of run r
		#pragma unroll
		for (int r = 0; r < RLERC_RW; r++) { sy1[r] = 0; sy2[r] = 0; }
		bool have = false;
		if (gl < nvalid)
		{
			const float4 ra = rec[gl], rb = rec[gl + 1];        // state before / after crossing gl
			const float db = ra.x, dn = rb.x;
			const int ib = __float_as_int(ra.w) & 1;
			cmip = __float_as_int(rb.w) >> 1;
			const int fix_x = (1 - ib) * fixx, fix_z = ib * fixz;
			const float ddelta = dn - db;
			const float vsx = ray_x * db, vsz = ray_z * db;
			const int voxel_x = f2i(vpx + ra.y) + fix_x;
			const int voxel_z = f2i(vpz + ra.z) + fix_z;
			const int gx = P.level[cmip].sx, gz = P.level[cmip].sz;
			const int vx = (voxel_x >> cmip) & (gx - 1);
			const int vz = (voxel_z >> cmip) & (gz - 1);
			cidx = vx + vz * gx;
			const float corx = ray_x * ddelta, corz = ray_z * ddelta;
			pz = cos_x * vsz + sin_x * mountain;
			py = vertical ? (cos_x * mountain - sin_x * vsz) : vsx;
			py *= rx2mr;
			czz = cos_x * corz;
			cyy = vertical ? (-sin_x * corz) : corx;
			cyy *= rx2mr;
			// The horizon only rises while this batch is consumed.  For pz > 0 a column culled
			// now stays culled; for pz <= 0 (or NaN) the test can flip, so keep those.
			have = !(pz * res_y2 + py <= pz * (float)ycmin) || !(pz > 0);
		}
		if (have)
		{
			const uint2 ent = __ldg(P.level[cmip].map + cidx);
			e0 = ent.x; e1 = ent.y;
			slen = (int)(e1 & 0xffffu);
			const uint16_t* runs = P.level[cmip].slabs + 2 + (size_t)e0;
			unsigned r1 = 0, r2 = 0, r3 = 0;
			if (slen > 1) r1 = __ldg(runs + 1);
			if (slen > 2) r2 = __ldg(runs + 2);
			if (slen > 3) r3 = __ldg(runs + 3);
			rw12 = r1 | (r2 << 16); rw3 = r3;
			nr = slen < RLERC_RW ? slen : RLERC_RW;
			longcol = slen > RLERC_RW;
			int blen = 0;
			bool stop = false;
			#pragma unroll
			for (int r = 0; r < RLERC_RW; r++)
			{
				const unsigned rw = (r == 0) ? (e1 >> 16) : (r == 1) ? r1 : (r == 2) ? r2 : r3;
				const int skip = (int)(rw & 1023u), solid = (int)(rw >> 10);
				const int top = (blen + skip) << cmip;
				const int bot = top + (solid << cmip);
				blen += skip + solid;
				if (r < nr && !stop && solid > 0)
				{
					const float ft = (float)top, fb = (float)bot;
					float zz1 = pz, yy1 = py;
					if (mountain + ft >= 0) { zz1 += czz; yy1 += cyy; }
					const float z1 = zz1 + pz_add * ft;
					if (!(z1 <= 0))
					{
						flags |= 1u << r;
						const float y1 = yy1 + py_add * ft;
						sy2[r] = f2i(res_y2 + y1 / z1);
						if (sy2[r] <= ycmin)
						{
							// breaks now, hence under every later (higher) horizon: later runs are dead
							stop = true; nr = r + 1; longcol = false;
						}
						else
						{
							float zz2 = pz, yy2 = py;
							if (mountain + fb < 0) { zz2 += czz; yy2 += cyy; }
							const float z2 = zz2 + pz_add * fb;
							if (!(z2 <= 0))
							{
								flags |= 1u << (8 + r);
								const float y2 = yy2 + py_add * fb;
								sy1[r] = f2i(res_y2 + y2 / z2 - 1);
							}
						}
					}
				}
			}
		}

		// ---- 3. consume: only columns that draw under the current bounds change anything ------
		unsigned todo = (nvalid >= 32) ? FULL : ((1u << nvalid) - 1u);
		while (true)
		{
			if (ycmin >= ycmax) { alive = false; break; }
			const bool mine = (todo >> gl) & 1u;
			const bool pass = mine && !(pz * res_y2 + py <= pz * (float)ycmin);       // Cuda_Render.h:467
			// does my column draw (or is it too long to tell)?  also: where would its run loop stop
			bool ev = false;
			int my_iter = slen, my_proc = 0, my_vox = 0;
			if (pass)
			{
				bool brk = false;
				#pragma unroll
				for (int r = 0; r < RLERC_RW; r++)
				{
					if (r < nr && !ev && !brk)
					{
						if (IDS)
						{
							const unsigned rw = (r == 0) ? (e1 >> 16) : (r == 1) ? (rw12 & 0xffffu) : (r == 2) ? (rw12 >> 16) : rw3;
							if (rw >> 10) { my_proc++; my_vox += (int)(rw >> 10) << cmip; }
						}
						if ((flags >> r) & 1u)
						{
							if (sy2[r] <= ycmin) { brk = true; if (IDS) my_iter = r + 1; }
							else if (((flags >> (8 + r)) & 1u) && !(sy1[r] >= ycmax)) ev = true;
						}
					}
				}
				if (!ev && !brk && longcol) ev = true;
			}
			const unsigned pb = __ballot_sync(FULL, pass);
			const unsigned eb = __ballot_sync(FULL, ev);
			const int L = eb ? (__ffs(eb) - 1) : 32;
			if (IDS)
			{
				// every passing column up to (and including) the event column is "fetched" in the serial order
				const unsigned upto = (L >= 31) ? FULL : ((2u << L) - 1u);
				if (pass && ((upto >> gl) & 1u))
				{
					c_cols++; c_total += slen; if (slen) c_cols1++;
					if (gl != L) { c_iter += my_iter; c_proc += my_proc; c_vox += my_vox; }
				}
			}
			(void)pb;
			if (!eb) break;
			todo &= ~((2u << L) - 1u);

			const bool Llong = __shfl_sync(FULL, (int)(longcol && slen > RLERC_RW), L) != 0;
			if (!Llong)
			{
				// ---- owner lane advances the state through its column, serial statement order ----
				int rstart = 0;
				while (true)
				{
					int act = 0;      // 0 = column finished, 1 = long pixel span handed to the warp
					if (gl == L)
					{
						int blen = 0, btex = 0;
						bool fin = false;
						#pragma unroll
						for (int r = 0; r < RLERC_RW; r++)
						{
							const unsigned rw = (r == 0) ? (e1 >> 16) : (r == 1) ? (rw12 & 0xffffu) : (r == 2) ? (rw12 >> 16) : rw3;
							const int skip = (int)(rw & 1023u), solid = (int)(rw >> 10);
							const int top = (blen + skip) << cmip;
							const int bot = top + (solid << cmip);
							const int texture = btex, texn = btex + solid;
							blen += skip + solid; btex += solid;
							if (r >= rstart && r < nr && !fin && act == 0)
							{
								if (IDS) { c_iter++; if (solid > 0) { c_proc++; c_vox += solid << cmip; } }
								if ((flags >> r) & 1u)
								{
									if (sy2[r] <= ycmin) fin = true;                                   // Cuda_Render.h:543
									else if (((flags >> (8 + r)) & 1u) && !(sy1[r] >= ycmax))
									{
										int s2 = sy2[r], s1 = sy1[r];
										if (s2 >= ycmax) { s2 = ycmax; ycmax = s1; }                     // Cuda_Render.h:564-580
										if (s1 <= ycmin)
										{
											s1 = ycmin;
											ycmin = s2;
											ycmin = first_clear(ymask, ycmin, ycmax);
										}
										int y = first_clear(ymask, s1, s2);
										if (y < s2)
										{
											if (IDS) c_rend++;
											const int n = s2 - y;
											if (n >= RLERC_COOP_MIN)
											{
												job->cpz = pz; job->cpy = py;
												job->y = y; job->s2 = s2; job->rtop = top; job->rbot = bot;
												job->rtex = texture; job->rtexn = texn;
												job->m = cmip; job->colid = cidx; job->e0 = e0; job->slen = (unsigned)slen;
												act = 1; rstart = r + 1;
											}
											else
											{
												// interpolants (Cuda_Render.h:645-680)
												const float ft = (float)top, fb2 = (float)bot;
												const float z1r = pz + pz_add * ft, y1r = py + py_add * ft;
												const float z2r = pz + pz_add * fb2, y2r = py + py_add * fb2;
												const float s2r = res_y2 + y1r / z1r;
												const float s1r = res_y2 + y2r / z2r;
												const float u1z = (float)texn / z2r;
												float u2dz = (float)texture / z1r - u1z;
												const float onez1 = 1.0f / z2r;
												float onedz2 = 1.0f / z1r - onez1;
												u2dz /= s2r - s1r;
												onedz2 /= s2r - s1r;
												const float mult = (float)(y + 1) - s1r;
												float uz = u1z + u2dz * mult;
												float onez = onez1 + onedz2 * mult;
												const int tex_hi = texn - 1;
												const uint16_t* send = P.level[cmip].slabs + 2 + (size_t)e0 + slen;
												for (; y < s2; ++y, uz += u2dz, onez += onedz2)               // Cuda_Render.h:687-733
												{
													const uint32_t bit = 1u << (y & 31);
													if (ymask[y >> 5] & bit) continue;
													int ui = f2i(uz / onez);
													ui = (ui > texture) ? ui : texture;
													ui = (ui < tex_hi) ? ui : tex_hi;
													const unsigned real_z = (unsigned)f2i(1.0f / onez) & 0xfffeu;
													if (pend_n == RLERC_PEND) RLERC_FLUSH_PENDING();
													#pragma unroll
													for (int k = 0; k < RLERC_PEND; k++)
														if (k == pend_n) { pend_y[k] = y; pend_z[k] = real_z; pend_c[k] = __ldg(send + ui); }
													pend_n++;
													ymask[y >> 5] |= bit;
													if (IDS)
													{
														c_pix++;
														uint32_t* id = P.ids + ((size_t)x * res_y + y) * 2;
														id[0] = (uint32_t)cidx;
														id[1] = ((uint32_t)cmip << 16) | (uint32_t)ui;
													}
												}
											}
										}
									}
								}
							}
						}
					}
					act = __shfl_sync(FULL, act, L);
					__syncwarp();
					if (act == 0) break;
					rstart = __shfl_sync(FULL, rstart, L);

					// ---- long pixel span: whole warp, 32 pixels at a time --------------------------
					const DrawJob J = *job;
					const float ft = (float)J.rtop, fb2 = (float)J.rbot;
					const float z1r = J.cpz + pz_add * ft, y1r = J.cpy + py_add * ft;
					const float z2r = J.cpz + pz_add * fb2, y2r = J.cpy + py_add * fb2;
					const float s2r = res_y2 + y1r / z1r;
					const float s1r = res_y2 + y2r / z2r;
					const float u1z = (float)J.rtexn / z2r;
					float u2dz = (float)J.rtex / z1r - u1z;
					const float onez1 = 1.0f / z2r;
					float onedz2 = 1.0f / z1r - onez1;
					u2dz /= s2r - s1r;
					onedz2 /= s2r - s1r;
					const float mult = (float)(J.y + 1) - s1r;
					float uz = u1z + u2dz * mult;
					float onez = onez1 + onedz2 * mult;
					const int tex_hi = J.rtexn - 1;
					const uint16_t* send = P.level[J.m].slabs + 2 + (size_t)J.e0 + J.slen;
					const int n = J.s2 - J.y;
					for (int c0 = 0; c0 < n; c0 += G)
					{
						const int steps = (n - c0 < G) ? (n - c0) : G;
						float muz = uz, monez = onez;
						for (int t = 0; t < steps; t++)
						{
							if (gl == t) { muz = uz; monez = onez; }
							uz += u2dz; onez += onedz2;
						}
						const int yy = J.y + c0 + gl;
						bool wr = false;
						if (gl < steps && !((ymask[yy >> 5] >> (yy & 31)) & 1u))
						{
							wr = true;
							int ui = f2i(muz / monez);
							ui = (ui > J.rtex) ? ui : J.rtex;
							ui = (ui < tex_hi) ? ui : tex_hi;
							const unsigned real_z = (unsigned)f2i(1.0f / monez) & 0xfffeu;
							const unsigned color16 = __ldg(send + ui);
							row[yy] = color16 + (real_z << 16);
							if (IDS)
							{
								c_pix++;
								uint32_t* id = P.ids + ((size_t)x * res_y + yy) * 2;
								id[0] = (uint32_t)J.colid;
								id[1] = ((uint32_t)J.m << 16) | (uint32_t)ui;
							}
						}
						const unsigned wb = __ballot_sync(FULL, wr);
						if (wb && gl == 0)
						{
							const int y0 = J.y + c0, wi = y0 >> 5, sh = y0 & 31;
							ymask[wi] |= wb << sh;
							if (sh && (wb >> (32 - sh))) ymask[wi + 1] |= wb >> (32 - sh);
						}
						__syncwarp();
					}
				}
				ycmin = __shfl_sync(FULL, ycmin, L);
				ycmax = __shfl_sync(FULL, ycmax, L);
			}
			else
			{
				// ---- long column: lane <-> run, as in k_traverse<32> --------------------------------
				const float cpz = __shfl_sync(FULL, pz, L);
				const float cpy = __shfl_sync(FULL, py, L);
				const float cczz = __shfl_sync(FULL, czz, L);
				const float ccyy = __shfl_sync(FULL, cyy, L);
				const int m = __shfl_sync(FULL, cmip, L);
				const unsigned ce0 = __shfl_sync(FULL, e0, L);
				const unsigned ce1 = __shfl_sync(FULL, e1, L);
				const int colid = IDS ? __shfl_sync(FULL, cidx, L) : 0;
				const int cslen = (int)(ce1 & 0xffffu);
				const unsigned first = ce1 >> 16;
				const uint16_t* runs = P.level[m].slabs + 2 + (size_t)ce0;
				const uint16_t* send = runs + cslen;
				int base_len = 0, base_tex = 0;
				bool done = false;
				for (int c = 0; c < cslen && !done; c += G)
				{
					const int j = c + gl;
					unsigned r = 0;
					if (j < cslen) r = (j == 0) ? first : (unsigned)__ldg(runs + j);
					const int skip = (int)(r & 1023u), solid = (int)(r >> 10);
					const unsigned v = ((unsigned)(skip + solid) << 16) | (unsigned)solid;
					unsigned inc = v;
					#pragma unroll
					for (int d = 1; d < G; d <<= 1)
					{
						const unsigned t = __shfl_up_sync(FULL, inc, d);
						if (gl >= d) inc += t;
					}
					const unsigned exc = inc - v;
					const int top = (base_len + (int)(exc >> 16) + skip) << m;
					const int bot = top + (solid << m);
					const int texture = base_tex + (int)(exc & 0xffffu);
					const int texn = texture + solid;
					const unsigned tot = __shfl_sync(FULL, inc, G - 1);
					base_len += (int)(tot >> 16);
					base_tex += (int)(tot & 0xffffu);

					bool v1 = false, v2 = false;
					int ry2 = 0, ry1 = 0;
					if (solid > 0)
					{
						const float ft = (float)top, fb = (float)bot;
						float zz1 = cpz, yy1 = cpy;
						if (mountain + ft >= 0) { zz1 += cczz; yy1 += ccyy; }
						const float z1 = zz1 + pz_add * ft;
						if (!(z1 <= 0))
						{
							v1 = true;
							const float y1 = yy1 + py_add * ft;
							ry2 = f2i(res_y2 + y1 / z1);
							float zz2 = cpz, yy2 = cpy;
							if (mountain + fb < 0) { zz2 += cczz; yy2 += ccyy; }
							const float z2 = zz2 + pz_add * fb;
							if (!(z2 <= 0))
							{
								v2 = true;
								const float y2 = yy2 + py_add * fb;
								ry1 = f2i(res_y2 + y2 / z2 - 1);
							}
						}
					}
					unsigned rem = (cslen - c >= 32) ? FULL : ((1u << (cslen - c)) - 1u);
					int limit = G - 1;
					while (true)
					{
						const bool inrem = (rem >> gl) & 1u;
						const bool brk = inrem && v1 && (ry2 <= ycmin);
						const bool drw = inrem && v1 && v2 && !brk && !(ry1 >= ycmax);
						const unsigned bb = __ballot_sync(FULL, brk);
						const unsigned bd = __ballot_sync(FULL, drw);
						if (!(bb | bd)) break;
						const int fb = bb ? (__ffs(bb) - 1) : 64;
						const int fd = bd ? (__ffs(bd) - 1) : 64;
						if (fb < fd) { done = true; limit = fb; break; }
						rem &= ~((2u << fd) - 1u);
						int s2 = __shfl_sync(FULL, ry2, fd);
						int s1 = __shfl_sync(FULL, ry1, fd);
						const int rtop = __shfl_sync(FULL, top, fd);
						const int rbot = __shfl_sync(FULL, bot, fd);
						const int rtex = __shfl_sync(FULL, texture, fd);
						const int rtexn = __shfl_sync(FULL, texn, fd);
						if (s2 >= ycmax) { s2 = ycmax; ycmax = s1; }
						if (s1 <= ycmin)
						{
							s1 = ycmin;
							ycmin = s2;
							ycmin = first_clear(ymask, ycmin, ycmax);
						}
						int y = first_clear(ymask, s1, s2);
						if (y >= s2) continue;
						const float ft = (float)rtop, fb2 = (float)rbot;
						const float z1r = cpz + pz_add * ft, y1r = cpy + py_add * ft;
						const float z2r = cpz + pz_add * fb2, y2r = cpy + py_add * fb2;
						const float s2r = res_y2 + y1r / z1r;
						const float s1r = res_y2 + y2r / z2r;
						const float u1z = (float)rtexn / z2r;
						float u2dz = (float)rtex / z1r - u1z;
						const float onez1 = 1.0f / z2r;
						float onedz2 = 1.0f / z1r - onez1;
						u2dz /= s2r - s1r;
						onedz2 /= s2r - s1r;
						if (IDS && gl == 0) c_rend++;
						const float mult = (float)(y + 1) - s1r;
						float uz = u1z + u2dz * mult;
						float onez = onez1 + onedz2 * mult;
						const int tex_hi = rtexn - 1;
						const int n = s2 - y;
						for (int c0 = 0; c0 < n; c0 += G)
						{
							const int steps = (n - c0 < G) ? (n - c0) : G;
							float muz = uz, monez = onez;
							for (int t = 0; t < steps; t++)
							{
								if (gl == t) { muz = uz; monez = onez; }
								uz += u2dz; onez += onedz2;
							}
							const int yy = y + c0 + gl;
							bool wr = false;
							if (gl < steps && !((ymask[yy >> 5] >> (yy & 31)) & 1u))
							{
								wr = true;
								int ui = f2i(muz / monez);
								ui = (ui > rtex) ? ui : rtex;
								ui = (ui < tex_hi) ? ui : tex_hi;
								const unsigned real_z = (unsigned)f2i(1.0f / monez) & 0xfffeu;
								const unsigned color16 = __ldg(send + ui);
								row[yy] = color16 + (real_z << 16);
								if (IDS)
								{
									c_pix++;
									uint32_t* id = P.ids + ((size_t)x * res_y + yy) * 2;
									id[0] = (uint32_t)colid;
									id[1] = ((uint32_t)m << 16) | (uint32_t)ui;
								}
							}
							const unsigned wb = __ballot_sync(FULL, wr);
							if (wb && gl == 0)
							{
								const int y0 = y + c0, wi = y0 >> 5, sh = y0 & 31;
								ymask[wi] |= wb << sh;
								if (sh && (wb >> (32 - sh))) ymask[wi + 1] |= wb >> (32 - sh);
							}
							__syncwarp();
						}
					}
					if (IDS)
					{
						const unsigned reach = (limit >= 31) ? FULL : ((2u << limit) - 1u);
						if ((j < cslen) && ((reach >> gl) & 1u) && solid > 0) { c_proc++; c_vox += solid << m; }
						if (done && gl == 0) c_iter += c + limit + 1;
					}
				}
				if (IDS && !done && gl == 0) c_iter += cslen;
			}
		}
		RLERC_FLUSH_PENDING();
		if (nvalid < G) alive = false;
	}
	RLERC_FLUSH_PENDING();
	__syncwarp();

	for (int y = ymin0 + gl; y <= ymax0; y += G)
		if (!((ymask[y >> 5] >> (y & 31)) & 1u)) row[y] = RLERC_SKY;

	if (IDS && P.counters)
	{
		atomicAdd(P.counters + 0, c_total);
		atomicAdd(P.counters + 1, c_proc);
		atomicAdd(P.counters + 2, c_vox);
		atomicAdd(P.counters + 3, c_rend);
		atomicAdd(P.counters + 4, c_pix);
		atomicAdd(P.counters + 5, c_cols);
		atomicAdd(P.counters + 6, c_iter);
		atomicAdd(P.counters + 7, c_cols1);
		if (gl == 0) atomicAdd(P.counters + 8, (unsigned long long)(ymax0 - ymin0 + 1));
		atomicAdd(P.counters + 9, c_steps);
	}
	#undef RLERC_FLUSH_PENDING
}

template <bool IDS>
static void launch_traverse_w(const TraverseParams& p, cudaStream_t st)
{
	const int wpb = RLERC_BLOCK / 32;
	const int rays = (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
	if (rays <= 0) return;
	const int blocks = (rays + wpb - 1) / wpb;
	const size_t smem = (size_t)wpb * ((33 * 4 + 16 + p.mask_words + 3) & ~3) * sizeof(uint32_t);
	static size_t configured = 0;
	if (smem > configured)
	{
		cudaFuncSetAttribute(k_traverse_w<IDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = smem;
	}
	k_traverse_w<IDS><<<blocks, RLERC_BLOCK, smem, st>>>(p);
}

template <int G, bool IDS>
static void launch_traverse_t(const TraverseParams& p, cudaStream_t st)
{
	const int gpb = RLERC_BLOCK / G;
	const int rays = (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
	if (rays <= 0) return;
	const int blocks = (rays + gpb - 1) / gpb;
	const size_t smem = ((size_t)gpb * G * 8 + (size_t)gpb * p.mask_words) * sizeof(uint32_t);
	static size_t configured = 0;
	if (smem > configured)
	{
		cudaFuncSetAttribute(k_traverse<G, IDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = smem;
	}
	k_traverse<G, IDS><<<blocks, RLERC_BLOCK, smem, st>>>(p);
}

void launch_traverse(const TraverseParams& p, int lanes, bool ids, cudaStream_t st)
{
	if (lanes == 0) { ids ? launch_traverse_w<true>(p, st) : launch_traverse_w<false>(p, st); return; }
	switch (lanes)
	{
	case 1:  ids ? launch_traverse_t<1, true>(p, st)  : launch_traverse_t<1, false>(p, st); break;
	case 2:  ids ? launch_traverse_t<2, true>(p, st)  : launch_traverse_t<2, false>(p, st); break;
	case 4:  ids ? launch_traverse_t<4, true>(p, st)  : launch_traverse_t<4, false>(p, st); break;
	case 8:  ids ? launch_traverse_t<8, true>(p, st)  : launch_traverse_t<8, false>(p, st); break;
	case 16: ids ? launch_traverse_t<16, true>(p, st) : launch_traverse_t<16, false>(p, st); break;
	default: ids ? launch_traverse_t<32, true>(p, st) : launch_traverse_t<32, false>(p, st); break;
	}
}

// ------------------------------------------------------------------------------------------
// Unwarp + shade.  One thread produces four horizontally adjacent pixels and writes them
// with one 128-bit store.  Pixel (px, py) with GL origin bottom-left; output row 0 = top.
// The statement order follows colorize_buddha_soft.frag line by line (cites inline) so
// that the texel chosen is bit-identical to the CPU restatement in oracle/.
__device__ __forceinline__ float stepf(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }

__device__ __forceinline__ uint32_t quant8(float c)
{
	c = fminf(fmaxf(c, 0.0f), 1.0f);
	return (uint32_t)__float2int_rz(c * 255.0f + 0.5f);
}

__device__ __forceinline__ uint32_t unwarp_pixel(const UnwarpParams& P, int px, int pyg)
{
	const float RESX = (float)P.W, RESY = (float)P.H;
	const float fx = (float)px + 0.5f, fy = (float)pyg + 0.5f;      // gl_FragCoord
	const float scx = fx / RESX;                                      // frag:19-20
	const float scy = fy / RESY;
	const float border = (RESX - RESY) / (RESX * 2);                  // frag:22
	const float scx1 = scx - P.vanish_x;                              // frag:24-25
	const float scy1 = scy - P.vanish_y;
	const float upper = stepf(scy1, 0.0f);                            // frag:27-29
	const float left = stepf(scx1, 0.0f);
	const float ostep = stepf(fabsf(scy1) - fabsf(scx1) * RESX / RESY, 0.0f);
	const float seg_up = (1 - upper) * (1 - ostep);                   // frag:31-34
	const float seg_dn = (upper) * (1 - ostep);
	const float seg_rt = (1 - left) * (ostep);
	const float seg_lt = (left) * (ostep);
	const float o2 = (ostep * fx + (1 - ostep) * fy) / RESX;          // frag:36
	const float ang2 = scx1 * fabsf(1 - upper - P.vanish_y) / scy1 +  // frag:38-40
	                   upper * (1 - P.vanish_x) +
	                   (1 - upper) * (P.vanish_x);
	float ang3 = scy1 * fabsf(1 - left - P.vanish_x) / scx1 +         // frag:42-44
	             left * (1 - P.vanish_y) +
	             (1 - left) * (P.vanish_y);
	ang3 = ang3 * RESY / RESX + border;                               // frag:46
	const float x_pre = (ostep * ang3 + ang2 * (1 - ostep));          // frag:53
	float ty = seg_dn * (P.ofs_add[1] + x_pre) +                      // frag:56-60
	           seg_up * (P.ofs_add[0] + 1.0f - x_pre) +
	           seg_lt * (P.ofs_add[3] + x_pre) +
	           seg_rt * (P.ofs_add[2] + 1.0f - x_pre);
	ty = ty * P.ratio * 0.25f;                                        // frag:62
	const float rg = P.rot_x_gt0 ? 1.0f : 0.0f;
	const float seg_up_x = rg * seg_up + (1.0f - rg) * seg_dn;        // frag:68-71
	const float seg_dn_x = rg * seg_dn + (1.0f - rg) * seg_up;
	const float seg_rt_x = rg * seg_rt + (1.0f - rg) * seg_lt;
	const float seg_lt_x = rg * seg_lt + (1.0f - rg) * seg_rt;
	const float tx = (seg_up_x) * (o2 + border)                       // frag:73-77
	               + (seg_dn_x) * (1.0f - (o2 + border))
	               + (seg_rt_x) * (o2)
	               + (seg_lt_x) * (1.0f - o2);
	// GL_NEAREST + CLAMP_TO_EDGE (R/src/GL_Main.cpp:154-157): texel = floor(coord * size), clamped
	int ix = f2i(floorf(tx * (float)P.RS));
	int iy = f2i(floorf(ty * (float)P.RC));
	ix = ix < 0 ? 0 : (ix >= P.RS ? P.RS - 1 : ix);
	iy = iy < 0 ? 0 : (iy >= P.RC ? P.RC - 1 : iy);
	if (P.ray_end >= 0 && (iy < P.ray_begin || iy >= P.ray_end)) return 0u;   // slice mode
	if (P.slice_n > 1 && (iy / P.slice_block) % P.slice_n != P.slice_rank) return 0u;
	const uint32_t t = __ldg(P.warp + (size_t)iy * P.RS + ix);
	const float cr = (float)(t & 255u) / 255.0f, cg = (float)((t >> 8) & 255u) / 255.0f;
	const float cb = (float)((t >> 16) & 255u) / 255.0f, ca = (float)(t >> 24) / 255.0f;
	float r, g, b, fragz = 0.0f;
	if (cb != 1.0f)                                                   // frag:89-121
	{
		const float zz = (cb * (1.0f / 256.0f) + ca);
		fragz = 0.001f / zz;
		const float light = (1.0f - cg) * 1.0f + (0.0f + cr) * 0.3f - 0.5f;
		const float pw = 1.2f * powf(fmaxf(light, 0.0f), 4.0f);
		r = light * 1.3f + pw * 1.2f;
		g = light * 0.9f + pw * 1.2f;
		b = light * 0.7f + pw * 1.2f;
	}
	else { r = 178.0f / 255.0f; g = 204.0f / 255.0f; b = 1.0f; }      // frag:125-126
	return quant8(r) | (quant8(g) << 8) | (quant8(b) << 16) | (quant8(fragz) << 24);
}

__global__ void __launch_bounds__(256) k_unwarp(const __grid_constant__ UnwarpParams P)
{
	const int qx = blockIdx.x * blockDim.x + threadIdx.x;           // group of 4 pixels
	const int rowi = P.row_begin + blockIdx.y;
	const int px0 = qx * 4;
	if (px0 >= P.W || rowi >= P.row_end) return;
	const int pyg = P.H - 1 - rowi;                                  // GL row
	uint32_t out[4];
	#pragma unroll
	for (int k = 0; k < 4; k++) out[k] = (px0 + k < P.W) ? unwarp_pixel(P, px0 + k, pyg) : 0u;
	uint32_t* dst = reinterpret_cast<uint32_t*>(P.rgba) + (size_t)rowi * P.W + px0;
	if (px0 + 3 < P.W && ((P.W & 3) == 0))
		*reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
	else
		for (int k = 0; k < 4 && px0 + k < P.W; k++) dst[k] = out[k];
}

void launch_unwarp(const UnwarpParams& p, cudaStream_t st)
{
	const int rows = p.row_end - p.row_begin;
	if (rows <= 0) return;
	const int quads = (p.W + 3) / 4;
	dim3 block(64, 1, 1);
	dim3 grid((quads + 63) / 64, rows, 1);
	k_unwarp<<<grid, block, 0, st>>>(p);
}

__global__ void k_fill_u32(uint32_t* p, uint32_t v, size_t n)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) p[i] = v;
}

void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t st)
{
	if (!n) return;
	k_fill_u32<<<148 * 8, 256, 0, st>>>(p, v, n);
}

} // namespace rlerc
