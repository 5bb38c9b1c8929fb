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
#include "device_common.cuh"

namespace rlerc {

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
	const float res_y2 = (float)(res_y / 2);             // Cuda_Render.h:108 (integer division)
	uint32_t* row = P.warp + (size_t)x * res_y;

	// ---- A/B. ray set-up and screen-space clip (device_common.cuh) ------------------------------
	RayInit ri;
	ray_init(P, x, ri);
	clear_outside<G>(row, res_y, ri, gl);
	if (ri.skip) return;
	const float ray_x = ri.ray_x, ray_z = ri.ray_z, rx2mr = ri.rx2mr;
	const bool vertical = ri.vertical;
	const float sin_x = P.sin_x, cos_x = P.cos_x;
	int ycmin = ri.ycmin, ycmax = ri.ycmax;
	const int ymin0 = ycmin, ymax0 = ycmax;

	// ---- C. occlusion mask clear; the sky sentinel is written at the end to the pixels
	//         that stayed open (same final row as clear-then-overwrite, Cuda_Render.h:255-264)
	for (int w = gl; w < P.mask_words; w += G) ymask[w] = 0;
	__syncwarp(gmask);

	// ---- D. DDA initialisation (device_common.cuh) ---------------------------------------------
	const float vpx = P.viewpos[0], mountain = P.viewpos[1], vpz = P.viewpos[2];
	Dda dd;
	dda_init(P, ray_x, ray_z, dd);
	const int fixx = dd.fixx, fixz = dd.fixz;
	float g0x = dd.g0x, g0y = dd.g0y, g1x = dd.g1x, g1y = dd.g1y, i0x = dd.i0x, i0y = dd.i0y, i1x = dd.i1x, i1y = dd.i1y;
	float gd0 = dd.gd0, gd1 = dd.gd1, d0 = dd.d0, d1 = dd.d1;
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


template <int G, bool IDS>
static void launch_traverse_t(const TraverseParams& p, cudaStream_t st)
{
	const int gpb = RLERC_BLOCK / G;
	const int rays = (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
	if (rays <= 0) return;
	const int blocks = (rays + gpb - 1) / gpb;
	const size_t smem = ((size_t)gpb * G * 8 + (size_t)gpb * p.mask_words) * sizeof(uint32_t);
	// dynamic shared memory above 48 KB is an opt-in per kernel AND per device
	static size_t configured_on[64] = { 0 };
	int dev = 0;
	cudaGetDevice(&dev);
	size_t& configured = configured_on[dev & 63];
	if (smem > configured)
	{
		cudaFuncSetAttribute(k_traverse<G, IDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = smem;
	}
	k_traverse<G, IDS><<<blocks, RLERC_BLOCK, smem, st>>>(p);
}

void launch_traverse(const TraverseParams& p, int lanes, bool ids, cudaStream_t st)
{
	if (lanes == 0 || lanes == 65) { launch_traverse_filter(p, ids, st); return; }
	if (lanes == 68) { if (ids) launch_traverse_filter(p, ids, st); else launch_traverse_pair(p, st); return; }
	if (lanes == 69) { if (ids) launch_traverse_filter(p, ids, st); else launch_traverse_quad(p, st); return; }
#if RLERC_VARIANTS
	if (lanes == 64) { launch_traverse_warp(p, ids, st); return; }
	if (lanes == 66 || lanes == 67) { launch_traverse_chunk(p, ids, lanes == 67 ? 7 : 3, st); return; }
#endif
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
// Unwarp + shade (GLSL pass 1, R/bin/shader/colorize_buddha_soft.frag:12-137, uniforms R/src/main.cpp:578-603).
// One thread produces a tile of 4 horizontally adjacent pixels x RLERC_UNWARP_ROWS rows and writes every row of it
// with one 128-bit store.  Pixel (px, py) with GL origin bottom-left; output row 0 = top.
//
// The shader evaluates ~12 IEEE divisions per fragment.  Most of them do not depend on the fragment: here everything
// that depends only on the column (scx, scx1, left, |scx1|*RESX/RESY, ...) is computed once per thread column and
// everything that depends only on the row (scy, scy1, upper, fy/RESX, ...) once per thread row; the step() selectors
// are 0.0 or 1.0, so each `sel * a + (1 - sel) * b` of the shader IS a or b (x*1 + y*0 == x exactly for finite
// operands), and only the selected branch is evaluated — with the same operations in the same order, so the texel
// chosen is bit-identical to the CPU restatement in oracle/ (checked per pixel by tests/test_gpu_parity.py through
// rlerc_debug_unwarp_texels).  One case needs care: the branch that is NOT selected can be non-finite (a division by
// scy1 == 0 or scx1 == 0) and then poisons x_pre through inf * 0 = NaN; that is reproduced explicitly.
// c / 255.0f of the four texel bytes comes from a 256-entry table in shared memory (the same IEEE quotient).
//
// Multi-GPU "pull" mode (P.peer_n > 1, csrc/group.cu): ray plane iy lives in the warped buffer of GPU
// (iy / slice_block) % peer_n; the texel is loaded from that GPU's buffer over NVLink (peer mapping), and P.rgba may
// point into another GPU's frame buffer: unwarp and compositing are one kernel, no collective moves pixel data.
// Thread <-> pixel mapping: a thread produces 4 horizontally adjacent pixels (one 128-bit store), a WARP a tile of
// RLERC_UNWARP_QW quads x (32 / QW) rows = 8 pixels x 16 rows, a block RLERC_UNWARP_WARPS such tiles side by side.
// In the upper / lower segments the texels of vertically adjacent pixels are adjacent along one ray (ix = row + const),
// in the left / right segments those of horizontally adjacent pixels are: the 128 gathers of a warp tile fall into
// 16-32 sectors of 32 bytes instead of 128 with a row-by-row order (one load instruction touches 4-16 sectors, the
// other three mostly hit L1) — fewer L2 requests locally, and what the NVLink pull of the multi-GPU mode pays for.  The
// stores stay full 32-byte sectors (two lanes per row).
#ifndef RLERC_UNWARP_QW
#define RLERC_UNWARP_QW 2
#endif
#ifndef RLERC_UNWARP_WARPS
#define RLERC_UNWARP_WARPS 4
#endif
#ifndef RLERC_UNWARP_MINB
#define RLERC_UNWARP_MINB 12
#endif
#define RLERC_UNWARP_TROWS (32 / RLERC_UNWARP_QW)

__device__ __forceinline__ uint32_t quant8(float c)
{
	c = fminf(fmaxf(c, 0.0f), 1.0f);
	return (uint32_t)__float2int_rz(c * 255.0f + 0.5f);
}

struct UnwarpCol {            // per pixel column (frag:19,24,28-29,42-44)
	float scx, scx1, axr, Bc, c3;
	bool left;
};
struct UnwarpRow {            // per pixel row (frag:20,25,27,36,38-40)
	float scy, scy1, ay, A, c2, fyx;
	bool upper;
};

__device__ __forceinline__ void unwarp_col(const UnwarpParams& P, int px, UnwarpCol& c)
{
	const float RESX = (float)P.W, RESY = (float)P.H;
	const float fx = (float)px + 0.5f;                               // gl_FragCoord.x
	c.scx = fx / RESX;                                               // frag:19
	c.scx1 = c.scx - P.vanish_x;                                     // frag:24
	c.left = 0.0f >= c.scx1;                                         // frag:28 step(scx1, 0)
	c.axr = fabsf(c.scx1) * RESX / RESY;                             // frag:29, the column half of ostep
	const float left = c.left ? 1.0f : 0.0f;
	c.Bc = fabsf(1 - left - P.vanish_x);                             // frag:42
	c.c3 = c.left ? (1 - P.vanish_y) : P.vanish_y;                   // frag:43-44: left*(1-vy) + (1-left)*vy
}

__device__ __forceinline__ void unwarp_row(const UnwarpParams& P, int pyg, UnwarpRow& r)
{
	const float RESX = (float)P.W, RESY = (float)P.H;
	const float fy = (float)pyg + 0.5f;                              // gl_FragCoord.y
	const float scy = fy / RESY;                                     // frag:20
	r.scy = scy;
	r.scy1 = scy - P.vanish_y;                                       // frag:25
	r.upper = 0.0f >= r.scy1;                                        // frag:27
	r.ay = fabsf(r.scy1);
	const float upper = r.upper ? 1.0f : 0.0f;
	r.A = fabsf(1 - upper - P.vanish_y);                             // frag:38
	r.c2 = r.upper ? (1 - P.vanish_x) : P.vanish_x;                  // frag:39-40
	r.fyx = fy / RESX;                                               // frag:36 for ostep == 0
}

// texel (iy = ray plane, ix = position along the ray) of one fragment
__device__ __forceinline__ void unwarp_texel(const UnwarpParams& P, const UnwarpCol& c, const UnwarpRow& r, int& iy, int& ix)
{
	const float RESX = (float)P.W, RESY = (float)P.H;
	const float border = P.border;                                   // frag:22, evaluated once on the host (same IEEE operations)
	const bool ostep = 0.0f >= r.ay - c.axr;                          // frag:29
	float x_pre, o2;
	if (ostep)
	{
		float ang3 = r.scy1 * c.Bc / c.scx1 + c.c3;                  // frag:42-44
		ang3 = ang3 * RESY / RESX + border;                          // frag:46
		x_pre = (r.scy1 == 0.0f) ? __int_as_float(0x7fc00000) : ang3;   // frag:53: ang2 = .../scy1 is inf or NaN, times 0: NaN
		o2 = c.scx;                                                  // frag:36: (1*fx + 0*fy) / RESX
	}
	else
	{
		const float ang2 = c.scx1 * r.A / r.scy1 + r.c2;             // frag:38-40
		x_pre = (c.scx1 == 0.0f) ? __int_as_float(0x7fc00000) : ang2;   // frag:53: 0 * ang3 with ang3 = .../scx1 non-finite
		o2 = r.fyx;
	}
	// frag:31-34,56-60: exactly one segment selector is 1
	float ty;
	if (!ostep) ty = r.upper ? (P.ofs_add[1] + x_pre) : (P.ofs_add[0] + 1.0f - x_pre);      // seg_dn : seg_up
	else        ty = c.left ? (P.ofs_add[3] + x_pre) : (P.ofs_add[2] + 1.0f - x_pre);       // seg_lt : seg_rt
	ty = ty * P.ratio * 0.25f;                                       // frag:62
	// frag:68-77: rot_x_greater_zero == 0 swaps up <-> dn and rt <-> lt
	const bool flip = !P.rot_x_gt0;
	float tx;
	if (!ostep) tx = ((r.upper != flip) ? (1.0f - (o2 + border)) : (o2 + border));           // seg_dn_x : seg_up_x
	else        tx = ((c.left != flip) ? (1.0f - o2) : o2);                                  // seg_lt_x : seg_rt_x
	// GL_NEAREST + CLAMP_TO_EDGE (R/src/GL_Main.cpp:154-157): texel = floor(coord * size), clamped
	ix = f2i(floorf(tx * (float)P.RS));
	iy = f2i(floorf(ty * (float)P.RC));
	ix = max(0, min(ix, P.RS - 1));
	iy = max(0, min(iy, P.RC - 1));
}

// The same texel with every statement of the shader evaluated as written (frag:19-77).  Used when the vanishing point is
// so far out (|vanish| >= 1e6, or not finite: camera pitch within ~1e-6 of level) that the unselected branch of x_pre
// could overflow to infinity without an exact division by zero — the only case unwarp_texel's shortcut does not cover.
__device__ __forceinline__ float stepf(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }

__device__ __noinline__ void unwarp_texel_generic(const UnwarpParams& P, int px, int pyg, int& iy, int& ix)
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
	ix = f2i(floorf(tx * (float)P.RS));
	iy = f2i(floorf(ty * (float)P.RC));
	ix = ix < 0 ? 0 : (ix >= P.RS ? P.RS - 1 : ix);
	iy = iy < 0 ? 0 : (iy >= P.RC ? P.RC - 1 : iy);
}

// frag:89-136 on the texel's four bytes.  The colour depends on the two attribute bytes only and the smoothing weight
// (alpha) on the two depth bytes only, so both are tabulated over their full 16-bit domains once per context
// (k_shade_tables: the formulas below on every input, hence bit-identical to evaluating them per pixel) and a pixel
// costs two table reads instead of ~60 instructions with five IEEE divisions.
__device__ __forceinline__ uint32_t shade_rgb(uint32_t lo16)          // lo16 = attribute: byte 0 -> c.r, byte 1 -> c.g
{
	const float cr = (float)(lo16 & 255u) / 255.0f, cg = (float)((lo16 >> 8) & 255u) / 255.0f;
	const float light = (1.0f - cg) * 1.0f + (0.0f + cr) * 0.3f - 0.5f;   // frag:90-121
	// pow(max(c, 0), 4) (frag:121) as two squarings: within 1 ulp of powf; the 8-bit result differs from the
	// oracle's powf in < 1e-4 of the pixels, by 1 LSB (the north star's tolerance)
	const float lp = fmaxf(light, 0.0f), lp2 = lp * lp;
	const float pw = 1.2f * (lp2 * lp2);
	const float r = light * 1.3f + pw * 1.2f;
	const float g = light * 0.9f + pw * 1.2f;
	const float b = light * 0.7f + pw * 1.2f;
	return quant8(r) | (quant8(g) << 8) | (quant8(b) << 16);
}
__device__ __forceinline__ uint32_t shade_alpha(uint32_t hi16)        // hi16 = depth: byte 0 -> c.b, byte 1 -> c.a
{
	const float cb = (float)(hi16 & 255u) / 255.0f, ca = (float)(hi16 >> 8) / 255.0f;
	const float zz = (cb * (1.0f / 256.0f) + ca);                         // frag:89,135
	return quant8(0.001f / zz);
}
#define RLERC_SKY_RGBA (178u | (204u << 8) | (255u << 16))              // frag:125-126: (178, 204, 255) / 255, weight 0

__global__ void k_shade_tables(uint32_t* rgb, uint8_t* alpha)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < 65536u) { rgb[i] = shade_rgb(i); alpha[i] = (uint8_t)shade_alpha(i); }
}
void launch_shade_tables(uint32_t* rgb, uint8_t* alpha, cudaStream_t st) { k_shade_tables<<<256, 256, 0, st>>>(rgb, alpha); }

__device__ __forceinline__ uint32_t unwarp_shade(const UnwarpParams& P, uint32_t t)
{
	if (((t >> 16) & 255u) == 255u) return RLERC_SKY_RGBA;                // c.b == 1.0: the sky sentinel (frag:89)
	return __ldg(P.shade_rgb + (t & 0xffffu)) | ((uint32_t)__ldg(P.shade_alpha + (t >> 16)) << 24);
}

// The pass-1 shader of the reference's -DANTIALIAS build (R/src/main.cpp:510-522), colorize_buddha_soft_2xAA.frag:79-139:
// the attribute read as a 5:5:5 normal, a point light at (10, -5) behind the eye, two-term tone curve, vertical sky
// gradient.  Depends on the fragment position and the depth, so it is evaluated per pixel (frame config flag
// RLERC_FLAG_SHADER_2XAA; the texel geometry is the same with the ray-row ratio fixed at 1, frag:62).
__device__ __forceinline__ uint32_t unwarp_shade_2xaa(uint32_t t, float scx, float scy, const float* q)
{
	const float cr = q[t & 255u], cg = q[(t >> 8) & 255u], cb = q[(t >> 16) & 255u], ca = q[t >> 24];
	const int x1 = f2i(cr * 255.0f);                                   // frag:84-88: int(c.r*255.0) truncates
	const int x2 = f2i(cg * 255.0f) * 256 + x1;
	const float col16b = (float)(x2 & 31) / 31.0f, col16g = (float)((x2 >> 5) & 31) / 31.0f, col16r = (float)((x2 >> 10) & 31) / 31.0f;
	float r, g, b, fragz = 0.0f;
	if (cb != 1.0f)
	{
		const float z = (cb * (1.0f / 256.0f) + ca);                    // frag:103-106
		fragz = 0.001f / z;
		const float pos3dx = z * (scx * 2.0f - 1.0f), pos3dy = z * (scy * 2.0f - 1.0f);
		const float nx = 2.0f * col16r - 1.0f, ny = 2.0f * col16g - 1.0f, nz = 2.0f * col16b - 1.0f;   // frag:108-113
		float lx = pos3dx - 10.0f, ly = pos3dy + 5.0f, lz = z;         // frag:115-121
		const float inv = 1.0f / sqrtf(lx * lx + ly * ly + lz * lz);
		lx *= inv; ly *= inv; lz *= inv;
		const float light = nx * lx + ny * ly + nz * lz;               // frag:123
		r = light * 0.9f + light * light * 0.5f;                        // frag:131
		g = light * 0.6f + light * light * 0.4f;
		b = light * 0.3f + light * light * 0.3f;
	}
	else { r = 0.3f * (1.0f - scy) + 0.8f * scy; g = r; b = r; }        // frag:134
	r = r * 1.1f; g = g * 1.1f; b = b * 1.1f;                           // frag:136
	return quant8(r) | (quant8(g) << 8) | (quant8(b) << 16) | (quant8(fragz) << 24);
}

// TEXELS: write (iy << 16) | ix instead of the colour (rlerc_debug_unwarp_texels: per-pixel texel parity against the oracle)
// RARE: any of the uncommon modes is on (multi-GPU slices / pull, the 2xAA shader, the generic texel arithmetic); the
// single-GPU frame of the shipped configuration runs the instantiation without them.
template <bool TEXELS, bool RARE>
__global__ void __launch_bounds__(32 * RLERC_UNWARP_WARPS, RLERC_UNWARP_MINB) k_unwarp(const __grid_constant__ UnwarpParams P)
{
	// the tile's column and row invariants, computed once per block
	constexpr int TW = RLERC_UNWARP_WARPS * RLERC_UNWARP_QW * 4;     // tile width in pixels
	__shared__ UnwarpCol s_col[TW];
	__shared__ UnwarpRow s_row[RLERC_UNWARP_TROWS];
	__shared__ float q255[256];                                      // i / 255.0f (2xAA shader only)
	const int lane = threadIdx.x;
	const int tid = threadIdx.y * 32 + lane;
	const int tile_px = (int)blockIdx.x * TW, tile_row = P.row_begin + (int)blockIdx.y * RLERC_UNWARP_TROWS;
	if (RARE && P.shader == 1) for (int i = tid; i < 256; i += 32 * RLERC_UNWARP_WARPS) q255[i] = (float)i / 255.0f;
	for (int i = tid; i < TW + RLERC_UNWARP_TROWS; i += 32 * RLERC_UNWARP_WARPS)
	{
		if (i < TW) unwarp_col(P, tile_px + i, s_col[i]);
		else unwarp_row(P, P.H - 1 - (tile_row + i - TW), s_row[i - TW]);     // GL row
	}
	__syncthreads();
	const int cx = ((int)threadIdx.y * RLERC_UNWARP_QW + (lane % RLERC_UNWARP_QW)) * 4;   // first pixel of my group of 4, in the tile
	const int ry = lane / RLERC_UNWARP_QW;
	const int px0 = tile_px + cx, rowi = tile_row + ry;
	if (px0 >= P.W || rowi >= P.row_end) return;
	const bool vec = (px0 + 3 < P.W) && ((P.W & 3) == 0);
	const UnwarpRow rw = s_row[ry];
	uint32_t out[4];
	const uint32_t* src[4];
	bool take[4];
	#pragma unroll
	for (int k = 0; k < 4; k++)
	{
		int iy, ix;
		if (RARE && P.generic) unwarp_texel_generic(P, px0 + k, P.H - 1 - rowi, iy, ix);
		else unwarp_texel(P, s_col[cx + k], rw, iy, ix);
		take[k] = px0 + k < P.W;
		if (RARE && P.ray_end >= 0 && (iy < P.ray_begin || iy >= P.ray_end)) take[k] = false;          // slice mode
		const uint32_t* base = P.warp;
		if (RARE && P.slice_n > 1)
		{
			const int owner = (iy / P.slice_block) % P.slice_n;
			if (P.peer_n > 1) base = P.warp_peer[owner];                                       // pull over NVLink
			else if (owner != P.slice_rank) take[k] = false;                                   // interleaved slice mode
		}
		src[k] = base + (size_t)iy * P.RS + ix;
		if (TEXELS) out[k] = ((uint32_t)iy << 16) | (uint32_t)ix;
	}
	if (!TEXELS)
	{
		uint32_t t[4];
		// all four gathers in flight before the first is used.  Read-only path (L1) also for peer memory: another GPU
		// rewrites it between frames, never during this kernel, and L1 does not outlive a kernel
		#pragma unroll
		for (int k = 0; k < 4; k++) t[k] = take[k] ? __ldg(src[k]) : 0u;
		#pragma unroll
		for (int k = 0; k < 4; k++)
			out[k] = !take[k] ? 0u : ((RARE && P.shader == 1) ? unwarp_shade_2xaa(t[k], s_col[cx + k].scx, rw.scy, q255) : unwarp_shade(P, t[k]));
	}
	uint32_t* dst = reinterpret_cast<uint32_t*>(P.rgba) + (size_t)rowi * P.W + px0;
	if (vec) *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
	else for (int k = 0; k < 4 && px0 + k < P.W; k++) dst[k] = out[k];
}

void launch_unwarp(const UnwarpParams& p, cudaStream_t st, bool texels)
{
	const int rows = p.row_end - p.row_begin;
	if (rows <= 0) return;
	const int quads = (p.W + 3) / 4;
	const int qpb = RLERC_UNWARP_QW * RLERC_UNWARP_WARPS;
	dim3 block(32, RLERC_UNWARP_WARPS, 1);
	dim3 grid((quads + qpb - 1) / qpb, (rows + RLERC_UNWARP_TROWS - 1) / RLERC_UNWARP_TROWS, 1);
	const bool rare = p.generic || p.shader != 0 || p.slice_n > 1 || p.ray_end >= 0;
	if (texels) k_unwarp<true, true><<<grid, block, 0, st>>>(p);
	else if (rare) k_unwarp<false, true><<<grid, block, 0, st>>>(p);
	else k_unwarp<false, false><<<grid, block, 0, st>>>(p);
}


// ------------------------------------------------------------------------------------------
// Depth-aware smoothing: GLSL pass 2 (R/bin/shader/soft.frag:1-75, drawn by R/src/main.cpp:628-653).
// The pass-1 image sits in the lower-left W x H corner of a square GL_RGBA8 FBO texture with
// GL_LINEAR / CLAMP_TO_EDGE sampling (R/src/GL_Main.h:151-164); texels outside the window are
// never rendered by the reference (undefined) and read as 0 here.  One thread per output pixel.
__device__ __forceinline__ float4 soft_texel(const SoftParams& P, int i, int j)
{
	if (i >= P.W || j >= P.H) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(P.in) + (size_t)(P.H - 1 - j) * P.W + i);
	return make_float4((float)(t & 255u) / 255.0f, (float)((t >> 8) & 255u) / 255.0f,
	                   (float)((t >> 16) & 255u) / 255.0f, (float)(t >> 24) / 255.0f);
}

// texture2D with GL_LINEAR, CLAMP_TO_EDGE at normalised coordinates (x, y)
__device__ __forceinline__ float4 soft_fetch(const SoftParams& P, float x, float y)
{
	const float u = x * (float)P.fbo - 0.5f, v = y * (float)P.fbo - 0.5f;
	const float fu0 = floorf(u), fv0 = floorf(v);
	const float fu = u - fu0, fv = v - fv0;
	int i0 = f2i(fu0), j0 = f2i(fv0), i1 = i0 + 1, j1 = j0 + 1;
	const int hi = P.fbo - 1;
	i0 = i0 < 0 ? 0 : (i0 > hi ? hi : i0); i1 = i1 < 0 ? 0 : (i1 > hi ? hi : i1);
	j0 = j0 < 0 ? 0 : (j0 > hi ? hi : j0); j1 = j1 < 0 ? 0 : (j1 > hi ? hi : j1);
	const float4 a = soft_texel(P, i0, j0), b = soft_texel(P, i1, j0), c = soft_texel(P, i0, j1), d = soft_texel(P, i1, j1);
	float4 r;
	r.x = (a.x * (1.0f - fu) + b.x * fu) * (1.0f - fv) + (c.x * (1.0f - fu) + d.x * fu) * fv;
	r.y = (a.y * (1.0f - fu) + b.y * fu) * (1.0f - fv) + (c.y * (1.0f - fu) + d.y * fu) * fv;
	r.z = (a.z * (1.0f - fu) + b.z * fu) * (1.0f - fv) + (c.z * (1.0f - fu) + d.z * fu) * fv;
	r.w = (a.w * (1.0f - fu) + b.w * fu) * (1.0f - fv) + (c.w * (1.0f - fu) + d.w * fu) * fv;
	return r;
}

__global__ void __launch_bounds__(256) k_soft(const __grid_constant__ SoftParams P)
{
	const int px = blockIdx.x * blockDim.x + threadIdx.x;
	const int row = blockIdx.y;
	if (px >= P.W || row >= P.H) return;
	const int j = P.H - 1 - row;                                       // GL row
	const float tx = ((float)px + 0.5f) / (float)P.fbo, ty = ((float)j + 0.5f) / (float)P.fbo;   // texCoord
	float4 col = soft_fetch(P, tx, ty);                                // soft.frag:8
	float radmax = col.w;                                              // soft.frag:11-24 (radmin/radavg are dead)
	#pragma unroll
	for (int k = 0; k < 7; k++)
	{
		const float w = soft_fetch(P, P.tap_x[k] + tx, P.tap_y[k] + ty).w;
		radmax = fmaxf(radmax, w);
	}
	const float rad = 0.0023f * radmax;                                // soft.frag:28
	if (rad > 0.00008f)                                                // soft.frag:31
	{
		float4 avg = col;
		float n = 1.0f;
		const float ofs[5] = { -1.0f, -0.6000000238418579f, -0.20000001788139343f, 0.19999998807907104f, 0.6000000238418579f };
		for (int ia = 0; ia < 5; ia++)                                 // soft.frag:38-56: a, b = -1 + k*(2.0/5.0) in float
		for (int ib = 0; ib < 5; ib++)
		{
			const float4 cin = soft_fetch(P, tx + ofs[ia] * rad, ty + ofs[ib] * rad);
			if (cin.w >= radmax * 0.7f)
			{
				avg.x += cin.x; avg.y += cin.y; avg.z += cin.z; avg.w += cin.w;
				n += 1.0f;
			}
		}
		if (n > 6.0f)                                                  // soft.frag:58-66 (colavg2/numtodiv2 mirror colavg/numtodiv)
		{
			const float s = 1.0f / n;
			col = make_float4(avg.x * s, avg.y * s, avg.z * s, avg.w * s);
		}
	}
	reinterpret_cast<uint32_t*>(P.out)[(size_t)row * P.W + px] =
		quant8(col.x) | (quant8(col.y) << 8) | (quant8(col.z) << 16) | (quant8(col.w) << 24);
}

void launch_soft(const SoftParams& p, cudaStream_t st)
{
	dim3 block(128, 1, 1), grid((p.W + 127) / 128, p.H, 1);
	k_soft<<<grid, block, 0, st>>>(p);
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

// Measurement only (tools/l2_peak.py): `iters` passes of 128-bit loads over a buffer; with the buffer L2-resident this is the
// L2 -> SM read bandwidth the roofline block quotes, with a buffer far larger than L2 it is the HBM read bandwidth.
__global__ void __launch_bounds__(512) k_read_u4(const uint4* __restrict__ p, size_t n, int iters, uint32_t* sink)
{
	uint32_t acc = 0;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (int it = 0; it < iters; it++)
	{
		size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
		for (; i + 3 * stride < n; i += 4 * stride)
		{
			const uint4 a = __ldcg(p + i), b = __ldcg(p + i + stride), c = __ldcg(p + i + 2 * stride), d = __ldcg(p + i + 3 * stride);
			acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
		}
		for (; i < n; i += stride) { const uint4 a = __ldcg(p + i); acc ^= a.x ^ a.y ^ a.z ^ a.w; }
	}
	if (acc == 0x9e3779b9u) *sink = acc;              // never true for the zero-filled buffer; keeps the loads alive
}

void launch_read_u4(const void* p, size_t bytes, int iters, uint32_t* sink, cudaStream_t st)
{
	k_read_u4<<<148 * 4, 512, 0, st>>>((const uint4*)p, bytes / 16, iters, sink);
}

} // namespace rlerc
