// k_traverse_w — the production traversal kernel: one WARP per ray plane, lane <-> COLUMN,
// software-pipelined over batches of 32 cell crossings.  Replaces cudaRender +
// Render::render_line (R/src/Cuda_Main.cu:150-181, R/src/Cuda_Render.h:96-737) and gives the
// same warped ray buffer bit for bit (arithmetic contract: see kernels.cu / DESIGN.md §3).
//
// The serial algorithm pays one dependent memory round trip chain per visited column
// (pointer-map entry -> run words -> attribute word) and one ray plane is inherently
// sequential in its occlusion state.  What is NOT sequential is everything that does not read
// that state, so the work of a ray plane is split into a state-free front end, which runs
// 32 columns wide and two batches ahead, and a state-carrying back end, which only spends
// time on the columns that actually draw:
//
//   A   DDA for the next 32 cell crossings (uniform serial float recurrence, integer-budgeted
//       so the inner loop has no LOD / z_far tests); every lane keeps one crossing, turns it
//       into a column address + projected cell, applies a conservative top-clip test and
//       issues its 8-byte pointer-map gather.                                  [batch b+2]
//   C1  the entry has arrived: issue the loads of the column's next run words (the first
//       run rides in the entry).                                                [batch b+1]
//   C2  the run words have arrived: project up to RW runs to screen rows (2 divides each)
//       into shared memory; a run that already breaks under the current horizon ends the
//       column for good (the horizon only rises).                               [batch b]
//   B   consume batch b in front-to-back order: each lane checks whether ITS column would
//       draw under the current floating-horizon bounds; a ballot finds the first such
//       column; all columns before it are provable no-ops and cost nothing.  The owner lane
//       advances the state through its column in the serial statement order; short pixel
//       spans are shaded by the owner (attribute gathers left in flight, stores flushed at the
//       end of the batch), long spans by the whole warp with coalesced stores.  Columns with
//       more than RW undecided runs use the lane <-> run scheme (as k_traverse<32>).
//
// Loads issued in A and C1 have a whole consume phase to land before they are used.
#include <stdint.h>
#include <limits.h>
#include "kernels.cuh"
#include "device_common.cuh"

#include "traverse_variants.cuh"

namespace rlerc {

// Merge-path batch.  The two tracks of the DDA (x-crossings and z-crossings) are independent recurrences
// state += gradient; the serial loop only MERGES them (fire the z-track when d1 < d0, else the x-track).
// So: lanes 0-15 generate the next 33 states of the x-track, lanes 16-31 those of the z-track (33 x 3 adds in
// lockstep instead of 32 x the whole step), and lane s finds crossing s of the merged order with a 5-step
// binary search along its merge-path diagonal (ties go to the x-track, exactly like the serial compare).
// A batch ends early at a LOD switch (the gradients change there).  trk: float4[66], trd: float[72] in shared
// memory.  Returns the crossings made; ra/rb = this lane's records before/after its crossing; *ended = z_far.
// Requires d0, d1 free of NaN (sorted tracks); callers route other rays through dda_serial_batch.
__device__ __forceinline__ int dda_merge_batch(DdaState& S, float4* trk, float* trd, int last_map, int zfar_i, int gl,
                                               float4& ra, float4& rb, bool& ended)
{
	const unsigned FULL = 0xffffffffu;
	while (S.zi > S.mapswitch) dda_lod_switch(S, last_map);
	const int lod_free = (S.mapswitch - S.zi) / S.dzi + 1;
	const int far_free = (zfar_i - S.zi) / S.dzi;
	if (far_free <= 0) { ended = true; return 0; }
	int n = 32;
	n = n < lod_free ? n : lod_free;
	n = n < far_free ? n : far_free;
	{
		const bool zt = gl >= 16;
		float hd = zt ? S.d1 : S.d0, hx = zt ? S.i1x : S.i0x, hy = zt ? S.i1y : S.i0y;
		const float ad = zt ? S.gd1 : S.gd0, ax = zt ? S.g1x : S.g0x, ay = zt ? S.g1y : S.g0y;
		float4* T = trk + (zt ? 33 : 0);
		float* D = trd + (zt ? 36 : 0);
		#pragma unroll 11
		for (int k = 0; k < 33; k++)
		{
			T[k] = make_float4(hd, hx, hy, 0.0f);
			D[k] = hd;
			hd += ad; hx += ax; hy += ay;
		}
	}
	__syncwarp();
	int lo = 0, hi = gl;                                  // x-track elements among the first gl crossings
	#pragma unroll
	for (int it = 0; it < 5; it++)
	{
		if (lo < hi)
		{
			const int mid = (lo + hi) >> 1;
			if (trd[mid] <= trd[36 + gl - mid - 1]) lo = mid + 1; else hi = mid;
		}
	}
	const int ia = lo, jb = gl - lo;
	const bool t1 = trd[36 + jb] < trd[ia];               // Cuda_Render.h:398
	const float4 me = t1 ? trk[33 + jb] : trk[ia];
	rb = make_float4(t1 ? -me.x : me.x, me.y, me.z, __int_as_float(S.mip));
	ra.x = __shfl_up_sync(FULL, rb.x, 1); ra.y = __shfl_up_sync(FULL, rb.y, 1); ra.z = __shfl_up_sync(FULL, rb.z, 1);
	ra.w = 0.0f;
	if (gl == 0) ra = make_float4(S.index ? -S.dist_now : S.dist_now, S.posx, S.posy, 0.0f);
	// state after the n-th crossing: last record + the heads of both tracks
	const int last = n - 1;
	const float lx = __shfl_sync(FULL, rb.x, last);
	S.dist_now = fabsf(lx); S.index = __float_as_int(lx) < 0 ? 1 : 0;
	S.posx = __shfl_sync(FULL, rb.y, last); S.posy = __shfl_sync(FULL, rb.z, last);
	const int in = __shfl_sync(FULL, ia + (t1 ? 0 : 1), last), jn = n - in;
	const float4 ha = trk[in], hb = trk[33 + jn];
	S.d0 = ha.x; S.i0x = ha.y; S.i0y = ha.z;
	S.d1 = hb.x; S.i1x = hb.y; S.i1y = hb.z;
	S.zi += n * S.dzi;
	__syncwarp();
	return n;
}

// ---- decoupled DDA producer ---------------------------------------------------------------------
// The DDA of a ray plane is a serial float recurrence that a consumer warp would execute redundantly
// in all 32 lanes (a third of its instructions).  With PC = true the first blocks of the grid are
// PRODUCERS instead: one lane per ray plane (32 independent recurrences per warp, no redundancy)
// writing batches of 32 crossing records into a small per-ray ring in global memory (L2-resident),
// DEPTH batches ahead of the consumer warp that owns the ray plane.  head[i] / tail[i] are the
// produced / consumed batch counts of launch-local ray i (tail = -1: the consumer is finished).
#define RLERC_RING_DEPTH 8
#define RLERC_RING_SLOT 33          // float4 records per batch: carry + 32 crossings
#define RLERC_SPIN_LIMIT (1 << 22)  // ~0.5 s of polling: a protocol failure must not hang the GPU

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile(int* p, int v) { asm volatile("st.volatile.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

__device__ __noinline__ void dda_producer(const TraverseParams& P, int rays)
{
	const unsigned FULL = 0xffffffffu;
	const int i = blockIdx.x * RLERC_BLOCK + threadIdx.x;      // launch-local ray index of this lane
	bool done = i >= rays;
	Dda dd;
	float posx = 0, posy = 0, dist_now = 0;
	int index = 0, mip = 0, zi = 0, dzi = 1, mapswitch = P.mapswitch0;
	const int last_map = P.nummaps - 1;
	dd.g0x = dd.g0y = dd.g1x = dd.g1y = dd.i0x = dd.i0y = dd.i1x = dd.i1y = dd.gd0 = dd.gd1 = dd.d0 = dd.d1 = 0; dd.fixx = dd.fixz = 0;
	if (!done)
	{
		const int x = owned_ray(P, i);
		RayInit ri;
		ray_init(P, x, ri);
		if (x >= P.ray_end || ri.skip) done = true;          // the consumer returns before it ever looks at the ring
		else
		{
			dda_init(P, ri.ray_x, ri.ray_z, dd);
			for (float yms = P.viewpos[1]; yms > 512.0f; yms = yms * 0.5f)     // Cuda_Render.h:343 (first crossing only)
			{
				if (mip < last_map) mip++;
				dd.g0x *= 2; dd.g0y *= 2; dd.g1x *= 2; dd.g1y *= 2; dd.gd0 *= 2; dd.gd1 *= 2;
				mapswitch *= 2; dzi *= 2;
			}
		}
	}
	int head = 0, spins = 0;
	while (true)
	{
		int tail = 0;
		if (!done) { tail = ld_volatile(P.dda_tail + i); if (tail < 0) done = true; }
		const bool can = !done && (head - tail) < RLERC_RING_DEPTH;
		if (!__any_sync(FULL, can))
		{
			if (__all_sync(FULL, done)) break;
			if (ld_volatile(P.dda_err) || ++spins > RLERC_SPIN_LIMIT) { st_volatile(P.dda_err, 1); break; }
			__nanosleep(256);
			continue;
		}
		spins = 0;
		// every lane fills all the room its ring has (up to DEPTH batches), then ONE fence and ONE
		// publication of head: the fence (all scattered record stores must be visible) is the expensive part
		int made = 0;
		while (__any_sync(FULL, !done && (head + made - tail) < RLERC_RING_DEPTH))
		{
			if (!done && (head + made - tail) < RLERC_RING_DEPTH)
			{
				float4* slot = P.dda_ring + ((size_t)i * RLERC_RING_DEPTH + ((head + made) & (RLERC_RING_DEPTH - 1))) * RLERC_RING_SLOT;
				const float4 carry = make_float4(index ? -dist_now : dist_now, posx, posy, 0.0f);
				int nvalid = 32;
				for (int s = 0; s < 32; s++)
				{
					while (zi > mapswitch)                               // Cuda_Render.h:343-365
					{
						if (mip < last_map) mip++;
						dd.g0x *= 2; dd.g0y *= 2; dd.g1x *= 2; dd.g1y *= 2; dd.gd0 *= 2; dd.gd1 *= 2;
						mapswitch *= 2; dzi *= 2;
					}
					if (zi + dzi > P.z_far) { nvalid = s; break; }       // Cuda_Render.h:366-367
					zi += dzi;
					const bool t1 = dd.d1 < dd.d0;                       // Cuda_Render.h:398-414
					dist_now = t1 ? dd.d1 : dd.d0;
					posx = t1 ? dd.i1x : dd.i0x;
					posy = t1 ? dd.i1y : dd.i0y;
					index = t1 ? 1 : 0;
					__stcg(slot + s + 1, make_float4(t1 ? -dist_now : dist_now, posx, posy, __int_as_float(mip)));
					if (t1) { dd.d1 += dd.gd1; dd.i1x += dd.g1x; dd.i1y += dd.g1y; }
					else    { dd.d0 += dd.gd0; dd.i0x += dd.g0x; dd.i0y += dd.g0y; }
				}
				__stcg(slot, make_float4(carry.x, carry.y, carry.z, __int_as_float(nvalid)));
				made++;
				if (nvalid < 32) done = true;
			}
		}
		__threadfence();
		if (made) { head += made; st_volatile(P.dda_head + i, head); }
	}
}

// MODE: how the DDA advances — 0 serial in every warp, 1 producer blocks + ring, 2 merge path, 3 closed form (lane-parallel)
#ifndef RLERC_MIN_CTAS
#define RLERC_MIN_CTAS 4
#endif
template <bool IDS, int MODE>
__global__ void __launch_bounds__(RLERC_BLOCK, RLERC_MIN_CTAS)
k_traverse_w(const __grid_constant__ TraverseParams P, int rays, int producer_blocks)
{
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int G = 32;
	constexpr int WPB = RLERC_BLOCK / 32;
	const int gl = threadIdx.x & 31;
	const int wid = threadIdx.x >> 5;
	const unsigned FULL = 0xffffffffu;
	constexpr bool PC = (MODE == 1);

	if (PC && (int)blockIdx.x < producer_blocks) { dda_producer(P, rays); return; }
	const int ray_i = ((int)blockIdx.x - (PC ? producer_blocks : 0)) * WPB + wid;    // launch-local ray index
	const int x = owned_ray(P, ray_i);
	if (x >= P.ray_end) return;

	// shared per warp: 33 crossing records (float4) | DrawJob (16 words) | RW x 32 projected runs (int2) |
	//                  RW x 32 deferred short spans | geometry of 3 batches | occlusion bits
	const int per_warp = (RLERC_DDA_WORDS + 16 + RLERC_PS_WORDS + 3 * 6 * 32 + P.mask_words + 3) & ~3;
	uint32_t* wbase = smem + (size_t)wid * per_warp;
	float4* rec = reinterpret_cast<float4*>(wbase);                 // serial DDA: 33 crossing records
	float4* trk = reinterpret_cast<float4*>(wbase);                 // merge-path DDA: 2 x 33 track states (same space)
	float* trd = reinterpret_cast<float*>(wbase + 66 * 4);          //                 2 x 36 track distances
	DrawJob* job = reinterpret_cast<DrawJob*>(wbase + RLERC_DDA_WORDS);
	int2* proj = reinterpret_cast<int2*>(wbase + RLERC_DDA_WORDS + 16);     // [r][lane] = {scr_y1, scr_y2}
	uint32_t* shade = wbase + RLERC_DDA_WORDS + 16 + RLERC_RW * 64;         // [r][lane] deferred short spans
	uint32_t* geo = wbase + RLERC_DDA_WORDS + 16 + RLERC_PS_WORDS;            // [3 batches][6 fields][lane]
	uint32_t* ymask = wbase + RLERC_DDA_WORDS + 16 + RLERC_PS_WORDS + 3 * 6 * 32;

	const int res_y = P.res_y;
	const float res_y2 = (float)(res_y / 2);             // Cuda_Render.h:108 (integer division)
	uint32_t* row = P.warp + (size_t)x * res_y;

	RayInit ri;
	ray_init(P, x, ri);
	clear_outside<G>(row, res_y, ri, gl);
	if (ri.skip) return;
	const float ray_x = ri.ray_x, ray_z = ri.ray_z, rx2mr = ri.rx2mr;
	const bool vertical = ri.vertical;
	const float sin_x = P.sin_x, cos_x = P.cos_x;
	HorizonState Hs;
	Hs.ycmin = ri.ycmin; Hs.ycmax = ri.ycmax; Hs.hiw = 0;
	const int ymin0 = Hs.ycmin, ymax0 = Hs.ycmax;

	// occlusion mask clear; the sky sentinel is written at the end to the pixels that stayed
	// open (same final row as clear-then-overwrite, Cuda_Render.h:255-264)
	for (int w = gl; w < P.mask_words; w += G) ymask[w] = 0;
	__syncwarp();

	const float vpx = P.viewpos[0], mountain = P.viewpos[1], vpz = P.viewpos[2];
	Dda dd;
	dda_init(P, ray_x, ray_z, dd);
	const int fixx = dd.fixx, fixz = dd.fixz;
	DdaState S;
	S.g0x = dd.g0x; S.g0y = dd.g0y; S.g1x = dd.g1x; S.g1y = dd.g1y; S.i0x = dd.i0x; S.i0y = dd.i0y; S.i1x = dd.i1x; S.i1y = dd.i1y;
	S.gd0 = dd.gd0; S.gd1 = dd.gd1; S.d0 = dd.d0; S.d1 = dd.d1;
	S.posx = 0; S.posy = 0; S.dist_now = 0; S.index = 0; S.mip = 0;
	S.zi = 0; S.dzi = 1;                                         // z and dz (Cuda_Render.h:181,325), integer valued
	S.mapswitch = P.mapswitch0;
	const float pz_add = sin_x;                                  // pos3d_z_add (Cuda_Render.h:313)
	const float py_add = (vertical ? cos_x : 0.0f) * rx2mr;      // pos3d_y_add (Cuda_Render.h:314-315)
	const int zfar_i = P.z_far;
	const int last_map = P.nummaps - 1;
	// The y_map_switch half of the LOD loop condition (Cuda_Render.h:343) can only be true on
	// the first crossing (it halves until <= 512 and never grows), where z = 0 < mapswitch.
	for (float yms = mountain; yms > 512.0f; yms = yms * 0.5f) dda_lod_switch(S, last_map);
	// merge path needs sorted tracks: a NaN distance (ray exactly along a grid axis through a lattice point)
	// falls back to the serial recurrence
	const bool sorted_tracks = !(S.d0 != S.d0) && !(S.d1 != S.d1) && !(S.gd0 != S.gd0) && !(S.gd1 != S.gd1);
	const bool merge_ok = (MODE == 2) && sorted_tracks;
	// closed-form DDA (MODE 3): lanes 0..5 own one variable each, the rest of the state is uniform
	const bool closed_ok = (MODE == 3) && sorted_tracks;
	DdaVar var;
	var.b = var.gb = var.F = var.D = var.L = 0;
	DdaUni U;
	U.mip = S.mip; U.zi = S.zi; U.dzi = S.dzi; U.mapswitch = S.mapswitch; U.csd = 0.0f; U.cpx = 0.0f; U.cpy = 0.0f;
	if (MODE == 3 && gl < 6)
	{
		const float v = gl == 0 ? S.d0 : gl == 1 ? S.i0x : gl == 2 ? S.i0y : gl == 3 ? S.d1 : gl == 4 ? S.i1x : S.i1y;
		const float g = gl == 0 ? S.gd0 : gl == 1 ? S.g0x : gl == 2 ? S.g0y : gl == 3 ? S.gd1 : gl == 4 ? S.g1x : S.g1y;
		dda_var_init(var, v, g);
	}

	// per-lane statistics (IDS build only)
	Counters Cn;
	memset(&Cn, 0, sizeof(Cn));

	Stage s0, s1, s2;
	s0.nvalid = s1.nvalid = s2.nvalid = 0;
	s0.have = s1.have = s2.have = false;
	s0.e0 = s0.e1 = 0;
	s1 = s0; s2 = s0;
	int bslot = 0;                    // geo ring slot of the batch produced by A in this iteration
	#pragma unroll
	for (int k = 0; k < 4; k++) { s0.rw[k] = 0; s1.rw[k] = 0; s2.rw[k] = 0; }
	bool dda_done = false;
	int consumed = 0;                 // PC: batches taken from the producer's ring
	float4 pra = make_float4(0, 0, 0, 0), prb = pra, nra = pra, nrb = pra;
	bool have_next = false;
	int known_head = 0;

	RayCtx R;
	R.row = row; R.ymask = ymask; R.ids = IDS ? P.ids + (size_t)x * res_y * 2 : nullptr;
	R.res_y2 = res_y2; R.pz_add = pz_add; R.py_add = py_add; R.mountain = mountain; R.gl = gl; R.stat = nullptr; R.hc_on = 0; R.hc = 0;

	while (true)
	{
		if (Hs.ycmin >= Hs.ycmax) break;                         // Cuda_Render.h:370
		const int ycmin = Hs.ycmin;                              // front end: conservative tests only
		s0 = s1; s1 = s2;
		s2.nvalid = 0; s2.have = false;
		bslot = bslot == 2 ? 0 : bslot + 1;
		const int slot0 = bslot == 2 ? 0 : bslot + 1;            // written two iterations ago: batch s0
		const int slot1 = slot0 == 2 ? 0 : slot0 + 1;            // written one iteration ago: batch s1
		Geo g0;
		{
			const uint32_t* gp = geo + slot0 * 192 + gl;
			g0.pz = __uint_as_float(gp[0]); g0.py = __uint_as_float(gp[32]);
			g0.czz = __uint_as_float(gp[64]); g0.cyy = __uint_as_float(gp[96]);
			g0.cmip = (int)gp[128]; g0.cidx = (int)gp[160];
		}

		// ---- C2. project the runs of batch s0 (their words were requested one iteration ago) ----
		int slen = 0, nr = 0;
		bool longcol = false;
		unsigned flags = 0;               // bit r: run r can be seen (z1 > 0); bit 8+r: its bottom too (z2 > 0)
		if (s0.have)
		{
			{	// run words as loaded in C1 -> runs 0..7, two per register (run 0 rides in the map entry)
				const unsigned first = s0.e1 >> 16;
				const unsigned a = s0.rw[0], b = s0.rw[1], c = s0.rw[2], d = s0.rw[3];
				if (!((2u + s0.e0) & 1u)) s0.rw[0] = first | (a & 0xffff0000u);
				else
				{
					s0.rw[0] = first | (a << 16);
					s0.rw[1] = __funnelshift_r(a, b, 16);
					s0.rw[2] = __funnelshift_r(b, c, 16);
					s0.rw[3] = __funnelshift_r(c, d, 16);
				}
			}
			slen = (int)(s0.e1 & 0xffffu);
			nr = slen < RLERC_RW ? slen : RLERC_RW;
			longcol = slen > RLERC_RW;
			int blen = 0;
			for (int r = 0; r < nr; r++)
			{
				const unsigned rw = run_word(s0.rw, r);
				const int skip = (int)(rw & 1023u), solid = (int)(rw >> 10);
				const int top = (blen + skip) << g0.cmip;                // sti_general_sti_skip
				const int bot = top + (solid << g0.cmip);                // sti_general
				blen += skip + solid;
				if (solid == 0) continue;
				const float ft = (float)top, fb = (float)bot;           // Cuda_Render.h:529-560
				float zz1 = g0.pz, yy1 = g0.py;
				if (mountain + ft >= 0) { zz1 += g0.czz; yy1 += g0.cyy; }
				const float z1 = zz1 + pz_add * ft;
				if (z1 <= 0) continue;
				flags |= 1u << r;
				const float y1 = yy1 + py_add * ft;
				const int sy2 = f2i(res_y2 + y1 / z1);
				int sy1 = 0;
				if (sy2 > ycmin)
				{
					float zz2 = g0.pz, yy2 = g0.py;
					if (mountain + fb < 0) { zz2 += g0.czz; yy2 += g0.cyy; }
					const float z2 = zz2 + pz_add * fb;
					if (!(z2 <= 0))
					{
						flags |= 1u << (8 + r);
						const float y2 = yy2 + py_add * fb;
						sy1 = f2i(res_y2 + y2 / z2 - 1);
					}
				}
				proj[r * 32 + gl] = make_int2(sy1, sy2);
				if (sy2 <= ycmin)
				{
					// breaks now, hence under every later (higher) horizon: later runs are dead
					nr = r + 1; longcol = false;
					break;
				}
			}
		}

		// ---- A. next batch of crossings: DDA, column address, conservative top clip, map gather ---
		if (!dda_done)
		{
			int nvalid = G;
			if (PC)
			{
				// the records of this batch were normally requested one iteration ago (see the end of A)
				if (!have_next)
				{
					int spins = 0;
					while ((known_head = ld_volatile(P.dda_head + ray_i)) <= consumed)
					{
						if (ld_volatile(P.dda_err) || ++spins > RLERC_SPIN_LIMIT) { st_volatile(P.dda_err, 1); nvalid = 0; break; }
						__nanosleep(128);
					}
					const float4* slot = P.dda_ring + ((size_t)ray_i * RLERC_RING_DEPTH + (consumed & (RLERC_RING_DEPTH - 1))) * RLERC_RING_SLOT;
					nra = __ldcg(slot + gl);
					nrb = __ldcg(slot + gl + 1);
				}
				pra = nra;                                            // state before crossing gl
				prb = nrb;                                            // state after crossing gl
				have_next = false;
				const int nv = __float_as_int(__shfl_sync(FULL, pra.w, 0));
				if (nvalid) nvalid = nv;
				// the slot of the PREVIOUS batch is free again: its records were consumed a whole iteration ago
				if (gl == 0 && consumed > 0) st_volatile(P.dda_tail + ray_i, consumed);
				consumed++;
			}
			else if (MODE == 3)
			{
				bool ended = false;
				nvalid = dda_closed_batch(var, U, reinterpret_cast<int*>(wbase + 136), rec, last_map, zfar_i, gl, closed_ok, pra, prb, ended);
				if (ended) dda_done = true;
			}
			else if (merge_ok)
			{
				bool ended = false;
				nvalid = dda_merge_batch(S, trk, trd, last_map, zfar_i, gl, pra, prb, ended);
				if (ended) dda_done = true;
			}
			else
			{
				if (MODE == 0) nvalid = dda_serial_batch_inl(S, rec, last_map, zfar_i);
				else
				{
					DdaState T = S;      // by reference into a non-inlined function: keep S itself in registers
					nvalid = dda_serial_batch(T, rec, last_map, zfar_i);
					S = T;
				}
				if (nvalid < G) dda_done = true;
				__syncwarp();          // every lane wrote every record (same values): ordering made explicit for racecheck
				pra = rec[gl]; prb = rec[gl + 1];
			}
			if (PC && nvalid < G) dda_done = true;
			if (IDS && gl == 0) Cn.c_steps += nvalid;
			s2.nvalid = nvalid;
			if (gl < nvalid)
			{
				Geo g2;
				const float4 ra = pra, rb = prb;                      // state before / after crossing gl
				const float db = fabsf(ra.x), dn = fabsf(rb.x);
				const int ib = __float_as_int(ra.x) < 0 ? 1 : 0;        // index_before: sign bit of the record
				g2.cmip = __float_as_int(rb.w);
				const int fix_x = (1 - ib) * fixx, fix_z = ib * fixz;    // Cuda_Render.h:418-419
				const float ddelta = dn - db;
				const float vsx = ray_x * db, vsz = ray_z * db;
				const int voxel_x = f2i(vpx + ra.y) + fix_x;             // Cuda_Render.h:429-430
				const int voxel_z = f2i(vpz + ra.z) + fix_z;
				const int gx = P.level[g2.cmip].sx, gz = P.level[g2.cmip].sz;
				const int vx = (voxel_x >> g2.cmip) & (gx - 1);          // Cuda_Render.h:441-442
				const int vz = (voxel_z >> g2.cmip) & (gz - 1);
				g2.cidx = vx + vz * gx;
				const float corx = ray_x * ddelta, corz = ray_z * ddelta;
				g2.pz = cos_x * vsz + sin_x * mountain;                  // Cuda_Render.h:459-464
				g2.py = vertical ? (cos_x * mountain - sin_x * vsz) : vsx;
				g2.py *= rx2mr;
				g2.czz = cos_x * corz;                                   // Cuda_Render.h:483-486
				g2.cyy = vertical ? (-sin_x * corz) : corx;
				g2.cyy *= rx2mr;
				// The horizon only rises until this batch is consumed.  For pz > 0 a column culled
				// now stays culled; for pz <= 0 (or NaN) the test can flip, so keep those.
				s2.have = !(g2.pz * res_y2 + g2.py <= g2.pz * (float)ycmin) || !(g2.pz > 0);
				if (s2.have)
				{
					const uint2 ent = __ldg(P.level[g2.cmip].map + g2.cidx);     // Cuda_Render.h:474-478
					s2.e0 = ent.x; s2.e1 = ent.y;
					uint32_t* gp = geo + bslot * 192 + gl;
					gp[0] = __float_as_uint(g2.pz); gp[32] = __float_as_uint(g2.py);
					gp[64] = __float_as_uint(g2.czz); gp[96] = __float_as_uint(g2.cyy);
					gp[128] = (uint32_t)g2.cmip; gp[160] = (uint32_t)g2.cidx;
				}
			}
		}

		// PC: request the crossing records of the next batch now; they land while this batch is consumed
		if (PC && !dda_done && !have_next)
		{
			if (known_head <= consumed) known_head = ld_volatile(P.dda_head + ray_i);
			if (known_head > consumed)
			{
				const float4* slot = P.dda_ring + ((size_t)ray_i * RLERC_RING_DEPTH + (consumed & (RLERC_RING_DEPTH - 1))) * RLERC_RING_SLOT;
				nra = __ldcg(slot + gl);
				nrb = __ldcg(slot + gl + 1);
				have_next = true;
			}
		}

		// ---- C1. run words of batch s1 (its entries were requested one iteration ago) ------------
		if (s1.have)
		{
			const int sl = (int)(s1.e1 & 0xffffu);
			// element i0 of the slab stream is run 0; runs 0..7 are fetched as aligned 32-bit words
			const unsigned i0 = 2u + s1.e0;
			const uint32_t* w32 = reinterpret_cast<const uint32_t*>(P.level[geo[slot1 * 192 + 128 + gl]].slabs);
			const unsigned first = s1.e1 >> 16;
			// only ISSUE the loads here (raw words); they are shifted into place in C2, one iteration later,
			// so that nothing in this iteration waits for them
			const uint32_t* p = w32 + ((i0 + (i0 & 1u)) >> 1);
			const int odd = (int)(i0 & 1u);                            // odd: words hold runs (1,2) (3,4) (5,6) (7,8)
			s1.rw[0] = (sl > 1) ? __ldg(p) : 0u;
			s1.rw[1] = (sl > 2 + odd) ? __ldg(p + 1) : 0u;
			s1.rw[2] = (sl > 4 + odd) ? __ldg(p + 2) : 0u;
			s1.rw[3] = (sl > 6 + odd) ? __ldg(p + 3) : 0u;
			(void)first;
		}

		// ---- B / B0 / S. consume batch s0 (traverse_common.cuh) --------------------------------------
		const bool finished = consume_batch<IDS>(P, R, Hs, Cn, s0, g0, slen, nr, longcol, flags, proj, shade, job);
		if (finished) break;
		if (dda_done && s1.nvalid == 0 && s2.nvalid == 0) break;      // pipeline drained (z > z_far, Cuda_Render.h:367)
	}
	if (PC && gl == 0) st_volatile(P.dda_tail + ray_i, -1);
	__syncwarp();

	// sky sentinel on every pixel of the clip range that no run covered
	for (int y = ymin0 + gl; y <= ymax0; y += G)
		if (!((ymask[y >> 5] >> (y & 31)) & 1u)) row[y] = RLERC_SKY;

	if (IDS) flush_counters(P, Cn, gl, ymax0 - ymin0 + 1);
}

template <bool IDS, int MODE>
static void launch_w(const TraverseParams& p, cudaStream_t st)
{
	const int wpb = RLERC_BLOCK / 32;
	const int rays = (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
	if (rays <= 0) return;
	const int producers = (MODE == 1) ? (rays + RLERC_BLOCK - 1) / RLERC_BLOCK : 0;
	const int blocks = producers + (rays + wpb - 1) / wpb;
	const size_t smem = (size_t)wpb * ((RLERC_DDA_WORDS + 16 + RLERC_PS_WORDS + 3 * 6 * 32 + p.mask_words + 3) & ~3) * sizeof(uint32_t);
	// dynamic shared memory above 48 KB is an opt-in per kernel AND per device
	static size_t configured_on[64] = { 0 };
	int dev = 0;
	cudaGetDevice(&dev);
	size_t& configured = configured_on[dev & 63];
	if (smem > configured)
	{
		cudaFuncSetAttribute(k_traverse_w<IDS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = smem;
	}
	k_traverse_w<IDS, MODE><<<blocks, RLERC_BLOCK, smem, st>>>(p, rays, producers);
}

size_t traverse_ring_bytes(int rays) { return (size_t)rays * RLERC_RING_DEPTH * RLERC_RING_SLOT * sizeof(float4); }

void launch_traverse_warp(const TraverseParams& p, bool ids, cudaStream_t st)
{
	const int mode = p.dda_ring != nullptr ? 1 : (p.dda_mode == 0 ? 0 : (p.dda_mode == 2 ? 2 : 3));
	if (ids)
	{
		if (mode == 1) launch_w<true, 1>(p, st); else if (mode == 0) launch_w<true, 0>(p, st); else if (mode == 2) launch_w<true, 2>(p, st); else launch_w<true, 3>(p, st);
	}
	else
	{
		if (mode == 1) launch_w<false, 1>(p, st); else if (mode == 0) launch_w<false, 0>(p, st); else if (mode == 2) launch_w<false, 2>(p, st); else launch_w<false, 3>(p, st);
	}
}

} // namespace rlerc
