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

namespace rlerc {

#define RLERC_RW 8          // runs pre-projected per column (first 8 run words)
#define RLERC_DDA_WORDS (66 * 4 + 72)   // shared words per warp for the DDA hand-over (merge path: 66 float4 + 72 float)
#define RLERC_COOP_MIN 12   // pixel spans at least this long are shaded by the whole warp

struct DrawJob {            // owner lane -> warp hand-off for a long pixel span (shared memory)
	float cpz, cpy;
	int y, s2, rtop, rbot, rtex, rtexn;
	int m, colid;
	unsigned e0, slen;
};

// One batch of 32 columns in flight: what a lane knows about its column.
// What stays in registers between iterations is only what is in flight from memory; the
// projected cell geometry of the two younger batches waits in shared memory (geo ring).
struct Stage {
	unsigned e0, e1;         // pointer-map entry
	unsigned rw[4];          // run words 0..7, two per register
	int nvalid;              // crossings in this batch (uniform); 0 = empty stage
	bool have;               // this lane's column may be visited
};
struct Geo {
	float pz, py, czz, cyy;  // pos3d_z, pos3d_y (scaled), corr_zz, corr_yy (Cuda_Render.h:459-464,483-486)
	int cmip, cidx;          // mip level and column index vx + vz*gridx
};

// run word r (0..7) of a stage; r is a run-time value, the words live in registers
__device__ __forceinline__ unsigned run_word(const unsigned (&rw)[4], int r)
{
	const unsigned w = (r < 4) ? ((r < 2) ? rw[0] : rw[1]) : ((r < 6) ? rw[2] : rw[3]);
	return (w >> ((r & 1) * 16)) & 0xffffu;
}

// Everything the span shader needs that is per ray plane (kept in one place so that the rarely
// taken paths can live in non-inlined functions and stay out of the instruction cache)
struct RayCtx {
	uint32_t* row;
	uint32_t* ymask;
	uint32_t* ids;           // IDS build: id words of this ray plane's row, else null
	float res_y2, pz_add, py_add, mountain;
	int gl;
};

// A pixel span [y, s2) of one run, shaded by the whole warp 32 pixels at a time, stores coalesced
// (Cuda_Render.h:645-733).  All arguments are warp-uniform.  Returns the number of pixels written.
template <bool IDS>
__device__ __noinline__ int coop_span(const RayCtx& R, const uint16_t* send, float cpz, float cpy,
                                      int y, int s2, int rtop, int rbot, int rtex, int rtexn, int m, int colid)
{
	const unsigned FULL = 0xffffffffu;
	const int gl = R.gl;
	const float ft = (float)rtop, fb2 = (float)rbot;
	const float z1r = cpz + R.pz_add * ft, y1r = cpy + R.py_add * ft;
	const float z2r = cpz + R.pz_add * fb2, y2r = cpy + R.py_add * fb2;
	const float s2r = R.res_y2 + y1r / z1r;
	const float s1r = R.res_y2 + y2r / z2r;
	const float u1z = (float)rtexn / z2r;
	float u2dz = (float)rtex / z1r - u1z;
	const float onez1 = 1.0f / z2r;
	float onedz2 = 1.0f / z1r - onez1;
	u2dz /= s2r - s1r;
	onedz2 /= s2r - s1r;
	const float mult = (float)(y + 1) - s1r;
	float uz = u1z + u2dz * mult;
	float onez = onez1 + onedz2 * mult;
	const int tex_hi = rtexn - 1;                      // int(float(tex-1.0))
	const int n = s2 - y;
	int written = 0;
	for (int c0 = 0; c0 < n; c0 += 32)
	{
		const int steps = (n - c0 < 32) ? (n - c0) : 32;
		float muz = uz, monez = onez;
		for (int t = 0; t < steps; t++)
		{
			if (gl == t) { muz = uz; monez = onez; }
			uz += u2dz; onez += onedz2;
		}
		const int yy = y + c0 + gl;
		bool wr = false;
		if (gl < steps && !((R.ymask[yy >> 5] >> (yy & 31)) & 1u))
		{
			wr = true;
			int ui = f2i(muz / monez);
			ui = (ui > rtex) ? ui : rtex;
			ui = (ui < tex_hi) ? ui : tex_hi;
			const unsigned real_z = (unsigned)f2i(1.0f / monez) & 0xfffeu;
			R.row[yy] = (unsigned)__ldg(send + ui) + (real_z << 16);
			if (IDS) { R.ids[yy * 2] = (uint32_t)colid; R.ids[yy * 2 + 1] = ((uint32_t)m << 16) | (uint32_t)ui; }
		}
		const unsigned wb = __ballot_sync(FULL, wr);
		if (wb && gl == 0)
		{
			const int y0 = y + c0, wi = y0 >> 5, sh = y0 & 31;
			R.ymask[wi] |= wb << sh;
			if (sh && (wb >> (32 - sh))) R.ymask[wi + 1] |= wb >> (32 - sh);
		}
		written += __popc(wb);
		__syncwarp();
	}
	return written;
}

// A column with more than RW undecided runs: lane <-> run, 32 runs at a time (the scheme of
// k_traverse<32>): coalesced run loads, shuffle prefix sum for the y extents and attribute
// offsets, all runs projected in parallel, ballots find the runs that change state in order.
// Arguments warp-uniform; ycmin/ycmax/hiw (one past the highest mask row set) are updated.  stats (IDS): [0] iterations [1] processed
// [2] voxels [3] rendered [4] pixels, per lane partial sums.
template <bool IDS>
__device__ __noinline__ void long_column(const RayCtx& R, const uint16_t* slabs, unsigned e0, unsigned e1,
                                         float cpz, float cpy, float cczz, float ccyy, int m, int colid,
                                         int& ycmin_io, int& ycmax_io, int& hiw_io, unsigned long long* stats)
{
	const unsigned FULL = 0xffffffffu;
	const int gl = R.gl;
	int ycmin = ycmin_io, ycmax = ycmax_io;
	const int cslen = (int)(e1 & 0xffffu);
	const unsigned first = e1 >> 16;
	const uint16_t* runs = slabs + 2 + (size_t)e0;         // Cuda_Render.h:498-499
	const uint16_t* send = runs + cslen;
	int base_len = 0, base_tex = 0;
	bool done = false;
	for (int c = 0; c < cslen && !done; c += 32)
	{
		const int j = c + gl;
		unsigned r = 0;
		if (j < cslen) r = (j == 0) ? first : (unsigned)__ldg(runs + j);
		const int skip = (int)(r & 1023u), solid = (int)(r >> 10);
		// inclusive prefix sum of {skip+solid, solid}, packed 16:16
		const unsigned v = ((unsigned)(skip + solid) << 16) | (unsigned)solid;
		unsigned inc = v;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const unsigned t = __shfl_up_sync(FULL, inc, d);
			if (gl >= d) inc += t;
		}
		const unsigned exc = inc - v;
		const int top = (base_len + (int)(exc >> 16) + skip) << m;      // sti_general_sti_skip
		const int bot = top + (solid << m);                              // sti_general
		const int texture = base_tex + (int)(exc & 0xffffu);
		const int texn = texture + solid;                                // tex
		const unsigned tot = __shfl_sync(FULL, inc, 31);
		base_len += (int)(tot >> 16);
		base_tex += (int)(tot & 0xffffu);

		bool v1 = false, v2 = false;                                     // Cuda_Render.h:529-560
		int ry2 = 0, ry1 = 0;
		if (solid > 0)
		{
			const float ft = (float)top, fb = (float)bot;
			float zz1 = cpz, yy1 = cpy;
			if (R.mountain + ft >= 0) { zz1 += cczz; yy1 += ccyy; }
			const float z1 = zz1 + R.pz_add * ft;
			if (!(z1 <= 0))
			{
				v1 = true;
				const float y1 = yy1 + R.py_add * ft;
				ry2 = f2i(R.res_y2 + y1 / z1);
				float zz2 = cpz, yy2 = cpy;
				if (R.mountain + fb < 0) { zz2 += cczz; yy2 += ccyy; }
				const float z2 = zz2 + R.pz_add * fb;
				if (!(z2 <= 0))
				{
					v2 = true;
					const float y2 = yy2 + R.py_add * fb;
					ry1 = f2i(R.res_y2 + y2 / z2 - 1);
				}
			}
		}
		// resolve the runs in order: only "break" and "draw" events change state
		unsigned rem = (cslen - c >= 32) ? FULL : ((1u << (cslen - c)) - 1u);
		int limit = 31;
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
			if (fb < fd) { done = true; limit = fb; break; }               // Cuda_Render.h:543
			rem &= ~((2u << fd) - 1u);
			int s2y = __shfl_sync(FULL, ry2, fd);
			int s1y = __shfl_sync(FULL, ry1, fd);
			const int rtop = __shfl_sync(FULL, top, fd);
			const int rbot = __shfl_sync(FULL, bot, fd);
			const int rtex = __shfl_sync(FULL, texture, fd);
			const int rtexn = __shfl_sync(FULL, texn, fd);
			if (s2y >= ycmax) { s2y = ycmax; ycmax = s1y; }                 // Cuda_Render.h:564-580
			if (s1y <= ycmin)
			{
				s1y = ycmin;
				ycmin = s2y;
				ycmin = first_clear(R.ymask, ycmin, ycmax);
			}
			const int y = first_clear(R.ymask, s1y, s2y);                  // Cuda_Render.h:639-640
			if (y >= s2y) continue;
			hiw_io = hiw_io > s2y ? hiw_io : s2y;
			const int w = coop_span<IDS>(R, send, cpz, cpy, y, s2y, rtop, rbot, rtex, rtexn, m, colid);
			if (IDS && gl == 0) { stats[3]++; stats[4] += w; }
		}
		if (IDS)
		{
			const unsigned reach = (limit >= 31) ? FULL : ((2u << limit) - 1u);
			if ((j < cslen) && ((reach >> gl) & 1u) && solid > 0) { stats[1]++; stats[2] += solid << m; }
			if (done && gl == 0) stats[0] += c + limit + 1;
		}
	}
	if (IDS && !done && gl == 0) stats[0] += cslen;
	ycmin_io = ycmin; ycmax_io = ycmax;
}


// ---- the DDA of one ray plane (state shared by the three ways of advancing it) ----------------------
struct DdaState {
	float g0x, g0y, g1x, g1y, i0x, i0y, i1x, i1y, gd0, gd1, d0, d1;   // Cuda_Render.h:286-300
	float posx, posy, dist_now;                                       // last crossing (pos_vxl, dds_dist_now)
	int index, mip, zi, dzi, mapswitch;                               // z and dz are integer valued
};

__device__ __forceinline__ void dda_lod_switch(DdaState& S, int last_map)          // Cuda_Render.h:343-365
{
	if (S.mip < last_map) S.mip++;
	S.g0x *= 2; S.g0y *= 2; S.g1x *= 2; S.g1y *= 2;
	S.gd0 *= 2; S.gd1 *= 2;
	S.mapswitch *= 2;
	S.dzi *= 2;
}

// Serial batch: up to 32 crossings, all lanes in lockstep; rec[s+1] = state after crossing s, rec[0] = state
// before the batch; record = {dist (negated when the z-track fired), pos.x, pos.y, mip}.  Returns the number
// of crossings made (< 32 only when z_far was reached, Cuda_Render.h:366-367).
__device__ __forceinline__ int dda_serial_batch_inl(DdaState& S, float4* rec, int last_map, int zfar_i)
{
	int nvalid = 32;
	rec[0] = make_float4(S.index ? -S.dist_now : S.dist_now, S.posx, S.posy, 0.0f);
	for (int s = 0; s < 32;)
	{
		while (S.zi > S.mapswitch) dda_lod_switch(S, last_map);
		const int lod_free = (S.mapswitch - S.zi) / S.dzi + 1;     // crossings before z > mapswitch
		const int far_free = (zfar_i - S.zi) / S.dzi;              // crossings with z + dz <= z_far
		if (far_free <= 0) { nvalid = s; break; }
		int n = 32 - s;
		n = n < lod_free ? n : lod_free;
		n = n < far_free ? n : far_free;
		const float mipf = __int_as_float(S.mip);
		float4* out = rec + s + 1;
		for (int j = 0; j < n; j++)
		{
			const bool t1 = S.d1 < S.d0;                           // Cuda_Render.h:398-414
			S.dist_now = t1 ? S.d1 : S.d0;
			S.posx = t1 ? S.i1x : S.i0x;
			S.posy = t1 ? S.i1y : S.i0y;
			out[j] = make_float4(t1 ? -S.d1 : S.d0, S.posx, S.posy, mipf);
			if (t1) { S.d1 += S.gd1; S.i1x += S.g1x; S.i1y += S.g1y; }
			else    { S.d0 += S.gd0; S.i0x += S.g0x; S.i0y += S.g0y; }
		}
		S.index = __float_as_int(out[n - 1].x) < 0 ? 1 : 0;
		S.zi += n * S.dzi;
		s += n;
	}
	return nvalid;
}

// out-of-line copy for the rare NaN fallback of the merge-path build
__device__ __noinline__ int dda_serial_batch(DdaState& S, float4* rec, int last_map, int zfar_i)
{
	return dda_serial_batch_inl(S, rec, last_map, zfar_i);
}

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

// MODE: how the DDA advances — 0 serial in every warp (default), 1 producer blocks + ring, 2 merge path
template <bool IDS, int MODE>
__global__ void __launch_bounds__(RLERC_BLOCK, 4)
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
	const int per_warp = (RLERC_DDA_WORDS + 16 + RLERC_RW * 96 + 3 * 6 * 32 + P.mask_words + 3) & ~3;
	uint32_t* wbase = smem + (size_t)wid * per_warp;
	float4* rec = reinterpret_cast<float4*>(wbase);                 // serial DDA: 33 crossing records
	float4* trk = reinterpret_cast<float4*>(wbase);                 // merge-path DDA: 2 x 33 track states (same space)
	float* trd = reinterpret_cast<float*>(wbase + 66 * 4);          //                 2 x 36 track distances
	DrawJob* job = reinterpret_cast<DrawJob*>(wbase + RLERC_DDA_WORDS);
	int2* proj = reinterpret_cast<int2*>(wbase + RLERC_DDA_WORDS + 16);     // [r][lane] = {scr_y1, scr_y2}
	uint32_t* shade = wbase + RLERC_DDA_WORDS + 16 + RLERC_RW * 64;         // [r][lane] deferred short spans
	uint32_t* geo = wbase + RLERC_DDA_WORDS + 16 + RLERC_RW * 96;            // [3 batches][6 fields][lane]
	uint32_t* ymask = wbase + RLERC_DDA_WORDS + 16 + RLERC_RW * 96 + 3 * 6 * 32;

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
	int ycmin = ri.ycmin, ycmax = ri.ycmax;
	const int ymin0 = ycmin, ymax0 = ycmax;

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
	const bool merge_ok = (MODE == 2) && !(S.d0 != S.d0) && !(S.d1 != S.d1) && !(S.gd0 != S.gd0) && !(S.gd1 != S.gd1);

	// per-lane statistics (IDS build only)
	unsigned long long c_total = 0, c_proc = 0, c_vox = 0, c_rend = 0, c_pix = 0, c_cols = 0, c_iter = 0, c_cols1 = 0, c_steps = 0;

	// one past the highest row of the occlusion mask that is set: above it the mask is clean
	int hiw = 0;

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
	R.res_y2 = res_y2; R.pz_add = pz_add; R.py_add = py_add; R.mountain = mountain; R.gl = gl;
	unsigned long long lstats[5] = { 0, 0, 0, 0, 0 };              // long_column's share of the counters
	unsigned long long dbg[11] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };  // IDS build: fast-path statistics

	while (true)
	{
		if (ycmin >= ycmax) break;                               // Cuda_Render.h:370
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
				pra = rec[gl]; prb = rec[gl + 1];
			}
			if (PC && nvalid < G) dda_done = true;
			if (IDS && gl == 0) c_steps += nvalid;
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

		// ---- B. consume batch s0: only columns that draw under the current bounds change anything ---
		unsigned todo = (s0.nvalid >= 32) ? FULL : ((1u << s0.nvalid) - 1u);
		bool finished = false;
		unsigned shade_runs = 0;          // runs of my column with a deferred short span

		// ---- B0. rising-horizon fast path -----------------------------------------------------------
		// The common far-field batch: the mask is clean above y_clip_min and every column either does
		// nothing or draws ONE bottom-attached short span [y_clip_min, sy2) and thereby raises
		// y_clip_min to sy2 (classic floating horizon).  Then the serial recurrence over the 32 columns
		// is y_{c+1} = max(y_c, sy2_c): a warp prefix maximum.  Every lane proves that its own column
		// is of that kind under ANY horizon it can meet in this batch; one failed proof sends the
		// whole batch through the general event loop below.
		if (IDS && gl == 0 && todo) dbg[0]++;
		if (todo && hiw <= ycmin)
		{
			if (IDS && gl == 0) dbg[1]++;
			const int y0 = ycmin;
			const bool mine = (todo >> gl) & 1u;
			const bool pass0 = mine && s0.have && !(g0.pz * res_y2 + g0.py <= g0.pz * (float)y0);
			bool ok = true;
			int T = INT_MIN, ra = -1;
			int why = 0;
			if (mine && s0.have && !pass0 && !(g0.pz > 0)) { ok = false; why |= 1; }      // culled now, may pass later
			if (pass0)
			{
				if (longcol) { ok = false; why |= 2; }
				int min_before = INT_MAX;
				for (int r = 0; r < nr; r++)
				{
					if (!((flags >> r) & 1u)) continue;
					const int2 sy = proj[r * 32 + gl];
					if (ra < 0)
					{
						if (!((flags >> (8 + r)) & 1u) || sy.x >= ycmax) { min_before = min_before < sy.y ? min_before : sy.y; continue; }
						ra = r; T = sy.y;
						if (T <= y0) { T = INT_MIN; break; }                  // breaks here under every horizon >= y0
						if (sy.x > y0) { ok = false; why |= 4; }
						if (T >= ycmax) { ok = false; why |= 8; }
						if (min_before < T) { ok = false; why |= 16; }
						if (T - y0 >= RLERC_COOP_MIN) { ok = false; why |= 32; }
					}
					else if (sy.y > T) { ok = false; why |= 64; }               // a later run must break after this one drew
				}
				if (ra < 0) T = INT_MIN;
			}
			// exclusive prefix maximum of T in column order = the horizon each column meets
			int inc = T;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const int t = __shfl_up_sync(FULL, inc, d);
				if (gl >= d) inc = inc > t ? inc : t;
			}
			int yc = __shfl_up_sync(FULL, inc, 1);
			if (gl == 0) yc = INT_MIN;
			yc = yc > y0 ? yc : y0;
			const bool draws = pass0 && T > yc;
			// a column that draws must still pass the top-clip test under the horizon it meets
			if (draws && (g0.pz * res_y2 + g0.py <= g0.pz * (float)yc)) { ok = false; why |= 128; }
			if (IDS)
			{
				const unsigned allwhy = __reduce_or_sync(FULL, (unsigned)why);
				if (gl == 0) for (int k = 0; k < 8; k++) if ((allwhy >> k) & 1u) dbg[3 + k]++;
			}
			if (__all_sync(FULL, ok))
			{
				if (IDS && gl == 0) dbg[2]++;
				int yend = __shfl_sync(FULL, inc, 31);
				yend = yend > y0 ? yend : y0;
				if (draws)
				{
					shade[ra * 32 + gl] = (unsigned)yc | ((unsigned)(T - yc) << 16) | (0xfffu << 20);
					shade_runs |= 1u << ra;
				}
				for (int w = (y0 >> 5) + gl; w <= ((yend - 1) >> 5) && yend > y0; w += 32)
				{
					const int wlo = w << 5;
					const int a = (y0 > wlo ? y0 : wlo) - wlo, b = (yend < wlo + 32 ? yend : wlo + 32) - wlo;
					ymask[w] |= ((b >= 32) ? 0xffffffffu : ((1u << b) - 1u)) & ~((1u << a) - 1u);
				}
				if (IDS && mine && s0.have && !(g0.pz * res_y2 + g0.py <= g0.pz * (float)yc))
				{
					c_cols++; c_total += slen; if (slen) c_cols1++;
					int y = yc, it = slen;
					for (int r = 0; r < nr; r++)
					{
						const unsigned rw = run_word(s0.rw, r);
						if (rw >> 10) { c_proc++; c_vox += (int)(rw >> 10) << g0.cmip; }
						if (!((flags >> r) & 1u)) continue;
						if (proj[r * 32 + gl].y <= y) { it = r + 1; break; }
						if (r == ra) { y = T; c_rend++; }
					}
					c_iter += it;
				}
				ycmin = yend;
				hiw = hiw > yend ? hiw : yend;
				todo = 0;
				__syncwarp();
			}
		}

		while (todo)
		{
			if (ycmin >= ycmax) { finished = true; break; }
			const bool mine = (todo >> gl) & 1u;
			const bool pass = mine && s0.have && !(g0.pz * res_y2 + g0.py <= g0.pz * (float)ycmin);   // Cuda_Render.h:467
			// does my column draw (or is it too long to tell)?  also: where would its run loop stop
			bool ev = false;
			int my_iter = slen, my_proc = 0, my_vox = 0;
			if (pass)
			{
				bool brk = false;
				for (int r = 0; r < nr; r++)
				{
					if (IDS)
					{
						const unsigned rw = run_word(s0.rw, r);
						if (rw >> 10) { my_proc++; my_vox += (int)(rw >> 10) << g0.cmip; }
					}
					if (!((flags >> r) & 1u)) continue;
					const int2 sy = proj[r * 32 + gl];
					if (sy.y <= ycmin) { brk = true; if (IDS) my_iter = r + 1; break; }
					if (((flags >> (8 + r)) & 1u) && !(sy.x >= ycmax)) { ev = true; break; }
				}
				if (!ev && !brk && longcol) ev = true;
			}
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
			if (!eb) break;
			todo &= ~((2u << L) - 1u);

			if (__shfl_sync(FULL, (int)longcol, L) == 0)
			{
				// ---- owner lane advances the state through its column, serial statement order ----
				int rnext = 0, blen = 0, btex = 0;
				while (true)
				{
					int act = 0;      // 0 = column finished, 1 = long pixel span handed to the warp
					if (gl == L)
					{
						for (int r = rnext; r < nr; r++)
						{
							const unsigned rw = run_word(s0.rw, r);
							const int skip = (int)(rw & 1023u), solid = (int)(rw >> 10);
							const int top = (blen + skip) << g0.cmip;
							const int bot = top + (solid << g0.cmip);
							const int texture = btex, texn = btex + solid;
							blen += skip + solid; btex += solid;
							if (IDS) { c_iter++; if (solid > 0) { c_proc++; c_vox += solid << g0.cmip; } }
							if (!((flags >> r) & 1u)) continue;
							const int2 sy = proj[r * 32 + gl];
							if (sy.y <= ycmin) break;                                          // Cuda_Render.h:543
							if (!((flags >> (8 + r)) & 1u) || sy.x >= ycmax) continue;
							int s2y = sy.y, s1y = sy.x;
							if (s2y >= ycmax) { s2y = ycmax; ycmax = s1y; }                      // Cuda_Render.h:564-580
							if (s1y <= ycmin)
							{
								s1y = ycmin;
								ycmin = s2y;
								ycmin = first_clear(ymask, ycmin, ycmax);
							}
							const int y = first_clear(ymask, s1y, s2y);                        // Cuda_Render.h:639-640
							if (y >= s2y) continue;
							if (IDS) c_rend++;
							hiw = hiw > s2y ? hiw : s2y;
							const int n = s2y - y;
							if (n >= RLERC_COOP_MIN)
							{
								job->cpz = g0.pz; job->cpy = g0.py;
								job->y = y; job->s2 = s2y; job->rtop = top; job->rbot = bot;
								job->rtex = texture; job->rtexn = texn;
								job->m = g0.cmip; job->colid = g0.cidx; job->e0 = s0.e0; job->slen = (unsigned)slen;
								act = 1; rnext = r + 1;
								break;
							}
							// short span: fix WHICH pixels it owns now (that is all the occlusion state
							// needs) and leave the shading to the end of the batch, where all lanes
							// shade their spans side by side
							const int w = y >> 5, sh = y & 31;
							unsigned bits = ymask[w] >> sh;
							if (sh) bits |= ymask[w + 1] << (32 - sh);
							const unsigned clear = ~bits & ((1u << n) - 1u);
							ymask[w] |= clear << sh;
							if (sh && (clear >> (32 - sh))) ymask[w + 1] |= clear >> (32 - sh);
							shade[r * 32 + gl] = (unsigned)y | ((unsigned)n << 16) | (clear << 20);
							shade_runs |= 1u << r;
						}
					}
					act = __shfl_sync(FULL, act, L);
					__syncwarp();
					if (act == 0) break;
					// long pixel span: whole warp, 32 pixels at a time
					const DrawJob J = *job;
					const int w = coop_span<IDS>(R, P.level[J.m].slabs + 2 + (size_t)J.e0 + J.slen, J.cpz, J.cpy,
					                             J.y, J.s2, J.rtop, J.rbot, J.rtex, J.rtexn, J.m, J.colid);
					if (IDS && gl == 0) c_pix += w;
				}
				ycmin = __shfl_sync(FULL, ycmin, L);
				ycmax = __shfl_sync(FULL, ycmax, L);
				hiw = __shfl_sync(FULL, hiw, L);
			}
			else
			{
				// ---- long column: lane <-> run -----------------------------------------------------
				const int m = __shfl_sync(FULL, g0.cmip, L);
				long_column<IDS>(R, P.level[m].slabs, __shfl_sync(FULL, s0.e0, L), __shfl_sync(FULL, s0.e1, L),
				                 __shfl_sync(FULL, g0.pz, L), __shfl_sync(FULL, g0.py, L),
				                 __shfl_sync(FULL, g0.czz, L), __shfl_sync(FULL, g0.cyy, L),
				                 m, __shfl_sync(FULL, g0.cidx, L), ycmin, ycmax, hiw, lstats);
			}
		}

		// ---- S. shade the short spans of this batch: every lane its own column, side by side -------
		if (shade_runs)
		{
			int blen = 0, btex = 0;
			const uint16_t* send = P.level[g0.cmip].slabs + 2 + (size_t)s0.e0 + slen;
			for (int r = 0; r < nr; r++)
			{
				const unsigned rw = run_word(s0.rw, r);
				const int skip = (int)(rw & 1023u), solid = (int)(rw >> 10);
				const int top = (blen + skip) << g0.cmip;
				const int bot = top + (solid << g0.cmip);
				const int texture = btex, texn = btex + solid;
				blen += skip + solid; btex += solid;
				if (!((shade_runs >> r) & 1u)) continue;
				const unsigned jw = shade[r * 32 + gl];
				int y = (int)(jw & 0xffffu);
				const int n = (int)((jw >> 16) & 15u);
				unsigned clear = jw >> 20;
				// interpolants (Cuda_Render.h:645-680)
				const float ft = (float)top, fb2 = (float)bot;
				const float z1r = g0.pz + pz_add * ft, y1r = g0.py + py_add * ft;
				const float z2r = g0.pz + pz_add * fb2, y2r = g0.py + py_add * fb2;
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
				// Cuda_Render.h:687-733, four pixels at a time: all attribute gathers of a group are in flight
				// before the first store needs one
				for (int k0 = 0; k0 < n; k0 += 4)
				{
					unsigned colr[4], zz[4];
					#pragma unroll
					for (int k = 0; k < 4; k++)
					{
						colr[k] = 0; zz[k] = 0;
						if (k0 + k < n)
						{
							if ((clear >> k) & 1u)
							{
								int ui = f2i(uz / onez);
								ui = (ui > texture) ? ui : texture;
								ui = (ui < tex_hi) ? ui : tex_hi;
								zz[k] = (unsigned)f2i(1.0f / onez) & 0xfffeu;
								colr[k] = __ldg(send + ui);
								if (IDS)
								{
									c_pix++;
									R.ids[(y + k) * 2] = (uint32_t)g0.cidx;
									R.ids[(y + k) * 2 + 1] = ((uint32_t)g0.cmip << 16) | (uint32_t)ui;
								}
							}
							uz += u2dz; onez += onedz2;
						}
					}
					#pragma unroll
					for (int k = 0; k < 4; k++)
						if (k0 + k < n && ((clear >> k) & 1u)) row[y + k] = colr[k] + (zz[k] << 16);
					y += 4; clear >>= 4;
				}
			}
		}
		if (finished) break;
		if (dda_done && s1.nvalid == 0 && s2.nvalid == 0) break;      // pipeline drained (z > z_far, Cuda_Render.h:367)
	}
	if (PC && gl == 0) st_volatile(P.dda_tail + ray_i, -1);
	if (IDS)
	{
		c_iter += lstats[0]; c_proc += lstats[1]; c_vox += lstats[2]; c_rend += lstats[3]; c_pix += lstats[4];
	}
	__syncwarp();

	// sky sentinel on every pixel of the clip range that no run covered
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
		if (gl == 0) for (int k = 0; k < 6; k++) atomicAdd(P.counters + 10 + k, k < 3 ? dbg[k] : 0ull);
		if (gl == 0) P.counters[15] = 0;
		if (gl == 0) for (int k = 0; k < 8; k++) atomicAdd(P.counters + 16 + k, dbg[3 + k]);
	}
}

template <bool IDS, int MODE>
static void launch_w(const TraverseParams& p, cudaStream_t st)
{
	const int wpb = RLERC_BLOCK / 32;
	const int rays = (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
	if (rays <= 0) return;
	const int producers = (MODE == 1) ? (rays + RLERC_BLOCK - 1) / RLERC_BLOCK : 0;
	const int blocks = producers + (rays + wpb - 1) / wpb;
	const size_t smem = (size_t)wpb * ((RLERC_DDA_WORDS + 16 + RLERC_RW * 96 + 3 * 6 * 32 + p.mask_words + 3) & ~3) * sizeof(uint32_t);
	static size_t configured = 0;
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
	const int mode = p.dda_ring != nullptr ? 1 : (p.dda_mode == 0 ? 0 : 2);
	if (ids)
	{
		if (mode == 1) launch_w<true, 1>(p, st); else if (mode == 0) launch_w<true, 0>(p, st); else launch_w<true, 2>(p, st);
	}
	else
	{
		if (mode == 1) launch_w<false, 1>(p, st); else if (mode == 0) launch_w<false, 0>(p, st); else launch_w<false, 2>(p, st);
	}
}

} // namespace rlerc
