// Pieces shared by the traversal kernels that work lane <-> column (traverse_filter.cu and the variant kernels):
// batch bookkeeping, the span shaders, the long-column path and consume_batch (the occlusion machinery).
#pragma once
#include <stdint.h>
#include <limits.h>
#include "kernels.cuh"
#include "device_common.cuh"

namespace rlerc {

#define RLERC_RW 8          // runs pre-projected per column (first 8 run words)
#ifndef RLERC_COOP_MIN
#define RLERC_COOP_MIN 33   // pixel spans at least this long are shaded by the whole warp (a deferred span record holds <= 32 rows)
#endif
#define RLERC_PS_WORDS (RLERC_RW * 128)   // shared words per warp: RW x 32 projected runs (int2) + RW x 32 deferred span records (uint2)

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

// the n lowest bits (0 <= n <= 32)
__device__ __forceinline__ unsigned row_bits(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

// HEIGHT_COLOR (Cuda_Render.h:716-722): the low attribute byte scaled by height_color >> 12, clamped to a byte, the
// high byte kept.  The product is unsigned as in the reference (uint * int), the clamp is the int min/max.
__device__ __forceinline__ unsigned height_color16(unsigned color16, int hc)
{
	const unsigned pal = color16 & 0xff00u;
	int v = (int)(((color16 & 0xffu) * (unsigned)hc) >> 12);
	v = v > 0 ? v : 0;
	v = v < 255 ? v : 255;
	return (unsigned)v | pal;
}

// bits [lo, hi) of one 32-bit word, lo / hi clamped to the word (empty when hi <= 0 or lo >= 32 or hi <= lo)
__device__ __forceinline__ unsigned bit_range(int lo, int hi)
{
	lo = lo < 0 ? 0 : lo;
	hi = hi > 32 ? 32 : hi;
	if (hi <= lo) return 0u;
	return (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
}

// Everything the span shader needs that is per ray plane (kept in one place so that the rarely
// taken paths can live in non-inlined functions and stay out of the instruction cache)
struct RayCtx {
	uint32_t* row;
	uint32_t* ymask;
	uint32_t* ids;           // IDS build: id words of this ray plane's row, else null
	float res_y2, pz_add, py_add, mountain;
	int gl;
	int hc_on, hc;           // HEIGHT_COLOR (R/src/core.h:22): on/off, height_color (Cuda_Render.h:675)
	long long* stat;         // STATS build: [0] batches [1] B0 taken [2] B1 taken [3] event-loop iterations [4..10] B1 fail reasons [11] batches that needed the event loop [12..15] cycles in B0, B1, event loop, S
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
			unsigned c16 = (unsigned)__ldg(send + ui);
			if (R.hc_on) c16 = height_color16(c16, R.hc);
			st_warp(R.row + yy, c16 + (real_z << 16));
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

// The same span shader for the ownership-resolved path (B1 in consume_batch): pixel y of [y, s2) is written iff its
// bit is set in the claim words cw (bit (y - wbase) of a 128-bit window); the occlusion mask is neither read nor
// written here.  All arguments warp-uniform.
static __device__ __noinline__ void coop_span_claim(const RayCtx& R, const uint16_t* send, float cpz, float cpy,
                                             int y, int s2, int rtop, int rbot, int rtex, int rtexn,
                                             unsigned c0w, unsigned c1w, unsigned c2w, unsigned c3w, int wbase)
{
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
		const int b = yy - wbase;                        // 0 <= b < 128 for every claimed pixel
		const unsigned w = (b < 64) ? ((b < 32) ? c0w : c1w) : ((b < 96) ? c2w : c3w);
		if (gl < steps && b >= 0 && b < 128 && ((w >> (b & 31)) & 1u))
		{
			int ui = f2i(muz / monez);
			ui = (ui > rtex) ? ui : rtex;
			ui = (ui < tex_hi) ? ui : tex_hi;
			const unsigned real_z = (unsigned)f2i(1.0f / monez) & 0xfffeu;
			unsigned c16 = (unsigned)__ldg(send + ui);
			if (R.hc_on) c16 = height_color16(c16, R.hc);
			st_warp(R.row + yy, c16 + (real_z << 16));
		}
	}
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


// ---- consuming one batch of 32 pre-projected columns (shared by k_traverse_w and k_traverse_p) ---------
struct HorizonState { int ycmin, ycmax, hiw; };   // y_clip_min / y_clip_max (Cuda_Render.h:178-179) and one past the highest mask row set
struct Counters {                                  // instrumented build: per-lane partial sums
	unsigned long long c_total, c_proc, c_vox, c_rend, c_pix, c_cols, c_iter, c_cols1, c_steps;
	unsigned long long lstats[5];                  // long_column's share
	unsigned long long dbg[11];                    // fast-path statistics
};

// RESOLVE: everything of a batch that touches the occlusion state (which rows go to which span, the horizon, the long
// spans); the short spans it assigns are left as records in `shade`, one bit per run in shade_runs_out, for shade_batch.
// Returns true when the ray plane is finished (y_clip_min >= y_clip_max, Cuda_Render.h:370).
template <bool IDS, bool STATS = false>
__device__ __forceinline__ bool resolve_batch(const TraverseParams& P, const RayCtx& R, HorizonState& H, Counters& C,
                                              const Stage& s0, const Geo& g0, int slen, int nr, bool longcol, unsigned flags,
                                              const int2* proj, uint32_t* shade, DrawJob* job, unsigned& shade_runs_out)
{
	const unsigned FULL = 0xffffffffu;
	const int gl = R.gl;
	uint32_t* const ymask = R.ymask;
	uint32_t* const row = R.row;
	const float res_y2 = R.res_y2, pz_add = R.pz_add, py_add = R.py_add;
	int ycmin = H.ycmin, ycmax = H.ycmax, hiw = H.hiw;
	unsigned long long& c_total = C.c_total; unsigned long long& c_proc = C.c_proc; unsigned long long& c_vox = C.c_vox;
	unsigned long long& c_rend = C.c_rend; unsigned long long& c_pix = C.c_pix; unsigned long long& c_cols = C.c_cols;
	unsigned long long& c_iter = C.c_iter; unsigned long long& c_cols1 = C.c_cols1;
	unsigned long long* const lstats = C.lstats; unsigned long long* const dbg = C.dbg;

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
	if (STATS && todo) R.stat[0]++;
	long long stick = STATS ? clock64() : 0;
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
			if (STATS) R.stat[1]++;
			int yend = __shfl_sync(FULL, inc, 31);
			yend = yend > y0 ? yend : y0;
			if (draws)
			{
				reinterpret_cast<uint2*>(shade)[ra * 32 + gl] = make_uint2((unsigned)yc | ((unsigned)(T - yc) << 16), row_bits(T - yc));
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

	if (STATS) { const long long now_ = clock64(); R.stat[12] += now_ - stick; stick = now_; }
	// ---- B1. ownership-resolved path (production build) ----------------------------------------------------
	// Every draw is first-come on the occlusion mask, and the reference's break / skip tests only ever discard
	// spans that lie entirely inside already covered rows.  So for REGULAR columns (sy2 non-increasing over the
	// visible runs) and as long as y_clip_max does not move, the outcome of the 32 columns is: row y goes to the
	// first span (column order, then run order) that covers it, y_clip_min is the first open row.  Each lane
	// builds the row set of its spans inside a 128-row window above y_clip_min, an exclusive prefix OR over the
	// lanes gives the rows covered before each column (hence the horizon it meets, for the exact top-clip test
	// of Cuda_Render.h:467) and its own rows.  One failed precondition sends the batch through the event loop.
	// B1 works on the longest PREFIX of the open columns that meets the per-column preconditions; a column that does
	// not (more than RW undecided runs, a span reaching y_clip_max, a span beyond the window) is left to one iteration
	// of the event loop below, after which B1 takes the rest of the batch (at most three attempts per batch).
	int b1_attempts = 0;
	bool ev_batch = false;            // STATS: this batch needed the event loop
	while (todo)
	{
	bool b1_done = false;
	if (!IDS && b1_attempts++ < 3)
	{
		const int y0 = ycmin;
		const int w0 = y0 >> 5, wbase = w0 << 5, wend = wbase + 128;
		bool mine = (todo >> gl) & 1u;
		const bool pass0 = mine && s0.have && !(g0.pz * res_y2 + g0.py <= g0.pz * (float)y0);
		bool ok = true;
		unsigned why = 0;
		if (mine && s0.have && !pass0 && !(g0.pz > 0)) { ok = false; why |= 1; }  // culled now, may pass later
		unsigned rg0 = 0, rg1 = 0, rg2 = 0, rg3 = 0;
		if (pass0)
		{
			if (longcol) { ok = false; why |= 2; }
			int prev2 = INT_MAX;
			for (int r = 0; r < nr; r++)
			{
				if (!((flags >> r) & 1u)) continue;
				const int2 sy = proj[r * 32 + gl];
				if (sy.y > prev2) { ok = false; why |= 4; }                       // irregular column
				prev2 = sy.y;
				if (sy.y <= y0) break;                                            // Cuda_Render.h:543, now and later
				if (!((flags >> (8 + r)) & 1u) || sy.x >= ycmax) continue;        // Cuda_Render.h:556,560
				if (sy.y >= ycmax || sy.y > wend) { ok = false; why |= (sy.y >= ycmax) ? 8 : 16; break; }   // y_clip_max would move / window too small
				const int lo = (sy.x > y0 ? sy.x : y0) - wbase, hi = sy.y - wbase; // 0 <= lo < hi <= 128
				rg0 |= bit_range(lo, hi); rg1 |= bit_range(lo - 32, hi - 32);
				rg2 |= bit_range(lo - 64, hi - 64); rg3 |= bit_range(lo - 96, hi - 96);
			}
		}
		const unsigned badmask = __ballot_sync(FULL, !ok) & todo;
		const int first_bad = badmask ? (__ffs(badmask) - 1) : 32;
		const unsigned sub = todo & (first_bad >= 32 ? 0xffffffffu : ((1u << first_bad) - 1u));   // the columns B1 takes
		if (gl >= first_bad) { rg0 = 0; rg1 = 0; rg2 = 0; rg3 = 0; mine = false; }
		ok = true;
		if (sub != 0)
		{
			// rows covered before my column: mask as it is, rows below y_clip_min, spans of the lanes before me
			unsigned i0 = rg0, i1 = rg1, i2 = rg2, i3 = rg3;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const unsigned t0 = __shfl_up_sync(FULL, i0, d), t1 = __shfl_up_sync(FULL, i1, d);
				const unsigned t2 = __shfl_up_sync(FULL, i2, d), t3 = __shfl_up_sync(FULL, i3, d);
				if (gl >= d) { i0 |= t0; i1 |= t1; i2 |= t2; i3 |= t3; }
			}
			unsigned c0 = __shfl_up_sync(FULL, i0, 1), c1 = __shfl_up_sync(FULL, i1, 1);
			unsigned c2 = __shfl_up_sync(FULL, i2, 1), c3 = __shfl_up_sync(FULL, i3, 1);
			if (gl == 0) { c0 = 0; c1 = 0; c2 = 0; c3 = 0; }
			const int mw = P.mask_words;
			c0 |= ymask[w0] | ((1u << (y0 & 31)) - 1u);
			c1 |= (w0 + 1 < mw) ? ymask[w0 + 1] : 0xffffffffu;
			c2 |= (w0 + 2 < mw) ? ymask[w0 + 2] : 0xffffffffu;
			c3 |= (w0 + 3 < mw) ? ymask[w0 + 3] : 0xffffffffu;
			unsigned m0 = rg0 & ~c0, m1 = rg1 & ~c1, m2 = rg2 & ~c2, m3 = rg3 & ~c3;   // my rows
			if (m0 | m1 | m2 | m3)
			{
				// the horizon this column meets = first open row before it
				const int yc = wbase + (~c0 ? __ffs(~c0) - 1 : (~c1 ? 31 + __ffs(~c1) : (~c2 ? 63 + __ffs(~c2) : 95 + __ffs(~c3))));
				if (g0.pz * res_y2 + g0.py <= g0.pz * (float)yc) { ok = false; why |= 32; }   // Cuda_Render.h:467 under the exact horizon
			}
			// split my rows among my runs (run order), note short spans for S and at most one long span
			int long_r = -1, long_y = 0, long_e = 0;
			unsigned my_shade = 0;
			if (ok && (m0 | m1 | m2 | m3))
			{
				unsigned q0 = m0, q1 = m1, q2 = m2, q3 = m3;
				for (int r = 0; r < nr; r++)
				{
					if (!((flags >> r) & 1u)) continue;
					const int2 sy = proj[r * 32 + gl];
					if (sy.y <= y0) break;
					if (!((flags >> (8 + r)) & 1u) || sy.x >= ycmax) continue;
					const int lo = (sy.x > y0 ? sy.x : y0) - wbase, hi = sy.y - wbase;
					const unsigned a0 = q0 & bit_range(lo, hi), a1 = q1 & bit_range(lo - 32, hi - 32);
					const unsigned a2 = q2 & bit_range(lo - 64, hi - 64), a3 = q3 & bit_range(lo - 96, hi - 96);
					if (!(a0 | a1 | a2 | a3)) continue;
					q0 &= ~a0; q1 &= ~a1; q2 &= ~a2; q3 &= ~a3;
					const int first = a0 ? __ffs(a0) - 1 : (a1 ? 31 + __ffs(a1) : (a2 ? 63 + __ffs(a2) : 95 + __ffs(a3)));
					const int last = a3 ? 127 - __clz(a3) : (a2 ? 95 - __clz(a2) : (a1 ? 63 - __clz(a1) : 31 - __clz(a0)));
					const int n = last - first + 1;
					if (n < RLERC_COOP_MIN)
					{
						// the 32 rows from `first` on, as a bit field
						const int k = first >> 5, sh = first & 31;
						const unsigned lo_w = k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
						const unsigned hi_w = k == 0 ? a1 : (k == 1 ? a2 : (k == 2 ? a3 : 0u));
						const unsigned bits = sh ? ((lo_w >> sh) | (hi_w << (32 - sh))) : lo_w;
						reinterpret_cast<uint2*>(shade)[r * 32 + gl] = make_uint2((unsigned)(wbase + first) | ((unsigned)n << 16), bits);   // scratch until shade_runs says so
						my_shade |= 1u << r;
					}
					else if (long_r < 0) { long_r = r; long_y = wbase + first; long_e = wbase + last + 1; }
					else { ok = false; why |= 64; }                                // two long spans in one column: rare
				}
			}
			if (__all_sync(FULL, ok))
			{
				// ---- commit: nothing was modified before this point ----
				const unsigned t0 = __shfl_sync(FULL, i0, 31), t1 = __shfl_sync(FULL, i1, 31);
				const unsigned t2 = __shfl_sync(FULL, i2, 31), t3 = __shfl_sync(FULL, i3, 31);
				shade_runs |= my_shade;
				// long spans: whole warp, in any order (the rows are already assigned)
				unsigned lb = __ballot_sync(FULL, long_r >= 0);
				while (lb)
				{
					const int L = __ffs(lb) - 1;
					lb &= lb - 1;
					if (gl == L)
					{
						int blen = 0, btex = 0, top = 0, bot = 0, texture = 0, texn = 0;
						for (int r = 0; r <= long_r; r++)
						{
							const unsigned rw = run_word(s0.rw, r);
							const int skip = (int)(rw & 1023u), solid = (int)(rw >> 10);
							top = (blen + skip) << g0.cmip; bot = top + (solid << g0.cmip);
							texture = btex; texn = btex + solid;
							blen += skip + solid; btex += solid;
						}
						job->cpz = g0.pz; job->cpy = g0.py;
						job->y = long_y; job->s2 = long_e; job->rtop = top; job->rbot = bot;
						job->rtex = texture; job->rtexn = texn;
						job->m = g0.cmip; job->colid = g0.cidx; job->e0 = s0.e0; job->slen = (unsigned)slen;
					}
					__syncwarp();
					const DrawJob J = *job;
					coop_span_claim(R, P.level[J.m].slabs + 2 + (size_t)J.e0 + J.slen, J.cpz, J.cpy, J.y, J.s2, J.rtop, J.rbot, J.rtex, J.rtexn,
					                __shfl_sync(FULL, m0, L), __shfl_sync(FULL, m1, L), __shfl_sync(FULL, m2, L), __shfl_sync(FULL, m3, L), wbase);
					__syncwarp();
				}
				// occlusion mask, horizon
				if (gl < 4 && w0 + gl < mw)
				{
					const unsigned t = gl == 0 ? t0 : (gl == 1 ? t1 : (gl == 2 ? t2 : t3));
					if (t) ymask[w0 + gl] |= t;
				}
				__syncwarp();
				const int top_row = (t3 ? 128 - __clz(t3) : (t2 ? 96 - __clz(t2) : (t1 ? 64 - __clz(t1) : (t0 ? 32 - __clz(t0) : 0))));
				if (t0 | t1 | t2 | t3) { const int h = wbase + top_row; hiw = hiw > h ? hiw : h; }
				ycmin = first_clear(ymask, y0, ycmax);                              // Cuda_Render.h:573-577
				todo &= ~sub;
				b1_done = true;
				if (STATS) R.stat[2]++;
			}
		}
		if (!b1_done && sub != 0) b1_attempts = 3;                                // refused as a whole: the event loop takes the batch
		if (STATS && !b1_done)
		{
			const unsigned allwhy = __reduce_or_sync(FULL, why);
			for (int k = 0; k < 7; k++) if ((allwhy >> k) & 1u) R.stat[4 + k]++;
		}
	}
	if (STATS) { const long long now_ = clock64(); R.stat[13] += now_ - stick; stick = now_; }
	if (b1_done) continue;
	if (STATS && !ev_batch) { ev_batch = true; R.stat[11]++; }

	if (STATS) { const long long now_ = clock64(); R.stat[13] += now_ - stick; stick = now_; }   // ---- E. one event-loop iteration
	{
		if (STATS) R.stat[3]++;
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
						const unsigned clear = ~bits & row_bits(n);
						ymask[w] |= clear << sh;
						if (sh && (clear >> (32 - sh))) ymask[w + 1] |= clear >> (32 - sh);
						reinterpret_cast<uint2*>(shade)[r * 32 + gl] = make_uint2((unsigned)y | ((unsigned)n << 16), clear);
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
	if (STATS) { const long long now_ = clock64(); R.stat[14] += now_ - stick; stick = now_; }
	}   // while (todo)

	if (STATS) { const long long now_ = clock64(); R.stat[14] += now_ - stick; stick = now_; }
	H.ycmin = ycmin; H.ycmax = ycmax; H.hiw = hiw;
	shade_runs_out = shade_runs;
	return finished;
}

// SHADE (S): the short spans resolve_batch assigned, every lane its own column, side by side.  Touches neither the
// occlusion mask nor the horizon: it may run later, or on another warp (k_traverse_q).
template <bool IDS, bool STATS = false>
__device__ __forceinline__ void shade_batch(const TraverseParams& P, const RayCtx& R, Counters& C, const Stage& s0, const Geo& g0,
                                            int slen, int nr, unsigned shade_runs, const uint32_t* shade)
{
	uint32_t* const row = R.row;
	const int gl = R.gl;
	const float res_y2 = R.res_y2, pz_add = R.pz_add, py_add = R.py_add;
	unsigned long long& c_pix = C.c_pix;
	long long stick = STATS ? clock64() : 0;
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
			const uint2 jw = reinterpret_cast<const uint2*>(shade)[r * 32 + gl];
			int y = (int)(jw.x & 0xffffu);
			const int n = (int)(jw.x >> 16);
			unsigned clear = jw.y;
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
					if (k0 + k < n && ((clear >> k) & 1u)) st_warp(row + y + k, (R.hc_on ? height_color16(colr[k], R.hc) : colr[k]) + (zz[k] << 16));
				y += 4; clear >>= 4;
			}
		}
	}
	if (STATS) { const long long now_ = clock64(); R.stat[15] += now_ - stick; stick = now_; }
}

// resolve + shade on one warp (k_traverse_f, k_traverse_p, k_traverse_w)
template <bool IDS, bool STATS = false>
__device__ __forceinline__ bool consume_batch(const TraverseParams& P, const RayCtx& R, HorizonState& H, Counters& C,
                                              const Stage& s0, const Geo& g0, int slen, int nr, bool longcol, unsigned flags,
                                              const int2* proj, uint32_t* shade, DrawJob* job)
{
	unsigned shade_runs = 0;
	const bool finished = resolve_batch<IDS, STATS>(P, R, H, C, s0, g0, slen, nr, longcol, flags, proj, shade, job, shade_runs);
	shade_batch<IDS, STATS>(P, R, C, s0, g0, slen, nr, shade_runs, shade);
	return finished;
}

// instrumented build: add this lane's partial sums to the launch totals (layout: include/rlerc.h rlerc_render_counters)
__device__ __forceinline__ void flush_counters(const TraverseParams& P, Counters& C, int gl, int cleared)
{
	if (!P.counters) return;
	atomicAdd(P.counters + 0, C.c_total);
	atomicAdd(P.counters + 1, C.c_proc + C.lstats[1]);
	atomicAdd(P.counters + 2, C.c_vox + C.lstats[2]);
	atomicAdd(P.counters + 3, C.c_rend + C.lstats[3]);
	atomicAdd(P.counters + 4, C.c_pix + C.lstats[4]);
	atomicAdd(P.counters + 5, C.c_cols);
	atomicAdd(P.counters + 6, C.c_iter + C.lstats[0]);
	atomicAdd(P.counters + 7, C.c_cols1);
	if (gl == 0 && cleared > 0) atomicAdd(P.counters + 8, (unsigned long long)cleared);
	atomicAdd(P.counters + 9, C.c_steps);
	if (gl == 0)
	{
		for (int k = 0; k < 3; k++) atomicAdd(P.counters + 10 + k, C.dbg[k]);
		for (int k = 0; k < 8; k++) atomicAdd(P.counters + 16 + k, C.dbg[3 + k]);
	}
}

} // namespace rlerc
