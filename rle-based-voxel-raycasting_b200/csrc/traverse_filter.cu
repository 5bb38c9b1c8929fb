// The production traversal: k_dda_states (pre-pass) + k_traverse_f (one WARP per ray plane) / k_traverse_p (two warps
// per ray plane).  Replaces cudaRender + Render::render_line (R/src/Cuda_Main.cu:150-181, R/src/Cuda_Render.h:96-737)
// and gives the same warped ray buffer bit for bit (arithmetic contract: DESIGN.md §3).
//
// Measured on B200 in round 1: a frame's traversal time is the serial chain of its longest ray planes (the ones that
// see sky and walk all ~7000 cell crossings to z_far); > 90 % of the visited columns are no-ops (their first visible
// run already projects at or below the floating horizon y_clip_min, Cuda_Render.h:542-543, a test that needs nothing
// but the 8-byte pointer-map entry and is monotone: the horizon only rises); and 48 % of the warp instructions of a
// frame were the DDA's float recurrence (Cuda_Render.h:398-414), which a warp that owns one ray plane can only
// execute redundantly in all 32 lanes.  Hence three pieces:
//
//   DDA      k_dda_states: one THREAD per ray plane runs the recurrence once, to z_far, and keeps the six floats of
//            its state at every 32nd crossing (24 bytes per 32 crossings; the LOD / z schedule is integer and the
//            same for every ray plane of a frame: LodSched).  In the traversal kernel lane j of a warp restarts
//            from saved state j and re-runs crossings 32j .. 32j+31 of the current 512-crossing chunk into shared
//            memory: the same float operations in the same order, 16 lanes wide instead of one.
//   FILTER   per 32 crossings (lane <-> crossing): column address, projected cell, conservative top-clip test
//            (Cuda_Render.h:467), pointer-map gather (two batches in flight) -> first-run test -> the LIVE columns
//            are compacted (ballot + popc) into a small queue in shared memory, in crossing order.
//   CONSUME  per 32 LIVE columns: run-word loads (issued one round early), projection of up to RW runs,
//            then consume_batch (traverse_common.cuh): rising-horizon prefix-max fast path, ownership-resolved
//            prefix-OR path, owner-lane event loop, cooperative long spans, deferred parallel shading — the exact
//            semantics, every column is re-tested under the exact state.
//
// The filter only ever removes columns that are provable no-ops, so the result does not depend on how far it
// runs ahead.  The instrumented build (IDS) passes every column that survives the top-clip test, because the
// work counters of the byte model count the no-op columns too.
#include <stdint.h>
#include <limits.h>
#include "kernels.cuh"
#include "device_common.cuh"
#include "traverse_common.cuh"

namespace rlerc {

#define RLERC_QCAP 64                       // queue capacity in columns (ring, power of two): < 32 left + <= 32 new
#ifndef RLERC_F_CH
#define RLERC_F_CH 16                       // 32-crossing batches per DDA chunk = lanes that run the recurrence side by side
#endif
#define RLERC_F_SLOTS (RLERC_F_CH * 33)     // record slots per chunk: crossing c of the chunk sits in slot c + (c >> 5), so that
                                            // the DDA lanes (stride 33) and the filter lanes (stride 1) are both conflict-free
#define RLERC_F_REC ((3 * RLERC_F_SLOTS + (RLERC_F_SLOTS + 3) / 4 + 3) & ~3)   // words: distance, pos.x, pos.y (float) + mip level (byte) per slot
#define RLERC_F_QUEUE (8 * RLERC_QCAP)      // words: 8 fields x QCAP, field-major

static_assert(RLERC_F_CH >= 1 && RLERC_F_CH <= 32, "one lane per batch of a chunk");
static_assert((RLERC_F_REC & 3) == 0, "the areas behind the records are 16-byte aligned");

// The DDA state as two register triples, one per track, laid out like the crossing record a lane consumes:
// {distance, pos.x, pos.y}.  The z-track keeps its distance NEGATED (the record marks the track that fired
// by the sign of its distance, and -(a + b) == (-a) + (-b) exactly), so a crossing is: compare, keep the triple of
// the track that fires, three adds.
struct DdaQ {
	float d0, x0, y0;        // x-track: dds_dist0, isect0           (Cuda_Render.h:286-300)
	float nd1, x1, y1;       // z-track: -dds_dist1, isect1
	float gd0, gx0, gy0;     // grad_dist0, grad0
	float ngd1, gx1, gy1;    // -grad_dist1, grad1
	int mip, zi, dzi, mapswitch;   // z and dz are integer valued
};

__device__ __forceinline__ void ddaq_lod_switch(DdaQ& Q, int last_map)          // Cuda_Render.h:343-365
{
	if (Q.mip < last_map) Q.mip++;
	Q.gx0 *= 2; Q.gy0 *= 2; Q.gx1 *= 2; Q.gy1 *= 2;
	Q.gd0 *= 2; Q.ngd1 *= 2;
	Q.mapswitch *= 2;
	Q.dzi *= 2;
}

// DDA state of a ray plane before its first crossing (Cuda_Render.h:270-305) and after the y_map_switch half of the
// LOD loop condition (Cuda_Render.h:343), which can only be true on the first crossing (it halves until <= 512 and
// never grows) where z = 0 <= mapswitch.
__device__ __forceinline__ void ddaq_init(const TraverseParams& P, float ray_x, float ray_z, DdaQ& Q, int& fixx, int& fixz)
{
	Dda dd;
	dda_init(P, ray_x, ray_z, dd);
	fixx = dd.fixx; fixz = dd.fixz;
	Q.d0 = dd.d0; Q.x0 = dd.i0x; Q.y0 = dd.i0y; Q.nd1 = -dd.d1; Q.x1 = dd.i1x; Q.y1 = dd.i1y;
	Q.gd0 = dd.gd0; Q.gx0 = dd.g0x; Q.gy0 = dd.g0y; Q.ngd1 = -dd.gd1; Q.gx1 = dd.g1x; Q.gy1 = dd.g1y;
	Q.mip = 0; Q.zi = 0; Q.dzi = 1;                              // z and dz (Cuda_Render.h:181,325)
	Q.mapswitch = P.mapswitch0;
}

// One crossing (Cuda_Render.h:398-414).  REC: write its record {+-distance, pos.x, pos.y, mip}.
template <bool REC>
__device__ __forceinline__ void ddaq_cross(DdaQ& Q, float* rd, float* rx, float* ry, uint8_t* rm, int i)
{
	const bool t1 = -Q.nd1 < Q.d0;
	if (REC)
	{
		rd[i] = t1 ? Q.nd1 : Q.d0; rx[i] = t1 ? Q.x1 : Q.x0; ry[i] = t1 ? Q.y1 : Q.y0;
		rm[i] = (uint8_t)Q.mip;
	}
	if (t1) { Q.nd1 += Q.ngd1; Q.x1 += Q.gx1; Q.y1 += Q.gy1; }
	else    { Q.d0 += Q.gd0; Q.x0 += Q.gx0; Q.y0 += Q.gy0; }
}

// Up to 32 crossings of ONE lane's DDA.  LOD / z_far budgets by shifts (dz is a power of two); stops early only
// when z_far is reached (Cuda_Render.h:366-367).
template <bool REC>
__device__ __forceinline__ void ddaq_run32(DdaQ& Q, int last_map, int zfar_i, float* rd, float* rx, float* ry, uint8_t* rm)
{
	for (int s = 0; s < 32;)
	{
		while (Q.zi > Q.mapswitch) ddaq_lod_switch(Q, last_map);
		const int sh = 31 - __clz(Q.dzi);
		const int lod_free = ((Q.mapswitch - Q.zi) >> sh) + 1;        // crossings before z > mapswitch
		const int far_free = (zfar_i - Q.zi) >> sh;                   // crossings with z + dz <= z_far (<= 0: none)
		if (far_free <= 0) break;
		int n = 32 - s;
		n = n < lod_free ? n : lod_free;
		n = n < far_free ? n : far_free;
		#pragma unroll 4
		for (int j = 0; j < n; j++) ddaq_cross<REC>(Q, rd, rx, ry, rm, s + j);
		Q.zi += n << sh;
		s += n;
	}
}

// ---- pre-pass: the DDA of every ray plane of the launch, one thread each, state kept at every 32nd crossing -------
// states[(ray_i * nb + b) * 3 + 0..2] = {d0, x0}, {y0, -d1}, {x1, y1} before crossing 32 b (ray_i: launch-local index)
//
// The traversal kernel is launched behind it with programmatic stream serialization and starts as soon as every block
// of this kernel is resident (griddepcontrol.launch_dependents at the top): a ray plane's chain of ~7000 dependent
// crossings takes ~0.1 ms here, which would otherwise sit in front of every frame.  Hand-over per ray plane:
// dda_progress = (epoch << 32) | batches published, written with release semantics after the states, at batch counts
// CH, 3 CH, 7 CH, ... (a fence per publication costs this thread ~1 us, so they thin out) and at the end.
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__global__ void __launch_bounds__(64)
k_dda_states(const __grid_constant__ TraverseParams P, int rays)
{
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	const int ray_i = (int)blockIdx.x * 64 + (int)threadIdx.x;
	if (ray_i >= rays) return;
	const int x = owned_ray(P, ray_i);
	if (x >= P.ray_end) return;
	RayInit ri;
	ray_init(P, x, ri);
	if (ri.skip) return;
	DdaQ Q;
	int fixx, fixz;
	ddaq_init(P, ri.ray_x, ri.ray_z, Q, fixx, fixz);
	const int last_map = P.nummaps - 1;
	for (float yms = P.viewpos[1]; yms > 512.0f; yms = yms * 0.5f) ddaq_lod_switch(Q, last_map);
	const int nb = (P.lod.k_total + 31) >> 5;
	float2* out = P.dda_states + (size_t)ray_i * nb * 3;
	unsigned long long* const flag = P.dda_progress + ray_i;
	const unsigned long long epoch = (unsigned long long)P.dda_epoch << 32;
	int publish_at = RLERC_F_CH;
	for (int b = 0; b < nb; b++, out += 3)
	{
		out[0] = make_float2(Q.d0, Q.x0); out[1] = make_float2(Q.y0, Q.nd1); out[2] = make_float2(Q.x1, Q.y1);
		if (b + 1 == publish_at || b + 1 == nb)
		{
			st_release_u64(flag, epoch | (unsigned)(b + 1));
			publish_at = 2 * publish_at + RLERC_F_CH;
		}
		ddaq_run32<false>(Q, last_map, P.z_far, nullptr, nullptr, nullptr, nullptr);
	}
}

// ---- the crossing records of one chunk, in shared memory -----------------------------------------------------------
struct RecView {
	float* d; float* x; float* y; uint8_t* m;        // [CH * 33] each
};
// words of the record area for chunks of CH batches
__host__ __device__ constexpr int rec_words(int ch) { return (3 * ch * 33 + (ch * 33 + 3) / 4 + 3) & ~3; }
static_assert(rec_words(RLERC_F_CH) == RLERC_F_REC, "record area");
template <int CH = RLERC_F_CH>
__device__ __forceinline__ RecView rec_view(uint32_t* base)
{
	RecView v;
	v.d = reinterpret_cast<float*>(base); v.x = v.d + CH * 33; v.y = v.x + CH * 33;
	v.m = reinterpret_cast<uint8_t*>(v.y + CH * 33);
	return v;
}

// Chunk `chunk` of the ray plane: lane j < CH restarts the DDA from saved state chunk * CH + j (the LOD / z side of
// it from the frame's schedule) and writes the records of its 32 crossings.  `carry` = record of the last crossing
// before this chunk (what the chunk's first crossing starts from).
#define RLERC_DDA_SPIN_MAX (1 << 22)          // x 100 ns: the pre-pass never came; fail loudly (launch error), never a wrong picture
__device__ __noinline__ void dda_states_missing() { __trap(); }

template <int CH = RLERC_F_CH>
__device__ __forceinline__ void dda_chunk(const TraverseParams& P, const float2* st, const unsigned long long* flag, int& published,
                                          float ray_x, float ray_z, int chunk, int gl, const RecView& rec, float3& carry)
{
	if (chunk > 0)
	{
		const int last = CH * 33 - 2;                                // slot of crossing CH * 32 - 1
		carry = make_float3(rec.d[last], rec.x[last], rec.y[last]);
	}
	// the states of this chunk's batches have to be published (k_dda_states runs concurrently, normally far ahead)
	{
		const int nb = (P.lod.k_total + 31) >> 5;
		int need = (chunk + 1) * CH;
		need = need < nb ? need : nb;
		int spins = 0;
		while (published < need)                                     // warp-uniform: every lane reads the same word
		{
			const unsigned long long v = ld_acquire_u64(flag);
			if ((unsigned)(v >> 32) == P.dda_epoch) published = (int)(unsigned)v;
			if (published >= need) break;
			__nanosleep(100);
			if (++spins > RLERC_DDA_SPIN_MAX) { dda_states_missing(); break; }
		}
	}
	__syncwarp();
	const int b = chunk * CH + gl;
	const int k0 = b << 5;
	if (gl < CH && k0 < P.lod.k_total)
	{
		const float2* s = st + (size_t)b * 3;
		const float2 a0 = __ldcg(s), a1 = __ldcg(s + 1), a2 = __ldcg(s + 2);   // L2: written by a kernel that is still running
		DdaQ Q;
		int fixx, fixz;
		ddaq_init(P, ray_x, ray_z, Q, fixx, fixz);
		Q.d0 = a0.x; Q.x0 = a0.y; Q.y0 = a1.x; Q.nd1 = a1.y; Q.x1 = a2.x; Q.y1 = a2.y;
		int p = 0;
		while (p + 1 < P.lod.nphase && P.lod.ph_k[p + 1] <= k0) p++;
		const int nsw = P.lod.ph_nsw[p];
		// nsw doublings of a float are one multiplication by 2^nsw (both exact, both saturate to inf alike)
		const float sc = __int_as_float((127 + nsw) << 23);
		Q.gd0 *= sc; Q.gx0 *= sc; Q.gy0 *= sc; Q.ngd1 *= sc; Q.gx1 *= sc; Q.gy1 *= sc;
		Q.dzi = 1 << nsw;
		Q.zi = P.lod.ph_z[p] + ((k0 - P.lod.ph_k[p]) << nsw);
		Q.mapswitch = P.mapswitch0 << nsw;
		Q.mip = nsw < P.nummaps - 1 ? nsw : P.nummaps - 1;
		const int s0 = gl * 33;
		ddaq_run32<true>(Q, P.nummaps - 1, P.z_far, rec.d + s0, rec.x + s0, rec.y + s0, rec.m + s0);
	}
	__syncwarp();
}

// ---- FILTER, geometry half: crossing gl of batch bl of the chunk -> projected cell, top clip, pointer-map gather -------
struct FilterRay {                                   // per ray plane, constant
	float ray_x, ray_z, rx2mr, sin_x, cos_x, vpx, vpz, mountain, res_y2, pz_add, py_add;
	int fixx, fixz;
	bool vertical;
};

__device__ __forceinline__ bool filter_geometry(const TraverseParams& P, const FilterRay& F, const RecView& rec, const float3& carry,
                                                int bl, int gl, int ycmin, Geo& g, unsigned& e0, unsigned& e1)
{
	const int slot = bl * 33 + gl;
	float ad, ax, ay;                                             // state before / after crossing gl
	if (gl > 0) { ad = rec.d[slot - 1]; ax = rec.x[slot - 1]; ay = rec.y[slot - 1]; }
	else if (bl > 0) { ad = rec.d[slot - 2]; ax = rec.x[slot - 2]; ay = rec.y[slot - 2]; }
	else { ad = carry.x; ax = carry.y; ay = carry.z; }
	const float bd = rec.d[slot];
	g.cmip = (int)rec.m[slot];
	const float db = fabsf(ad), dn = fabsf(bd);
	const int ib = __float_as_int(ad) < 0 ? 1 : 0;                // index_before: sign bit of the record
	const int fix_x = (1 - ib) * F.fixx, fix_z = ib * F.fixz;     // Cuda_Render.h:418-419
	const float ddelta = dn - db;
	const float vsx = F.ray_x * db, vsz = F.ray_z * db;
	const int voxel_x = f2i(F.vpx + ax) + fix_x;                  // Cuda_Render.h:429-430
	const int voxel_z = f2i(F.vpz + ay) + fix_z;
	const int gx = P.level[g.cmip].sx, gz = P.level[g.cmip].sz;
	// CLIPREGION (Cuda_Render.h:432-437): finite scene, columns outside the grid are skipped
	const bool outside = (P.flags & 1) && (voxel_x < 0 || voxel_z < 0 || (voxel_x >> g.cmip) > gx - 1 || (voxel_z >> g.cmip) > gz - 1);
	const int vx = (voxel_x >> g.cmip) & (gx - 1);                // Cuda_Render.h:441-442
	const int vz = (voxel_z >> g.cmip) & (gz - 1);
	g.cidx = vx + vz * gx;
	const float corx = F.ray_x * ddelta, corz = F.ray_z * ddelta;
	g.pz = F.cos_x * vsz + F.sin_x * F.mountain;                  // Cuda_Render.h:459-464
	g.py = F.vertical ? (F.cos_x * F.mountain - F.sin_x * vsz) : vsx;
	g.py *= F.rx2mr;
	g.czz = F.cos_x * corz;                                       // Cuda_Render.h:483-486
	g.cyy = F.vertical ? (-F.sin_x * corz) : corx;
	g.cyy *= F.rx2mr;
	// The horizon only rises.  For pz > 0 a column culled now stays culled; for pz <= 0 (or NaN)
	// the test can flip, so keep those.
	const bool have = !outside && (!(g.pz * F.res_y2 + g.py <= g.pz * (float)ycmin) || !(g.pz > 0));   // Cuda_Render.h:467
	if (have)
	{
		const uint2 ent = __ldg(P.level[g.cmip].map + g.cidx);    // Cuda_Render.h:474-478
		e0 = ent.x; e1 = ent.y;
	}
	return have;
}

// ---- FILTER, test half: can the column do anything under horizon ycmin (or any later, higher one)? -------------------
template <bool IDS>
__device__ __forceinline__ bool filter_live(const FilterRay& F, const Geo& g, unsigned e1, int ycmin)
{
	const int slen = (int)(e1 & 0xffffu);
	const unsigned first = e1 >> 16;
	const int solid = (int)(first >> 10), skip = (int)(first & 1023u);
	if (IDS) return true;                                 // the byte model counts no-op columns too
	if (slen == 0) return false;                          // empty column: the run loop does not execute
	if (solid == 0) return true;                          // pure skip run: undecided, let the machinery look
	const float ft = (float)(skip << g.cmip);             // Cuda_Render.h:529-543 for run 0
	float zz1 = g.pz, yy1 = g.py;
	if (F.mountain + ft >= 0) { zz1 += g.czz; yy1 += g.cyy; }
	const float z1 = zz1 + F.pz_add * ft;
	if (z1 <= 0) return true;                             // `continue`: a later run may be the first visible one
	const float y1 = yy1 + F.py_add * ft;
	return f2i(F.res_y2 + y1 / z1) > ycmin;               // else: break, now and under every later horizon
}

// entry of the live-column queue / ring: 8 words, field-major with stride `cap`
__device__ __forceinline__ void column_put(uint32_t* q, int cap, const Geo& g, unsigned e0, unsigned e1)
{
	q[0 * cap] = __float_as_uint(g.pz); q[1 * cap] = __float_as_uint(g.py);
	q[2 * cap] = __float_as_uint(g.czz); q[3 * cap] = __float_as_uint(g.cyy);
	q[4 * cap] = (uint32_t)g.cmip; q[5 * cap] = (uint32_t)g.cidx;
	q[6 * cap] = e0; q[7 * cap] = e1;
}

// ---- C1: take a live column, request its run words (runs 0..7 as aligned 32-bit words; run 0 rides in the map entry) --
__device__ __forceinline__ void column_take(const TraverseParams& P, const uint32_t* q, int cap, Geo& g, Stage& s)
{
	g.pz = __uint_as_float(q[0 * cap]); g.py = __uint_as_float(q[1 * cap]);
	g.czz = __uint_as_float(q[2 * cap]); g.cyy = __uint_as_float(q[3 * cap]);
	g.cmip = (int)q[4 * cap]; g.cidx = (int)q[5 * cap];
	s.e0 = q[6 * cap]; s.e1 = q[7 * cap];
	const int sl = (int)(s.e1 & 0xffffu);
	const unsigned i0 = 2u + s.e0;                                 // element i0 of the slab stream is run 0
	const uint32_t* w32 = reinterpret_cast<const uint32_t*>(P.level[g.cmip].slabs);
	const uint32_t* p = w32 + ((i0 + (i0 & 1u)) >> 1);
	const int odd = (int)(i0 & 1u);                                // odd: words hold runs (1,2) (3,4) (5,6) (7,8)
	s.rw[0] = (sl > 1) ? __ldg(p) : 0u;
	s.rw[1] = (sl > 2 + odd) ? __ldg(p + 1) : 0u;
	s.rw[2] = (sl > 4 + odd) ? __ldg(p + 2) : 0u;
	s.rw[3] = (sl > 6 + odd) ? __ldg(p + 3) : 0u;
}

__device__ __forceinline__ void stage_clear(Stage& s, Geo& g)
{
	s.nvalid = 0; s.have = false; s.e0 = s.e1 = 0;
	#pragma unroll
	for (int k = 0; k < 4; k++) s.rw[k] = 0;
	g.pz = g.py = g.czz = g.cyy = 0; g.cmip = 0; g.cidx = 0;
}

// ---- C2: project the runs of a lane's column to screen rows (Cuda_Render.h:529-560) under horizon ycmin ---------------
// proj[r * 32 + gl] = {sy1, sy2}; flags bit r: run r can be seen (z1 > 0); bit 8 + r: its bottom too (z2 > 0)
__device__ __forceinline__ void column_project(const FilterRay& F, Stage& s0, const Geo& g0, int ycmin, int gl, int2* proj,
                                               int& slen, int& nr, bool& longcol, unsigned& flags)
{
	{	// run words as loaded in C1 -> runs 0..7, two per register
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
		const float ft = (float)top, fb = (float)bot;
		float zz1 = g0.pz, yy1 = g0.py;
		if (F.mountain + ft >= 0) { zz1 += g0.czz; yy1 += g0.cyy; }
		const float z1 = zz1 + F.pz_add * ft;
		if (z1 <= 0) continue;
		flags |= 1u << r;
		const float y1 = yy1 + F.py_add * ft;
		const int sy2 = f2i(F.res_y2 + y1 / z1);
		int sy1 = 0;
		if (sy2 > ycmin)
		{
			float zz2 = g0.pz, yy2 = g0.py;
			if (F.mountain + fb < 0) { zz2 += g0.czz; yy2 += g0.cyy; }
			const float z2 = zz2 + F.pz_add * fb;
			if (!(z2 <= 0))
			{
				flags |= 1u << (8 + r);
				const float y2 = yy2 + F.py_add * fb;
				sy1 = f2i(F.res_y2 + y2 / z2 - 1);
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

__device__ __forceinline__ void filter_ray_init(const TraverseParams& P, const RayInit& ri, FilterRay& F)
{
	F.ray_x = ri.ray_x; F.ray_z = ri.ray_z; F.rx2mr = ri.rx2mr; F.vertical = ri.vertical;
	F.sin_x = P.sin_x; F.cos_x = P.cos_x;
	F.vpx = P.viewpos[0]; F.mountain = P.viewpos[1]; F.vpz = P.viewpos[2];
	F.res_y2 = (float)(P.res_y / 2);                              // Cuda_Render.h:108 (integer division)
	F.pz_add = P.sin_x;                                           // pos3d_z_add (Cuda_Render.h:313)
	F.py_add = (ri.vertical ? P.cos_x : 0.0f) * ri.rx2mr;         // pos3d_y_add (Cuda_Render.h:314-315)
	Dda dd;
	dda_init(P, ri.ray_x, ri.ray_z, dd);
	F.fixx = dd.fixx; F.fixz = dd.fixz;
}

__device__ __forceinline__ void ray_ctx_init(const TraverseParams& P, const FilterRay& F, uint32_t* row, uint32_t* ymask, uint32_t* ids,
                                             long long* stat, int gl, RayCtx& R)
{
	R.row = row; R.ymask = ymask; R.ids = ids;
	R.res_y2 = F.res_y2; R.pz_add = F.pz_add; R.py_add = F.py_add; R.mountain = F.mountain; R.gl = gl;
	R.stat = stat;
	R.hc_on = (P.flags & 2) ? 1 : 0;
	R.hc = f2i((4095.0f - F.mountain) + P.viewpos[1]);            // int height_color = 4095-mountain+viewpos.y (Cuda_Render.h:675)
}

// PROF (tools/ray_profile.py only): per ray plane, clock64() cycles spent in each phase, written as
// unsigned long long[24] {total, dda, filter test + queue, geometry + gather, C1, C2, consume, steps | batches << 32}
// to the buffer passed in P.ids.
#define RLERC_TICK(slot) do { if (PROF) { const long long now_ = clock64(); prof[slot] += now_ - tick; tick = now_; } } while (0)

#ifndef RLERC_F_WPB
#define RLERC_F_WPB 8                       // ray planes (warps) per block (16 warps per SM at 128 registers); adjacent ray
                                            // planes share pointer-map lines in L1, 8 per block measured best at 4K
#endif

#ifndef RLERC_F_MINB
#define RLERC_F_MINB (16 / RLERC_F_WPB)       // resident blocks per SM the register allocation is capped for (128 registers)
#endif

// The traversal kernels are launched with programmatic stream serialization behind k_dda_states and never wait for it
// as a whole, but whatever follows in the stream — the next frame's k_dda_states into the same buffers — must not
// start while the old one still writes: ONE thread of the grid waits for the pre-pass before it leaves, which keeps the
// grid (not its other blocks) alive until then.  (Until round 2 every thread waited: a ray plane that looks at the
// ground is done in a third of the pre-pass's ~0.1 ms and then held its block's registers and shared memory idle —
// invisible in a full frame, a fifth of the time of the small slice launches of the multi-GPU mode.)
#define RLERC_EXIT_AFTER_PREPASS() do { if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)

__host__ __device__ inline int f_words_per_warp(int mask_words)
{
	return (RLERC_F_REC + RLERC_F_QUEUE + 16 + RLERC_PS_WORDS + mask_words + 3) & ~3;
}

template <bool IDS, bool PROF>
__global__ void __launch_bounds__(RLERC_F_WPB * 32, RLERC_F_MINB)
k_traverse_f(const __grid_constant__ TraverseParams P, int rays)
{
	long long prof[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	long long stat[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
	long long tick = PROF ? clock64() : 0;
	const long long t_begin = tick;
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int G = 32;
	constexpr int WPB = RLERC_F_WPB;
	const int gl = threadIdx.x & 31;
	const int wid = threadIdx.x >> 5;
	const unsigned FULL = 0xffffffffu;
	const unsigned lt_mask = (1u << gl) - 1u;

	const int ray_i = (int)blockIdx.x * WPB + wid;                  // launch-local ray index
	const int x = owned_ray(P, ray_i);
	if (ray_i >= rays || x >= P.ray_end) { RLERC_EXIT_AFTER_PREPASS(); return; }

	// shared per warp: crossing records of a chunk | live-column queue | DrawJob | RW x 32 projected runs (int2) |
	//                  RW x 32 deferred short spans | occlusion bits
	uint32_t* wbase = smem + (size_t)wid * f_words_per_warp(P.mask_words);
	const RecView rec = rec_view(wbase);
	uint32_t* queue = wbase + RLERC_F_REC;                           // [8][QCAP]
	DrawJob* job = reinterpret_cast<DrawJob*>(wbase + RLERC_F_REC + RLERC_F_QUEUE);
	int2* proj = reinterpret_cast<int2*>(wbase + RLERC_F_REC + RLERC_F_QUEUE + 16);
	uint32_t* shade = wbase + RLERC_F_REC + RLERC_F_QUEUE + 16 + RLERC_RW * 64;
	uint32_t* ymask = wbase + RLERC_F_REC + RLERC_F_QUEUE + 16 + RLERC_PS_WORDS;

	const int res_y = P.res_y;
	uint32_t* row = P.warp + (size_t)x * res_y;

	RayInit ri;
	ray_init(P, x, ri);
	clear_outside<G>(row, res_y, ri, gl);
	if (ri.skip) { RLERC_EXIT_AFTER_PREPASS(); return; }
	FilterRay F;
	filter_ray_init(P, ri, F);
	HorizonState Hs;
	Hs.ycmin = ri.ycmin; Hs.ycmax = ri.ycmax; Hs.hiw = 0;
	const int ymin0 = Hs.ycmin, ymax0 = Hs.ycmax;

	// occlusion mask clear; the sky sentinel is written at the end to the pixels that stayed
	// open (same final row as clear-then-overwrite, Cuda_Render.h:255-264)
	for (int w = gl; w < P.mask_words; w += G) ymask[w] = 0;
	__syncwarp();

	Counters Cn;
	memset(&Cn, 0, sizeof(Cn));
	RayCtx R;
	ray_ctx_init(P, F, row, ymask, (IDS && !PROF) ? P.ids + (size_t)x * res_y * 2 : nullptr, PROF ? stat : nullptr, gl, R);

	// DDA: this ray plane's saved states, the chunk in shared memory
	const int k_total = P.lod.k_total;
	const float2* const st = P.dda_states + (size_t)ray_i * ((k_total + 31) >> 5) * 3;
	const unsigned long long* const flag = P.dda_progress + ray_i;
	int published = 0;                                           // batches of this ray plane k_dda_states is known to have published
	float3 carry = make_float3(0.0f, 0.0f, 0.0f);                 // no crossing yet: distance 0, x-track (Cuda_Render.h:302-305)
	int chunk = -1, bl = RLERC_F_CH;                             // next batch: bl of chunk (bl == CH: the next chunk is due)
	int k_next = 0;                                              // first crossing of the next batch
	// filter: the batch whose pointer-map gather is in flight
	Geo fg;
	fg.pz = fg.py = fg.czz = fg.cyy = 0; fg.cmip = 0; fg.cidx = 0;
	unsigned fe0 = 0, fe1 = 0;
	bool fhave = false;
	int fn = 0;                        // crossings in that batch (0: none in flight)
	int qhead = 0, qcount = 0;         // live-column queue (uniform)
	// consume: the batch whose run words are in flight
	Stage s0;
	Geo g0;
	stage_clear(s0, g0);

	while (true)
	{
		if (Hs.ycmin >= Hs.ycmax) break;                         // Cuda_Render.h:370

		// ---- FILTER: until a full batch of live columns is queued (or the ray plane has reached z_far) -------------
		while (qcount < 32 && (k_next < k_total || fn > 0))
		{
			RLERC_TICK(0);
			if (PROF) prof[7] += 1;
			// F1. geometry of the next 32 crossings, conservative top clip, pointer-map gather; the gather of the
			//     batch before it (fg, fe0, fe1) lands meanwhile
			Geo ng;
			ng.pz = ng.py = ng.czz = ng.cyy = 0; ng.cmip = 0; ng.cidx = 0;
			unsigned ne0 = 0, ne1 = 0;
			bool nhave = false;
			int nn = 0;
			if (k_next < k_total)
			{
				if (bl == RLERC_F_CH)
				{
					dda_chunk(P, st, flag, published, F.ray_x, F.ray_z, ++chunk, gl, rec, carry);
					bl = 0;
					RLERC_TICK(1);
				}
				nn = k_total - k_next < G ? k_total - k_next : G;
				if (gl < nn) nhave = filter_geometry(P, F, rec, carry, bl, gl, Hs.ycmin, ng, ne0, ne1);
				bl++;
				k_next += nn;
				if (IDS && gl == 0) Cn.c_steps += nn;
			}
			RLERC_TICK(3);
			// F2. first-run test of the batch in flight, live columns -> queue
			if (fn > 0)
			{
				const bool live = gl < fn && fhave && filter_live<IDS>(F, fg, fe1, Hs.ycmin);
				const unsigned lb = __ballot_sync(FULL, live);
				if (live) column_put(queue + ((qhead + qcount + __popc(lb & lt_mask)) & (RLERC_QCAP - 1)), RLERC_QCAP, fg, fe0, fe1);
				qcount += __popc(lb);
			}
			fg = ng; fe0 = ne0; fe1 = ne1; fhave = nhave; fn = nn;
			__syncwarp();
			RLERC_TICK(2);
		}
		RLERC_TICK(0);

		// ---- C1. take the next batch of live columns off the queue, request their run words ---------------------
		Stage s1;
		Geo g1;
		stage_clear(s1, g1);
		{
			const int n1 = qcount < 32 ? qcount : 32;
			s1.nvalid = n1; s1.have = gl < n1;
			if (s1.have) column_take(P, queue + ((qhead + gl) & (RLERC_QCAP - 1)), RLERC_QCAP, g1, s1);
			qhead = (qhead + n1) & (RLERC_QCAP - 1);
			qcount -= n1;
		}
		__syncwarp();
		RLERC_TICK(4);

		// ---- C2. project the runs of batch s0 (their words were requested one round ago) -------------------------
		if (s0.nvalid > 0)
		{
			int slen = 0, nr = 0;
			bool longcol = false;
			unsigned flags = 0;
			if (s0.have) column_project(F, s0, g0, Hs.ycmin, gl, proj, slen, nr, longcol, flags);
			RLERC_TICK(5);
			if (PROF) prof[7] += 1ll << 32;
			// ---- B / B0 / S. consume batch s0 (traverse_common.cuh) ----------------------------------------------
			const bool finished = consume_batch<IDS, PROF>(P, R, Hs, Cn, s0, g0, slen, nr, longcol, flags, proj, shade, job);
			RLERC_TICK(6);
			if (finished) break;
		}
		else if (s1.nvalid == 0 && k_next >= k_total && fn == 0 && qcount == 0) break;   // drained (z > z_far, Cuda_Render.h:367)

		s0 = s1; g0 = g1;
	}
	__syncwarp();

	// sky sentinel on every pixel of the clip range that no run covered
	for (int y = ymin0 + gl; y <= ymax0; y += G)
		if (!((ymask[y >> 5] >> (y & 31)) & 1u)) st_warp(row + y, RLERC_SKY);

	if (IDS) flush_counters(P, Cn, gl, ymax0 - ymin0 + 1);
	if (PROF && gl == 0)
	{
		prof[0] = clock64() - t_begin;
		unsigned long long* out = reinterpret_cast<unsigned long long*>(P.ids) + (size_t)x * 24;
		for (int k = 0; k < 8; k++) out[k] = (unsigned long long)prof[k];
		for (int k = 0; k < 16; k++) out[8 + k] = (unsigned long long)stat[k];
	}
	RLERC_EXIT_AFTER_PREPASS();
}

static int launch_rays(const TraverseParams& p)
{
	return (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
}

// dynamic shared memory above 48 KB is an opt-in per kernel AND per device
template <typename K>
static void opt_in_smem(K kernel, size_t smem, size_t (&configured_on)[64])
{
	int dev = 0;
	cudaGetDevice(&dev);
	size_t& configured = configured_on[dev & 63];
	if (smem > configured)
	{
		cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = smem;
	}
}

size_t traverse_dda_state_bytes(const TraverseParams& p)
{
	const int rays = launch_rays(p);
	return (size_t)(rays > 0 ? rays : 0) * ((p.lod.k_total + 31) >> 5) * 3 * sizeof(float2);
}

size_t traverse_dda_progress_bytes(const TraverseParams& p)
{
	const int rays = launch_rays(p);
	return (size_t)(rays > 0 ? rays : 0) * sizeof(unsigned long long);
}

void launch_dda_states(const TraverseParams& p, cudaStream_t st)
{
	const int rays = launch_rays(p);
	if (rays <= 0 || p.lod.k_total <= 0) return;
	k_dda_states<<<(rays + 63) / 64, 64, 0, st>>>(p, rays);
}

// launch behind k_dda_states with programmatic stream serialization: the kernel may start while the pre-pass still runs
template <typename K>
static void launch_overlapped(K kernel, int blocks, int threads, size_t smem, cudaStream_t st, const TraverseParams& p, int rays)
{
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3((unsigned)threads);
	cfg.dynamicSmemBytes = smem; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	cudaLaunchKernelEx(&cfg, kernel, p, rays);
}

template <bool IDS, bool PROF>
static void launch_f(const TraverseParams& p, cudaStream_t st)
{
	const int wpb = RLERC_F_WPB;
	const int rays = launch_rays(p);
	if (rays <= 0) return;
	const int blocks = (rays + wpb - 1) / wpb;
	const size_t smem = (size_t)wpb * f_words_per_warp(p.mask_words) * sizeof(uint32_t);
	static size_t configured_on[64] = { 0 };
	opt_in_smem(k_traverse_f<IDS, PROF>, smem, configured_on);
	launch_overlapped(k_traverse_f<IDS, PROF>, blocks, wpb * 32, smem, st, p, rays);
}


// ================================================================================================================
// k_traverse_p (variant 68; picked automatically for launches of at most 12 ray planes per SM, capi.cu pick_lanes):
// the same algorithm with the two loops of k_traverse_f on TWO warps per
// ray plane.  Warp F runs the FILTER (DDA chunks, pointer-map gather, first-run test) and pushes the live columns into
// a ring in shared memory; warp C pops batches of 32, loads and projects their runs and runs consume_batch.  A ray
// plane's chain becomes max(filter, consume) instead of their sum.  The filter only ever drops provable no-ops under
// a horizon that can only rise, so a stale y_clip_min (published by C after every batch) lets more columns through
// but cannot change the result: the picture does not depend on timing.
// Protocol (volatile words in shared memory, one writer each): tail (F), head (C), ycmin (C), done (F), closed (C).
// F waits while the ring has no room for 32 more entries, C waits until 32 entries are there or F is done; both
// waits are bounded (a broken protocol traps after >= 1 s of polling: a launch error, neither a hang nor a wrong picture).  Waiting is __nanosleep(64)
// polling by the whole warp.  Tried instead in round 1, all bit-exact, all slower: back-off polling, mbarrier.try_wait as
// "sleep until the partner signals", lane 0 polling alone with the other lanes parked at a shuffle.
#ifndef RLERC_P_RCAP
#define RLERC_P_RCAP 128
#endif
//                    // ring capacity in columns (power of two, >= 64)
#define RLERC_P_RING (8 * RLERC_P_RCAP)     // words: 8 fields x RCAP, field-major
#define RLERC_P_PLANES 4                    // ray planes per block (8 warps: F0..F3 | C0..C3, roles by warpgroup)
#define RLERC_P_SPIN_MAX (1 << 24)            // >= 1 s of polling: the protocol is broken; fail loudly (launch error), never a wrong picture

// Register re-balancing between the roles (setmaxnreg, sm_90+): launch with a cap that admits 3 blocks per SM
// (84 registers), the filter warpgroup gives registers back, the consume warpgroup takes them: 24 warps and 12 ray
// planes per SM instead of 16 and 8.  RLERC_P_REG_F = 0 switches it off (2 blocks per SM, 128 registers for both roles).
#ifndef RLERC_P_REG_F
#define RLERC_P_REG_F 48
#endif
#ifndef RLERC_P_REG_C
#define RLERC_P_REG_C 112
#endif
// The pool is what the block was LAUNCHED with: ptxas rounds the 3-blocks-per-SM cap (85) down to 80 registers per
// thread, so the consume side can take at most 2 * 80 - RLERC_P_REG_F; asking for more spins in USETMAXREG.TRY_ALLOC
// forever (measured the hard way).
static_assert(RLERC_P_REG_F == 0 || RLERC_P_REG_F + RLERC_P_REG_C <= 160, "setmaxnreg: the consume warpgroup cannot take more registers than the filter warpgroup gives back");
#define RLERC_STR2(x) #x
#define RLERC_STR(x) RLERC_STR2(x)

__device__ __noinline__ void pair_protocol_broken() { __trap(); }     // cold: keeps the trap out of the wait loops' code

__host__ __device__ inline int p_words_per_plane(int mask_words)
{
	return (RLERC_F_REC + RLERC_P_RING + 8 + 16 + RLERC_PS_WORDS + mask_words + 3) & ~3;
}

__global__ void __launch_bounds__(RLERC_P_PLANES * 64, RLERC_P_REG_F ? 3 : 2)
k_traverse_p(const __grid_constant__ TraverseParams P, int rays)
{
	constexpr bool IDS = false, PROF = false;
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int G = 32;
	const int gl = threadIdx.x & 31;
	const int wid = threadIdx.x >> 5;
	const bool roleC = wid >= RLERC_P_PLANES;
	const int pl = roleC ? wid - RLERC_P_PLANES : wid;               // ray plane of this warp inside the block
	const unsigned FULL = 0xffffffffu;
	const unsigned lt_mask = (1u << gl) - 1u;

	const int ray_i = (int)blockIdx.x * RLERC_P_PLANES + pl;         // launch-local ray index
	// no early return before setmaxnreg (every warp of a warpgroup has to execute it): an out-of-range pair
	// computes on a valid dummy ray plane and stays inactive
	const bool inrange = ray_i < rays && owned_ray(P, ray_i) < P.ray_end;
	const int x = inrange ? owned_ray(P, ray_i) : P.ray_begin;

	// shared per ray plane: crossing records | ring | ctl | DrawJob | projections + span records | occlusion bits
	uint32_t* wbase = smem + (size_t)pl * p_words_per_plane(P.mask_words);
	const RecView rec = rec_view(wbase);
	uint32_t* ring = wbase + RLERC_F_REC;                            // [8][RCAP]
	volatile int* ctl = reinterpret_cast<volatile int*>(wbase + RLERC_F_REC + RLERC_P_RING);   // tail, head, ycmin, done, closed
	DrawJob* job = reinterpret_cast<DrawJob*>(wbase + RLERC_F_REC + RLERC_P_RING + 8);
	int2* proj = reinterpret_cast<int2*>(wbase + RLERC_F_REC + RLERC_P_RING + 8 + 16);
	uint32_t* shade = wbase + RLERC_F_REC + RLERC_P_RING + 8 + 16 + RLERC_RW * 64;
	uint32_t* ymask = wbase + RLERC_F_REC + RLERC_P_RING + 8 + 16 + RLERC_PS_WORDS;

	const int res_y = P.res_y;
	uint32_t* row = P.warp + (size_t)x * res_y;
	RayInit ri;
	ray_init(P, x, ri);
	FilterRay F;
	filter_ray_init(P, ri, F);
	const int ymin0 = ri.ycmin, ymax0 = ri.ycmax;

	// control words: C initialises, the named barrier below (both warps of the pair) publishes them
	if (roleC)
	{
		if (inrange) clear_outside<G>(row, res_y, ri, gl);
		for (int w = gl; w < P.mask_words; w += G) ymask[w] = 0;
		if (gl == 0) { ctl[0] = 0; ctl[1] = 0; ctl[2] = ri.ycmin; ctl[3] = 0; ctl[4] = (ri.skip || !inrange) ? 1 : 0; }
	}
	asm volatile("bar.sync %0, 64;" :: "r"(1 + pl) : "memory");     // named barrier per ray plane: 2 warps
	const bool active = inrange && !ri.skip;

	if (!roleC)
	{
#if RLERC_P_REG_F
		asm volatile("setmaxnreg.dec.sync.aligned.u32 " RLERC_STR(RLERC_P_REG_F) ";");
#endif
		if (!active) { RLERC_EXIT_AFTER_PREPASS(); return; }
		// ================= warp F: DDA chunks -> geometry + gather -> first-run test -> ring ====================
		const int k_total = P.lod.k_total;
		const float2* const st = P.dda_states + (size_t)ray_i * ((k_total + 31) >> 5) * 3;
		const unsigned long long* const flag = P.dda_progress + ray_i;
		int published = 0;
		float3 carry = make_float3(0.0f, 0.0f, 0.0f);
		int chunk = -1, bl = RLERC_F_CH, k_next = 0;
		Geo fg;
		fg.pz = fg.py = fg.czz = fg.cyy = 0; fg.cmip = 0; fg.cidx = 0;
		unsigned fe0 = 0, fe1 = 0;
		bool fhave = false;
		int fn = 0;
		int tail = 0;
		bool closed = false;
		while (k_next < k_total || fn > 0)
		{
			// room for 32 more entries?  (head only grows)
			// (lane 0 reads the control words and broadcasts them: every decision on them is warp-uniform)
			int spins = 0, ycmin = 0;
			while (true)
			{
				int h = 0, c = 0, y = 0;
				if (gl == 0) { h = ctl[1]; c = ctl[4]; y = ctl[2]; }
				h = __shfl_sync(FULL, h, 0); c = __shfl_sync(FULL, c, 0); ycmin = __shfl_sync(FULL, y, 0);
				if (c) { closed = true; break; }
				if (tail + 32 - h <= RLERC_P_RCAP) break;
				__nanosleep(64);
				if (++spins > RLERC_P_SPIN_MAX) { pair_protocol_broken(); if (gl == 0) ctl[4] = 1; closed = true; break; }
			}
			if (closed) break;                                         // ycmin: possibly stale, lower than the truth, never higher
			Geo ng;
			ng.pz = ng.py = ng.czz = ng.cyy = 0; ng.cmip = 0; ng.cidx = 0;
			unsigned ne0 = 0, ne1 = 0;
			bool nhave = false;
			int nn = 0;
			if (k_next < k_total)
			{
				if (bl == RLERC_F_CH) { dda_chunk(P, st, flag, published, F.ray_x, F.ray_z, ++chunk, gl, rec, carry); bl = 0; }
				nn = k_total - k_next < G ? k_total - k_next : G;
				if (gl < nn) nhave = filter_geometry(P, F, rec, carry, bl, gl, ycmin, ng, ne0, ne1);
				bl++;
				k_next += nn;
			}
			if (fn > 0)
			{
				const bool live = gl < fn && fhave && filter_live<false>(F, fg, fe1, ycmin);
				const unsigned lb = __ballot_sync(FULL, live);
				if (live) column_put(ring + ((tail + __popc(lb & lt_mask)) & (RLERC_P_RCAP - 1)), RLERC_P_RCAP, fg, fe0, fe1);
				tail += __popc(lb);
				__threadfence_block();
				__syncwarp();
				if (gl == 0 && lb) ctl[0] = tail;
			}
			fg = ng; fe0 = ne0; fe1 = ne1; fhave = nhave; fn = nn;
			__syncwarp();
		}
		__threadfence_block();
		__syncwarp();
		if (gl == 0 && !closed) { ctl[0] = tail; __threadfence_block(); ctl[3] = 1; }
		RLERC_EXIT_AFTER_PREPASS();
		return;
	}

	// ===================== warp C: ring -> run words -> projection -> consume_batch ============================
#if RLERC_P_REG_F
	asm volatile("setmaxnreg.inc.sync.aligned.u32 " RLERC_STR(RLERC_P_REG_C) ";");
#endif
	if (!active) { RLERC_EXIT_AFTER_PREPASS(); return; }
	HorizonState Hs;
	Hs.ycmin = ri.ycmin; Hs.ycmax = ri.ycmax; Hs.hiw = 0;
	Counters Cn;
	memset(&Cn, 0, sizeof(Cn));
	RayCtx R;
	ray_ctx_init(P, F, row, ymask, nullptr, nullptr, gl, R);
	Geo g0;
	Stage s0;
	stage_clear(s0, g0);
	int head = 0;
	while (true)
	{
		if (Hs.ycmin >= Hs.ycmax) break;
		// C1. wait for a full batch (or the end of the ray plane), pop it, request its run words
		int done = 0, avail = 0, spins = 0;
		while (true)
		{
			int d = 0, t = 0;
			if (gl == 0)
			{
				d = ctl[3];                                           // read BEFORE tail: done implies the final tail is visible
				__threadfence_block();
				t = ctl[0];
			}
			done = __shfl_sync(FULL, d, 0);
			avail = __shfl_sync(FULL, t, 0) - head;
			if (avail >= 32 || done) break;
			__nanosleep(64);
			if (++spins > RLERC_P_SPIN_MAX) { pair_protocol_broken(); done = 1; break; }
		}
		__syncwarp();
		Stage s1;
		Geo g1;
		stage_clear(s1, g1);
		{
			const int n1 = avail < 32 ? avail : 32;
			s1.nvalid = n1; s1.have = gl < n1;
			if (s1.have) column_take(P, ring + ((head + gl) & (RLERC_P_RCAP - 1)), RLERC_P_RCAP, g1, s1);
			head += n1;
			__syncwarp();
			if (gl == 0 && n1) ctl[1] = head;                          // the slots may be overwritten from now on
		}
		// C2. project the runs of batch s0
		if (s0.nvalid > 0)
		{
			int slen = 0, nr = 0;
			bool longcol = false;
			unsigned flags = 0;
			if (s0.have) column_project(F, s0, g0, Hs.ycmin, gl, proj, slen, nr, longcol, flags);
			const bool finished = consume_batch<IDS, PROF>(P, R, Hs, Cn, s0, g0, slen, nr, longcol, flags, proj, shade, job);
			if (gl == 0) ctl[2] = Hs.ycmin;                            // the filter's (stale) horizon
			if (finished) break;
		}
		else if (s1.nvalid == 0 && done) break;                        // drained
		s0 = s1; g0 = g1;
	}
	__syncwarp();
	if (gl == 0) ctl[4] = 1;                                           // releases a filter warp that is still running
	for (int y = ymin0 + gl; y <= ymax0; y += G)
		if (!((ymask[y >> 5] >> (y & 31)) & 1u)) st_warp(row + y, RLERC_SKY);
	RLERC_EXIT_AFTER_PREPASS();
}

void launch_traverse_pair(const TraverseParams& p, cudaStream_t st)
{
	const int rays = launch_rays(p);
	if (rays <= 0) return;
	const int blocks = (rays + RLERC_P_PLANES - 1) / RLERC_P_PLANES;
	const size_t smem = (size_t)RLERC_P_PLANES * p_words_per_plane(p.mask_words) * sizeof(uint32_t);
	static size_t configured_on[64] = { 0 };
	opt_in_smem(k_traverse_p, smem, configured_on);
	launch_overlapped(k_traverse_p, blocks, RLERC_P_PLANES * 64, smem, st, p, rays);
}

// ================================================================================================================
// k_traverse_q (variant 69): FOUR warps per ray plane, one per stage of the traversal, for launches that are bound by the
// serial chain of their longest ray planes (single frames, multi-GPU slices).  Measured in round 2 on the sky ray planes of
// a 4K frame: of the ~1.3 M cycles of such a chain the filter is 30 %, run loads + projection 16 %, the occlusion
// machinery 28 % and the deferred shading 26 %; only the machinery (which row goes to which span, the horizon) is a true
// recurrence from column to column.  So:
//   F  filter    DDA chunks -> geometry + pointer-map gather -> first-run test -> live columns into a ring   (as warp F of k_traverse_p)
//   P  project   per 32 live columns: run-word loads (one batch early), projection of up to RW runs to screen rows
//   R  resolve   resolve_batch (traverse_common.cuh): rising-horizon / ownership-resolved / event-loop paths, long spans; owns
//                the occlusion mask and the horizon, publishes y_clip_min for F and P
//   S  shade     shade_batch: the short spans R assigned (interpolants, attribute gathers, pixel stores)
// The chain of a ray plane becomes max(F, P, R, S) instead of their sum.  F and P work under a STALE horizon (the last
// one R published): the filter only drops columns that are no-ops under any later horizon and the projection only stops
// at a run that breaks under any later horizon, R re-tests everything under the exact state — the picture does not
// depend on timing (same argument as k_traverse_p).  Batches travel through three buffers per ray plane that go round
// P -> R -> S -> P; hand-over is by mbarrier (arrive.release / try_wait.acquire, one arrival per hand-over after a
// __syncwarp), so a waiting warp sleeps in hardware instead of polling.  Roles are by warpgroup (4 ray planes per block of
// 16 warps) so that setmaxnreg can move registers to where they are needed: launch cap 64, F 40, P 56, R 104, S 56.
#define RLERC_Q_CH 8                        // DDA chunk of the filter warp (8 x 32 crossings: half the record area of k_traverse_f)
#define RLERC_Q_RING 128                    // live-column ring, four quarters of 32 (one hand-over each)
#define RLERC_Q_BUFS 3
#define RLERC_Q_PLANES 4
#define RLERC_Q_STAGE 12                    // words per lane of a batch buffer's column record
#define RLERC_Q_DEPTH 0                     // (stash of a deeper gather pipeline in the filter warp: measured slower, not used)
#define RLERC_Q_BUF_WORDS (RLERC_Q_STAGE * 32 + RLERC_PS_WORDS + 8)      // column records | projections + span records | header
#ifndef RLERC_Q_REG_F
#define RLERC_Q_REG_F 64
#endif
#ifndef RLERC_Q_REG_P
#define RLERC_Q_REG_P 48
#endif
#ifndef RLERC_Q_REG_R
#define RLERC_Q_REG_R 96
#endif
#ifndef RLERC_Q_REG_S
#define RLERC_Q_REG_S 48
#endif
static_assert(RLERC_Q_REG_F + RLERC_Q_REG_P + RLERC_Q_REG_R + RLERC_Q_REG_S <= 4 * 64, "setmaxnreg: the roles share the block's launch allocation");
#define RLERC_Q_SPIN_MAX (1 << 21)          // try_wait rounds (each blocks for up to the hardware's time limit): the protocol is broken; fail loudly

__device__ __forceinline__ void mbar_init(uint64_t* b, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
	asm volatile("{ .reg .b64 t; mbarrier.arrive.release.cta.shared::cta.b64 t, [%0]; }" :: "r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __noinline__ void quad_protocol_broken() { __trap(); }
// all lanes wait for the completion of the phase with this parity
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity)
{
	const uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
	for (int spins = 0;; spins++)
	{
		unsigned ok;
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
		             : "=r"(ok) : "r"(a), "r"(parity) : "memory");
		if (ok) return;
		if (spins > RLERC_Q_SPIN_MAX) { quad_protocol_broken(); return; }
	}
}

// RLERC_Q_PROF (tools/quad_profile.py, an A/B build only): per ray plane and role {cycles from start to exit, cycles spent in
// mbar_wait, hand-overs}, unsigned long long[24] per ray plane in the buffer passed as P.ids: [role * 4 + {0, 1, 2}]
#ifndef RLERC_Q_PROF
#define RLERC_Q_PROF 0
#endif
#if RLERC_Q_PROF
#define QPROF_DECL long long qp_t0 = clock64(), qp_wait = 0, qp_n = 0
#define QPROF_WAIT(stmt) do { const long long a_ = clock64(); stmt; qp_wait += clock64() - a_; } while (0)
#define QPROF_ITEM() (qp_n++)
#define QPROF_END(role_) do { if (gl == 0 && P.ids) { unsigned long long* o_ = reinterpret_cast<unsigned long long*>(P.ids) + (size_t)x * 24 + (role_) * 4; \
	o_[0] = (unsigned long long)(clock64() - qp_t0); o_[1] = (unsigned long long)qp_wait; o_[2] = (unsigned long long)qp_n; } } while (0)
#else
#define QPROF_DECL
#define QPROF_WAIT(stmt) stmt
#define QPROF_ITEM()
#define QPROF_END(role_)
#endif

struct QPlane {                              // shared memory of one ray plane (word offsets from its base)
	uint32_t* base;
	int mask_words;
	__device__ __forceinline__ uint32_t* rec() const { return base; }
	__device__ __forceinline__ uint32_t* ring() const { return base + rec_words(RLERC_Q_CH); }                     // [8][RING]
	__device__ __forceinline__ uint32_t* buf(int i) const { return ring() + 8 * RLERC_Q_RING + i * RLERC_Q_BUF_WORDS; }
	__device__ __forceinline__ DrawJob* job() const { return reinterpret_cast<DrawJob*>(buf(RLERC_Q_BUFS)); }
	__device__ __forceinline__ volatile int* ctl() const { return reinterpret_cast<volatile int*>(buf(RLERC_Q_BUFS) + 16); }   // ycmin, closed, f_total
	__device__ __forceinline__ uint64_t* bars() const { return reinterpret_cast<uint64_t*>(buf(RLERC_Q_BUFS) + 16 + 8); }     // 8-byte aligned
	__device__ __forceinline__ uint32_t* stash() const { return buf(RLERC_Q_BUFS) + 16 + 8 + 2 * 20; }                       // [DEPTH][5][32]
	__device__ __forceinline__ uint32_t* ymask() const { return stash() + RLERC_Q_DEPTH * 160; }
};
// mbarriers of a ray plane: ring_full[4] ring_empty[4] p_full[3] r_full[3] s_free[3]
#define QB_RING_FULL 0
#define QB_RING_EMPTY 4
#define QB_P_FULL 8
#define QB_R_FULL 11
#define QB_S_FREE 14
#define QB_COUNT 17

__host__ __device__ inline int q_words_per_plane(int mask_words)
{
	return (rec_words(RLERC_Q_CH) + 8 * RLERC_Q_RING + RLERC_Q_BUFS * RLERC_Q_BUF_WORDS + 16 + 8 + 2 * 20 + RLERC_Q_DEPTH * 160 + mask_words + 3) & ~3;
}

// column record of a batch buffer, field-major: e0 rw0 rw1 rw2 rw3 pz py czz cyy meta slen shade_runs
// meta = cmip | nr << 8 | longcol << 12 | have << 13 | flags << 16
__device__ __forceinline__ void qbuf_put(uint32_t* b, int gl, const Stage& s, const Geo& g, int slen, int nr, bool longcol, unsigned flags)
{
	b[0 * 32 + gl] = s.e0;
	b[1 * 32 + gl] = s.rw[0]; b[2 * 32 + gl] = s.rw[1]; b[3 * 32 + gl] = s.rw[2]; b[4 * 32 + gl] = s.rw[3];
	b[5 * 32 + gl] = __float_as_uint(g.pz); b[6 * 32 + gl] = __float_as_uint(g.py);
	b[7 * 32 + gl] = __float_as_uint(g.czz); b[8 * 32 + gl] = __float_as_uint(g.cyy);
	b[9 * 32 + gl] = (uint32_t)g.cmip | ((uint32_t)nr << 8) | ((longcol ? 1u : 0u) << 12) | ((s.have ? 1u : 0u) << 13) | (flags << 16);
	b[10 * 32 + gl] = (uint32_t)slen;
}
__device__ __forceinline__ void qbuf_get(const uint32_t* b, int gl, Stage& s, Geo& g, int& slen, int& nr, bool& longcol, unsigned& flags)
{
	s.e0 = b[0 * 32 + gl];
	s.rw[0] = b[1 * 32 + gl]; s.rw[1] = b[2 * 32 + gl]; s.rw[2] = b[3 * 32 + gl]; s.rw[3] = b[4 * 32 + gl];
	g.pz = __uint_as_float(b[5 * 32 + gl]); g.py = __uint_as_float(b[6 * 32 + gl]);
	g.czz = __uint_as_float(b[7 * 32 + gl]); g.cyy = __uint_as_float(b[8 * 32 + gl]);
	const uint32_t meta = b[9 * 32 + gl];
	g.cmip = (int)(meta & 255u); g.cidx = 0;
	nr = (int)((meta >> 8) & 15u); longcol = (meta >> 12) & 1u; s.have = (meta >> 13) & 1u; flags = meta >> 16;
	slen = (int)b[10 * 32 + gl];
	s.e1 = (uint32_t)slen | (s.rw[0] << 16);                         // n_runs | first run << 16 (column_project left run 0 in the low half of rw[0])
}

__global__ void __launch_bounds__(RLERC_Q_PLANES * 128, 2)
k_traverse_q(const __grid_constant__ TraverseParams P, int rays)
{
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int G = 32;
	const int gl = threadIdx.x & 31;
	const int wid = threadIdx.x >> 5;
	const int role = wid / RLERC_Q_PLANES;                           // 0 F, 1 P, 2 R, 3 S (one warpgroup each)
	const int pl = wid % RLERC_Q_PLANES;
	const unsigned FULL = 0xffffffffu;
	const unsigned lt_mask = (1u << gl) - 1u;

	const int ray_i = (int)blockIdx.x * RLERC_Q_PLANES + pl;         // launch-local ray index
	// no early return before setmaxnreg (every warp of a warpgroup has to execute it)
	const bool inrange = ray_i < rays && owned_ray(P, ray_i) < P.ray_end;
	const int x = inrange ? owned_ray(P, ray_i) : P.ray_begin;

	QPlane Q;
	Q.base = smem + (size_t)pl * q_words_per_plane(P.mask_words);
	Q.mask_words = P.mask_words;
	volatile int* const ctl = Q.ctl();
	uint64_t* const bars = Q.bars();
	uint32_t* const ymask = Q.ymask();

	const int res_y = P.res_y;
	uint32_t* row = P.warp + (size_t)x * res_y;
	RayInit ri;
	ray_init(P, x, ri);
	const bool active = inrange && !ri.skip;

	if (role == 2)
	{
		if (inrange) clear_outside<G>(row, res_y, ri, gl);
		for (int w = gl; w < P.mask_words; w += G) ymask[w] = 0;
		if (gl == 0)
		{
			ctl[0] = ri.ycmin; ctl[1] = 0; ctl[2] = -1;
			for (int k = 0; k < QB_COUNT; k++) mbar_init(bars + k, 1);
		}
	}
	__syncthreads();

	if (role == 0)
	{
		// ================= F: DDA chunks -> geometry + gather -> first-run test -> ring =========================
		asm volatile("setmaxnreg.dec.sync.aligned.u32 " RLERC_STR(RLERC_Q_REG_F) ";");
		if (!active) { RLERC_EXIT_AFTER_PREPASS(); return; }
		FilterRay F;
		filter_ray_init(P, ri, F);
		const RecView rec = rec_view<RLERC_Q_CH>(Q.rec());
		uint32_t* const ring = Q.ring();
		const int k_total = P.lod.k_total;
		const float2* const st = P.dda_states + (size_t)ray_i * ((k_total + 31) >> 5) * 3;
		const unsigned long long* const flag = P.dda_progress + ray_i;
		int published = 0;
		float3 carry = make_float3(0.0f, 0.0f, 0.0f);
		int chunk = -1, bl = RLERC_Q_CH, k_next = 0;
		// one batch of 32 crossings is in flight between its pointer-map gather and the first-run test that needs the result
		// (a deeper pipeline was measured: the filter warp is bound by its own instruction stream, ~400 dependent instructions
		// per step, not by the gather; four batches in flight made a step 20 % slower)
		Geo fg;
		fg.pz = fg.py = fg.czz = fg.cyy = 0; fg.cmip = 0; fg.cidx = 0;
		unsigned fe0 = 0, fe1 = 0;
		bool fhave = false;
		int fn = 0;
		int tail = 0;                       // entries written
		int acquired = 0, filled = 0;       // quarters this warp may write / has handed over
		QPROF_DECL;
		while (k_next < k_total || fn > 0)
		{
			if (ctl[1]) break;                                           // R closed the ray plane
			const int ycmin = ctl[0];                                    // possibly stale: lower than the truth, never higher
			QPROF_ITEM();
			// room for 32 more entries
			while (((tail + 63) >> 5) > acquired)
			{
				QPROF_WAIT(mbar_wait(bars + QB_RING_EMPTY + (acquired & 3), ((acquired >> 2) & 1) ^ 1));
				acquired++;
			}
			Geo ng;
			ng.pz = ng.py = ng.czz = ng.cyy = 0; ng.cmip = 0; ng.cidx = 0;
			unsigned ne0 = 0, ne1 = 0;
			bool nhave = false;
			int nn = 0;
			if (k_next < k_total)
			{
				if (bl == RLERC_Q_CH) { dda_chunk<RLERC_Q_CH>(P, st, flag, published, F.ray_x, F.ray_z, ++chunk, gl, rec, carry); bl = 0; }
				nn = k_total - k_next < G ? k_total - k_next : G;
				if (gl < nn) nhave = filter_geometry(P, F, rec, carry, bl, gl, ycmin, ng, ne0, ne1);
				bl++;
				k_next += nn;
			}
			if (fn > 0)
			{
				const bool live = gl < fn && fhave && filter_live<false>(F, fg, fe1, ycmin);
				const unsigned lb = __ballot_sync(FULL, live);
				if (live)
				{
					column_put(ring + ((tail + __popc(lb & lt_mask)) & (RLERC_Q_RING - 1)), RLERC_Q_RING, fg, fe0, fe1);
					// P will want the column's run words: start them on their way into this SM's L1 now
					if ((fe1 & 0xffffu) > 1u)
						asm volatile("prefetch.global.L1 [%0];" :: "l"(P.level[fg.cmip].slabs + 2 + (size_t)fe0 + 1));
				}
				tail += __popc(lb);
				__syncwarp();
				while ((filled + 1) * 32 <= tail)
				{
					if (gl == 0) mbar_arrive(bars + QB_RING_FULL + (filled & 3));
					filled++;
				}
			}
			fg = ng; fe0 = ne0; fe1 = ne1; fhave = nhave; fn = nn;
		}
		// end of the stream: the total, then one more hand-over (the last, partial or empty, quarter); its barrier must
		// not be signalled before P has consumed the quarter that used it last
		while (acquired <= filled)
		{
			QPROF_WAIT(mbar_wait(bars + QB_RING_EMPTY + (acquired & 3), ((acquired >> 2) & 1) ^ 1));
			acquired++;
		}
		__syncwarp();
		if (gl == 0)
		{
			ctl[2] = tail;
			mbar_arrive(bars + QB_RING_FULL + (filled & 3));
		}
		QPROF_END(0);
		RLERC_EXIT_AFTER_PREPASS();
		return;
	}

	if (role == 1)
	{
		// ================= P: ring -> run words -> projection -> batch buffer ==================================
		asm volatile("setmaxnreg.dec.sync.aligned.u32 " RLERC_STR(RLERC_Q_REG_P) ";");
		if (!active) { RLERC_EXIT_AFTER_PREPASS(); return; }
		FilterRay F;
		filter_ray_init(P, ri, F);
		const uint32_t* const ring = Q.ring();
		Stage s0;
		Geo g0;
		stage_clear(s0, g0);
		bool s0_valid = false, s0_last = false;
		bool ended = false;
		int k = 0, b = 0;
		QPROF_DECL;
		while (true)
		{
			// C1: the next quarter of the ring
			Stage s1;
			Geo g1;
			stage_clear(s1, g1);
			bool s1_valid = false, s1_last = false;
			if (!ended)
			{
				QPROF_WAIT(mbar_wait(bars + QB_RING_FULL + (k & 3), (k >> 2) & 1));
				const int ft = ctl[2];
				int n = 32;
				if (ft >= 0 && ft - k * 32 < 32) { n = ft - k * 32; s1_last = true; ended = true; }
				s1.nvalid = n; s1.have = gl < n;
				if (s1.have) column_take(P, ring + ((k * 32 + gl) & (RLERC_Q_RING - 1)), RLERC_Q_RING, g1, s1);
				s1_valid = true;
				__syncwarp();
				if (gl == 0) mbar_arrive(bars + QB_RING_EMPTY + (k & 3));
				k++;
			}
			// C2: project s0 into the next batch buffer
			if (s0_valid)
			{
				QPROF_WAIT(mbar_wait(bars + QB_S_FREE + (b % RLERC_Q_BUFS), ((b / RLERC_Q_BUFS) & 1) ^ 1));
				QPROF_ITEM();
				uint32_t* const bw = Q.buf(b % RLERC_Q_BUFS);
				int2* const proj = reinterpret_cast<int2*>(bw + RLERC_Q_STAGE * 32);
				int slen = 0, nr = 0;
				bool longcol = false;
				unsigned flags = 0;
				const bool closed = ctl[1] != 0;
				if (!closed && s0.have) column_project(F, s0, g0, ctl[0], gl, proj, slen, nr, longcol, flags);
				qbuf_put(bw, gl, s0, g0, slen, nr, longcol, flags);
				if (gl == 0) { bw[RLERC_Q_STAGE * 32 + RLERC_PS_WORDS] = closed ? 0u : (uint32_t)s0.nvalid; bw[RLERC_Q_STAGE * 32 + RLERC_PS_WORDS + 1] = s0_last ? 1u : 0u; }
				__syncwarp();
				if (gl == 0) mbar_arrive(bars + QB_P_FULL + (b % RLERC_Q_BUFS));
				b++;
				if (s0_last) break;
			}
			s0 = s1; g0 = g1; s0_valid = s1_valid; s0_last = s1_last;
		}
		QPROF_END(1);
		RLERC_EXIT_AFTER_PREPASS();
		return;
	}

	if (role == 2)
	{
		// ================= R: resolve — the occlusion state of the ray plane ===================================
		asm volatile("setmaxnreg.inc.sync.aligned.u32 " RLERC_STR(RLERC_Q_REG_R) ";");
		if (!active) { RLERC_EXIT_AFTER_PREPASS(); return; }
		FilterRay F;
		filter_ray_init(P, ri, F);
		const int ymin0 = ri.ycmin, ymax0 = ri.ycmax;
		HorizonState Hs;
		Hs.ycmin = ri.ycmin; Hs.ycmax = ri.ycmax; Hs.hiw = 0;
		Counters Cn;
		memset(&Cn, 0, sizeof(Cn));
		RayCtx R;
		ray_ctx_init(P, F, row, ymask, nullptr, nullptr, gl, R);
		bool finished = false;
		QPROF_DECL;
		for (int b = 0;; b++)
		{
			QPROF_WAIT(mbar_wait(bars + QB_P_FULL + (b % RLERC_Q_BUFS), (b / RLERC_Q_BUFS) & 1));
			QPROF_ITEM();
			uint32_t* const bw = Q.buf(b % RLERC_Q_BUFS);
			const int nvalid = (int)bw[RLERC_Q_STAGE * 32 + RLERC_PS_WORDS];
			const bool last = bw[RLERC_Q_STAGE * 32 + RLERC_PS_WORDS + 1] != 0;
			unsigned shade_runs = 0;
			if (nvalid > 0 && !finished)
			{
				const int2* const proj = reinterpret_cast<const int2*>(bw + RLERC_Q_STAGE * 32);
				uint32_t* const shade = bw + RLERC_Q_STAGE * 32 + RLERC_RW * 64;
				Stage s0;
				Geo g0;
				int slen, nr;
				bool longcol;
				unsigned flags;
				qbuf_get(bw, gl, s0, g0, slen, nr, longcol, flags);
				s0.nvalid = nvalid;
				// P projected under an older horizon: a run that breaks under the current one ends the column here
				// (Cuda_Render.h:542-543), which also settles a column P had to leave open as "long"
				if (s0.have)
					for (int r = 0; r < nr; r++)
						if (((flags >> r) & 1u) && proj[r * 32 + gl].y <= Hs.ycmin) { nr = r + 1; longcol = false; break; }
				finished = resolve_batch<false, false>(P, R, Hs, Cn, s0, g0, slen, nr, longcol, flags, proj, shade, Q.job(), shade_runs);
				if (Hs.ycmin >= Hs.ycmax) finished = true;               // Cuda_Render.h:370
				if (gl == 0) { ctl[0] = Hs.ycmin; if (finished) ctl[1] = 1; }
			}
			bw[11 * 32 + gl] = shade_runs;
			__syncwarp();
			if (gl == 0) mbar_arrive(bars + QB_R_FULL + (b % RLERC_Q_BUFS));
			if (last) break;
		}
		__syncwarp();
		// sky sentinel on every pixel of the clip range that no run covered
		for (int y = ymin0 + gl; y <= ymax0; y += G)
			if (!((ymask[y >> 5] >> (y & 31)) & 1u)) st_warp(row + y, RLERC_SKY);
		QPROF_END(2);
		RLERC_EXIT_AFTER_PREPASS();
		return;
	}

	// ===================== S: shade the short spans R assigned ===================================================
	asm volatile("setmaxnreg.dec.sync.aligned.u32 " RLERC_STR(RLERC_Q_REG_S) ";");
	if (!active) { RLERC_EXIT_AFTER_PREPASS(); return; }
	{
		FilterRay F;
		filter_ray_init(P, ri, F);
		Counters Cn;
		memset(&Cn, 0, sizeof(Cn));
		RayCtx R;
		ray_ctx_init(P, F, row, ymask, nullptr, nullptr, gl, R);
		QPROF_DECL;
		for (int b = 0;; b++)
		{
			QPROF_WAIT(mbar_wait(bars + QB_R_FULL + (b % RLERC_Q_BUFS), (b / RLERC_Q_BUFS) & 1));
			QPROF_ITEM();
			uint32_t* const bw = Q.buf(b % RLERC_Q_BUFS);
			const bool last = bw[RLERC_Q_STAGE * 32 + RLERC_PS_WORDS + 1] != 0;
			const unsigned shade_runs = bw[11 * 32 + gl];
			if (__any_sync(FULL, shade_runs != 0))
			{
				Stage s0;
				Geo g0;
				int slen, nr;
				bool longcol;
				unsigned flags;
				qbuf_get(bw, gl, s0, g0, slen, nr, longcol, flags);
				shade_batch<false, false>(P, R, Cn, s0, g0, slen, nr, shade_runs, bw + RLERC_Q_STAGE * 32 + RLERC_RW * 64);
			}
			__syncwarp();
			if (gl == 0) mbar_arrive(bars + QB_S_FREE + (b % RLERC_Q_BUFS));
			if (last) break;
		}
		QPROF_END(3);
	}
	RLERC_EXIT_AFTER_PREPASS();
}

void launch_traverse_quad(const TraverseParams& p, cudaStream_t st)
{
	const int rays = launch_rays(p);
	if (rays <= 0) return;
	const int blocks = (rays + RLERC_Q_PLANES - 1) / RLERC_Q_PLANES;
	const size_t smem = (size_t)RLERC_Q_PLANES * q_words_per_plane(p.mask_words) * sizeof(uint32_t);
	static size_t configured_on[64] = { 0 };
	opt_in_smem(k_traverse_q, smem, configured_on);
	launch_overlapped(k_traverse_q, blocks, RLERC_Q_PLANES * 128, smem, st, p, rays);
}

void launch_traverse_filter(const TraverseParams& p, bool ids, cudaStream_t st)
{
	if (p.dda_mode == 99) launch_f<false, true>(p, st);          // tools/ray_profile.py
	else if (ids) launch_f<true, false>(p, st);
	else launch_f<false, false>(p, st);
}

} // namespace rlerc
