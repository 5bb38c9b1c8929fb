// k_traverse_f — the production traversal kernel: one WARP per ray plane, a column FILTER in front of the
// occlusion machinery.  Replaces cudaRender + Render::render_line (R/src/Cuda_Main.cu:150-181,
// R/src/Cuda_Render.h:96-737) and gives the same warped ray buffer bit for bit (arithmetic contract: DESIGN.md §3).
//
// Measured on B200 (tools/chain_probe.py): a frame's traversal time is the serial chain of its longest ray planes
// (the ones that see sky and walk all ~7000 cell crossings to z_far), not throughput; and on those ray planes
// > 90 % of the visited columns are no-ops: their first visible run already projects at or below the floating
// horizon y_clip_min, so the reference breaks out of the run loop without touching any state
// (Cuda_Render.h:542-543).  That test needs nothing but the 8-byte pointer-map entry (the first run rides in it)
// and it is monotone: the horizon only rises, so a column that is dead under today's horizon is dead under
// every later one.  Hence two loops instead of one pipeline over all columns:
//
//   FILTER   per 32 crossings: DDA (uniform serial recurrence) -> per lane column address, projected cell,
//            conservative top-clip test (Cuda_Render.h:467), pointer-map gather (consumed one step later, the
//            DDA of the next step hides its latency) -> first-run test -> the LIVE columns are compacted
//            (ballot + popc) into a small queue in shared memory, in crossing order.
//   CONSUME  per 32 LIVE columns: run-word loads (issued one round early), projection of up to RW runs,
//            then consume_batch (traverse_common.cuh): rising-horizon prefix-max fast path, owner-lane event
//            loop, cooperative long spans, deferred parallel shading — unchanged exact semantics, it re-tests
//            every column under the exact state.
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
#define RLERC_F_REC 136                     // words: 33 crossing records (float4), padded
#define RLERC_F_QUEUE (8 * RLERC_QCAP)      // words: 8 fields x QCAP, field-major

// The DDA state as two register quads, one per track, laid out like the crossing record a lane consumes:
// {distance, pos.x, pos.y, mip}.  The z-track keeps its distance NEGATED (the record marks the track that fired
// by the sign of its distance, and -(a + b) == (-a) + (-b) exactly), so a crossing is: compare, store the quad of
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

// EXPERIMENTAL DDA crossings as hand-scheduled PTX (build knob RLERC_DDA_ASM, default 0 = the C++ recurrence in
// ddaq_batch).  All three are bit-exact on B200 and all three are SLOWER than what nvcc makes of the C++ loop
// (tools/ab_libs.py, 1080p fly-through, four frames in flight: 0.546 ms per frame for the C++ loop):
//   1: predicates folded with the writer flag (only the writer lane's float state is ever read); x-track quad
//      stored by the writer, z-track quad stored over it when that track fires; 3 adds @w1, 3 adds @!w1
//      (9 instructions + overhead, two STS.128 per crossing)                                             0.572 ms
//   2: setp w1|w0, one store @w1, one @w0, 3 adds @w1, 3 adds @w0 (10 instructions, one active store)     0.557 ms
//   3: no shared memory: every lane runs the recurrence, lane j keeps the record of crossing j in registers
//      (guarded selp per field), the consumer takes record j-1 with one shuffle-up (11 instructions, no STS/LDS,
//      no __syncwarp); same instruction total per frame, more instruction-cache misses (ncu no_instruction
//      stall 0.21 -> 0.48 per issue), IPC 2.73 -> 2.51                                                    0.607 ms
// The packed add.rn.f32x2 (FADD2, sm_100+) was tried too: ptxas does not predicate it (FADD2 + two SEL), a loss.
// IEEE adds (add.rn.f32, no FMA, no FTZ), ordered compare: NaN -> x-track, like `<`.
#ifndef RLERC_DDA_ASM
#define RLERC_DDA_ASM 0
#endif
#ifndef RLERC_DDA_UNROLL
#define RLERC_DDA_UNROLL 4
#endif
__device__ __forceinline__ void ddaq_cross(DdaQ& Q, float mipf, uint32_t out_s, int writer)
{
#if RLERC_DDA_ASM == 2
	asm volatile("{\n"
		" .reg .pred pw, w1, w0;\n"
		" .reg .f32 pd1;\n"
		" setp.ne.s32 pw, %13, 0;\n"
		" neg.f32 pd1, %3;\n"
		" setp.lt.and.f32 w1|w0, pd1, %0, pw;\n"
		" @w1 st.shared.v4.f32 [%12], {%3, %4, %5, %14};\n"
		" @w0 st.shared.v4.f32 [%12], {%0, %1, %2, %14};\n"
		" @w1 add.rn.f32 %3, %3, %9;\n"
		" @w0 add.rn.f32 %0, %0, %6;\n"
		" @w1 add.rn.f32 %4, %4, %10;\n"
		" @w0 add.rn.f32 %1, %1, %7;\n"
		" @w1 add.rn.f32 %5, %5, %11;\n"
		" @w0 add.rn.f32 %2, %2, %8;\n"
		"}\n"
		: "+f"(Q.d0), "+f"(Q.x0), "+f"(Q.y0), "+f"(Q.nd1), "+f"(Q.x1), "+f"(Q.y1)
		: "f"(Q.gd0), "f"(Q.gx0), "f"(Q.gy0), "f"(Q.ngd1), "f"(Q.gx1), "f"(Q.gy1), "r"(out_s), "r"(writer), "f"(mipf)
		: "memory");
#else
	asm volatile("{\n"
		" .reg .pred pw, w1;\n"
		" .reg .f32 pd1;\n"
		" setp.ne.s32 pw, %13, 0;\n"
		" neg.f32 pd1, %3;\n"
		" setp.lt.and.f32 w1, pd1, %0, pw;\n"
		" @pw st.shared.v4.f32 [%12], {%0, %1, %2, %14};\n"
		" @w1 st.shared.v4.f32 [%12], {%3, %4, %5, %14};\n"
		" @w1 add.rn.f32 %3, %3, %9;\n"
		" @!w1 add.rn.f32 %0, %0, %6;\n"
		" @w1 add.rn.f32 %4, %4, %10;\n"
		" @!w1 add.rn.f32 %1, %1, %7;\n"
		" @w1 add.rn.f32 %5, %5, %11;\n"
		" @!w1 add.rn.f32 %2, %2, %8;\n"
		"}\n"
		: "+f"(Q.d0), "+f"(Q.x0), "+f"(Q.y0), "+f"(Q.nd1), "+f"(Q.x1), "+f"(Q.y1)
		: "f"(Q.gd0), "f"(Q.gx0), "f"(Q.gy0), "f"(Q.ngd1), "f"(Q.gx1), "f"(Q.gy1), "r"(out_s), "r"(writer), "f"(mipf)
		: "memory");
#endif
}

// RLERC_DDA_ASM 3: no shared memory at all.  Every lane runs the recurrence (the state stays warp-uniform) and lane j
// keeps the record of crossing j in its own registers: a guarded select per field, nothing to store, nothing to
// synchronise; the consumer of crossing j needs the records j-1 and j, i.e. its own and one shuffle-up.
//   setp t1 | setp pj = (lane == crossing) | 3 x @pj selp | 3 adds @t1 | 3 adds @!t1    -> 11 instructions, 0 stores
struct DdaCap {
	float cx, cy, cz; int cmip;      // record of crossing `lane` of the current batch {+-distance, pos.x, pos.y}, its mip level
	float kx, ky, kz;                // record before crossing 0: the last crossing of the previous batch (warp-uniform)
};
template <int K>
__device__ __forceinline__ void ddaq_cross_cap(DdaQ& Q, DdaCap& C, int rel)          // lane keeps the record iff rel == K
{
	asm volatile("{\n"
		" .reg .pred t1, pj;\n"
		" .reg .f32 pd1;\n"
		" neg.f32 pd1, %3;\n"
		" setp.lt.f32 t1, pd1, %0;\n"
		" setp.eq.s32 pj, %15, %16;\n"
		" @pj selp.f32 %6, %3, %0, t1;\n"
		" @pj selp.f32 %7, %4, %1, t1;\n"
		" @pj selp.f32 %8, %5, %2, t1;\n"
		" @t1 add.rn.f32 %3, %3, %12;\n"
		" @!t1 add.rn.f32 %0, %0, %9;\n"
		" @t1 add.rn.f32 %4, %4, %13;\n"
		" @!t1 add.rn.f32 %1, %1, %10;\n"
		" @t1 add.rn.f32 %5, %5, %14;\n"
		" @!t1 add.rn.f32 %2, %2, %11;\n"
		"}\n"
		: "+f"(Q.d0), "+f"(Q.x0), "+f"(Q.y0), "+f"(Q.nd1), "+f"(Q.x1), "+f"(Q.y1), "+f"(C.cx), "+f"(C.cy), "+f"(C.cz)
		: "f"(Q.gd0), "f"(Q.gx0), "f"(Q.gy0), "f"(Q.ngd1), "f"(Q.gx1), "f"(Q.gy1), "r"(rel), "n"(K));
}

// Same contract as ddaq_batch below, records captured in registers (C) instead of shared memory.
__device__ __forceinline__ int ddaq_batch_cap(DdaQ& Q, DdaCap& C, int prev_n, int last_map, int zfar_i, int gl)
{
	int nvalid = 32;
	if (prev_n > 0)
	{
		C.kx = __shfl_sync(0xffffffffu, C.cx, prev_n - 1);
		C.ky = __shfl_sync(0xffffffffu, C.cy, prev_n - 1);
		C.kz = __shfl_sync(0xffffffffu, C.cz, prev_n - 1);
	}
	for (int s = 0; s < 32;)
	{
		while (Q.zi > Q.mapswitch) ddaq_lod_switch(Q, last_map);
		const int sh = 31 - __clz(Q.dzi);
		const int lod_free = ((Q.mapswitch - Q.zi) >> sh) + 1;        // crossings before z > mapswitch
		const int far_free = (zfar_i - Q.zi) >> sh;                   // crossings with z + dz <= z_far (<= 0: none)
		if (far_free <= 0) { nvalid = s; break; }
		int n = 32 - s;
		n = n < lod_free ? n : lod_free;
		n = n < far_free ? n : far_free;
		const int rel = gl - s;                                       // this lane keeps crossing j == rel of the segment
		if (rel >= 0 && rel < n) C.cmip = Q.mip;
		int r = rel, j = 0;
		for (; j + 4 <= n; j += 4, r -= 4)
		{
			ddaq_cross_cap<0>(Q, C, r); ddaq_cross_cap<1>(Q, C, r); ddaq_cross_cap<2>(Q, C, r); ddaq_cross_cap<3>(Q, C, r);
		}
		#pragma unroll 1
		for (; j < n; j++, r--) ddaq_cross_cap<0>(Q, C, r);
		Q.zi += n << sh;
		s += n;
	}
	return nvalid;
}

// Up to 32 crossings, all lanes in lockstep (one lane writes: 32 lanes storing the same 16 bytes cost four
// shared-memory passes per crossing, which made the DDA store-bound); rec[s+1] = record of crossing s; rec[0] = the last record of the
// previous batch (rec[prev_n], or zeros before the first).  LOD / z_far budgets by shifts (dz is a power of two).
// Returns the number of crossings made (< 32 only when z_far was reached, Cuda_Render.h:366-367).
__device__ __forceinline__ int ddaq_batch(DdaQ& Q, float4* rec, int prev_n, int last_map, int zfar_i, bool writer_b)
{
	const int writer = writer_b ? 1 : 0;
	int nvalid = 32;
	{
		const float4 carry = rec[prev_n];
		__syncwarp();
		if (writer) rec[0] = make_float4(carry.x, carry.y, carry.z, 0.0f);
	}
	for (int s = 0; s < 32;)
	{
		while (Q.zi > Q.mapswitch) ddaq_lod_switch(Q, last_map);
		const int sh = 31 - __clz(Q.dzi);
		const int lod_free = ((Q.mapswitch - Q.zi) >> sh) + 1;        // crossings before z > mapswitch
		const int far_free = (zfar_i - Q.zi) >> sh;                   // crossings with z + dz <= z_far (<= 0: none)
		if (far_free <= 0) { nvalid = s; break; }
		int n = 32 - s;
		n = n < lod_free ? n : lod_free;
		n = n < far_free ? n : far_free;
		const float mipf = __int_as_float(Q.mip);
		float4* out = rec + s + 1;
#if RLERC_DDA_ASM
		const uint32_t out_s = (uint32_t)__cvta_generic_to_shared(out);
		constexpr int DDA_UNROLL = RLERC_DDA_UNROLL;
		#pragma unroll DDA_UNROLL
		for (int j = 0; j < n; j++) ddaq_cross(Q, mipf, out_s + 16u * (uint32_t)j, writer);
#else
		constexpr int DDA_UNROLL = RLERC_DDA_UNROLL;
		#pragma unroll DDA_UNROLL
		for (int j = 0; j < n; j++)
		{
			const bool t1 = -Q.nd1 < Q.d0;                            // Cuda_Render.h:398-414
			const float4 cur = make_float4(t1 ? Q.nd1 : Q.d0, t1 ? Q.x1 : Q.x0, t1 ? Q.y1 : Q.y0, mipf);
			if (writer) out[j] = cur;
			if (t1) { Q.nd1 += Q.ngd1; Q.x1 += Q.gx1; Q.y1 += Q.gy1; }
			else    { Q.d0 += Q.gd0; Q.x0 += Q.gx0; Q.y0 += Q.gy0; }
		}
#endif
		Q.zi += n << sh;
		s += n;
	}
	return nvalid;
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
	if (ray_i >= rays || x >= P.ray_end) return;

	// shared per warp: crossing records | live-column queue | DrawJob | RW x 32 projected runs (int2) |
	//                  RW x 32 deferred short spans | occlusion bits
	const int per_warp = (RLERC_F_REC + RLERC_F_QUEUE + 16 + RLERC_PS_WORDS + P.mask_words + 3) & ~3;
	uint32_t* wbase = smem + (size_t)wid * per_warp;
	float4* rec = reinterpret_cast<float4*>(wbase);
	uint32_t* queue = wbase + RLERC_F_REC;                           // [8][QCAP]
	DrawJob* job = reinterpret_cast<DrawJob*>(wbase + RLERC_F_REC + RLERC_F_QUEUE);
	int2* proj = reinterpret_cast<int2*>(wbase + RLERC_F_REC + RLERC_F_QUEUE + 16);
	uint32_t* shade = wbase + RLERC_F_REC + RLERC_F_QUEUE + 16 + RLERC_RW * 64;
	uint32_t* ymask = wbase + RLERC_F_REC + RLERC_F_QUEUE + 16 + RLERC_PS_WORDS;

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
	int fixx, fixz;
	DdaQ Q;
	{
		Dda dd;
		dda_init(P, ray_x, ray_z, dd);
		fixx = dd.fixx; fixz = dd.fixz;
		Q.d0 = dd.d0; Q.x0 = dd.i0x; Q.y0 = dd.i0y; Q.nd1 = -dd.d1; Q.x1 = dd.i1x; Q.y1 = dd.i1y;
		Q.gd0 = dd.gd0; Q.gx0 = dd.g0x; Q.gy0 = dd.g0y; Q.ngd1 = -dd.gd1; Q.gx1 = dd.g1x; Q.gy1 = dd.g1y;
	}
	Q.mip = 0;
	Q.zi = 0; Q.dzi = 1;                                         // z and dz (Cuda_Render.h:181,325), integer valued
	Q.mapswitch = P.mapswitch0;
	DdaCap cap;
	cap.cx = cap.cy = cap.cz = 0.0f; cap.cmip = 0; cap.kx = cap.ky = cap.kz = 0.0f;   // (register-capture DDA)
	if (gl == 0) rec[0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // no crossing yet: distance 0, x-track (Cuda_Render.h:302-305)
	int prev_n = 0;
	__syncwarp();
	const float pz_add = sin_x;                                  // pos3d_z_add (Cuda_Render.h:313)
	const float py_add = (vertical ? cos_x : 0.0f) * rx2mr;      // pos3d_y_add (Cuda_Render.h:314-315)
	const int zfar_i = P.z_far;
	const int last_map = P.nummaps - 1;
	// The y_map_switch half of the LOD loop condition (Cuda_Render.h:343) can only be true on
	// the first crossing (it halves until <= 512 and never grows), where z = 0 < mapswitch.
	for (float yms = mountain; yms > 512.0f; yms = yms * 0.5f) ddaq_lod_switch(Q, last_map);

	Counters Cn;
	memset(&Cn, 0, sizeof(Cn));

	RayCtx R;
	R.row = row; R.ymask = ymask; R.ids = (IDS && !PROF) ? P.ids + (size_t)x * res_y * 2 : nullptr;
	R.res_y2 = res_y2; R.pz_add = pz_add; R.py_add = py_add; R.mountain = mountain; R.gl = gl;
	R.stat = PROF ? stat : nullptr;
	R.hc_on = (P.flags & 2) ? 1 : 0;
	R.hc = f2i((4095.0f - mountain) + P.viewpos[1]);              // int height_color = 4095-mountain+viewpos.y (Cuda_Render.h:675)

	// filter: the batch whose pointer-map gather is in flight
	Geo fg;
	fg.pz = fg.py = fg.czz = fg.cyy = 0; fg.cmip = 0; fg.cidx = 0;
	unsigned fe0 = 0, fe1 = 0;
	bool fhave = false;
	int fn = 0;                        // crossings in that batch (0: none in flight)
	bool dda_done = false;
	int qhead = 0, qcount = 0;         // live-column queue (uniform)
	// consume: the batch whose run words are in flight
	Stage s0;
	Geo g0 = fg;
	s0.nvalid = 0; s0.have = false; s0.e0 = s0.e1 = 0;
	#pragma unroll
	for (int k = 0; k < 4; k++) s0.rw[k] = 0;

	while (true)
	{
		if (Hs.ycmin >= Hs.ycmax) break;                         // Cuda_Render.h:370

		// ---- FILTER: until a full batch of live columns is queued (or the ray plane has reached z_far) -------------
		while (qcount < 32 && (!dda_done || fn > 0))
		{
			// F1. DDA for the next 32 crossings; the gather of the batch in flight lands meanwhile
			int nvalid = 0;
			RLERC_TICK(0);
			if (PROF) prof[7] += 1;
			if (!dda_done)
			{
#if RLERC_DDA_ASM == 3
				nvalid = ddaq_batch_cap(Q, cap, prev_n, last_map, zfar_i, gl);
#else
				nvalid = ddaq_batch(Q, rec, prev_n, last_map, zfar_i, gl == 0);
#endif
				prev_n = nvalid;
				if (nvalid < G) dda_done = true;
				if (IDS && gl == 0) Cn.c_steps += nvalid;
			}
#if RLERC_DDA_ASM != 3
			__syncwarp();
#endif
			RLERC_TICK(1);
			// F2. first-run test of the batch in flight, live columns -> queue
			if (fn > 0)
			{
				const int ycmin = Hs.ycmin;
				bool live = false;
				if (gl < fn && fhave)
				{
					const int slen = (int)(fe1 & 0xffffu);
					const unsigned first = fe1 >> 16;
					const int solid = (int)(first >> 10), skip = (int)(first & 1023u);
					if (IDS) live = true;                                // the byte model counts no-op columns too
					else if (slen == 0) live = false;                    // empty column: the run loop does not execute
					else if (solid == 0) live = true;                    // pure skip run: undecided, let the machinery look
					else
					{
						const float ft = (float)(skip << fg.cmip);         // Cuda_Render.h:529-543 for run 0
						float zz1 = fg.pz, yy1 = fg.py;
						if (mountain + ft >= 0) { zz1 += fg.czz; yy1 += fg.cyy; }
						const float z1 = zz1 + pz_add * ft;
						if (z1 <= 0) live = true;                          // `continue`: a later run may be the first visible one
						else
						{
							const float y1 = yy1 + py_add * ft;
							live = f2i(res_y2 + y1 / z1) > ycmin;          // else: break, now and under every later horizon
						}
					}
				}
				const unsigned lb = __ballot_sync(FULL, live);
				if (live)
				{
					uint32_t* q = queue + ((qhead + qcount + __popc(lb & lt_mask)) & (RLERC_QCAP - 1));
					q[0 * RLERC_QCAP] = __float_as_uint(fg.pz); q[1 * RLERC_QCAP] = __float_as_uint(fg.py);
					q[2 * RLERC_QCAP] = __float_as_uint(fg.czz); q[3 * RLERC_QCAP] = __float_as_uint(fg.cyy);
					q[4 * RLERC_QCAP] = (uint32_t)fg.cmip; q[5 * RLERC_QCAP] = (uint32_t)fg.cidx;
					q[6 * RLERC_QCAP] = fe0; q[7 * RLERC_QCAP] = fe1;
				}
				qcount += __popc(lb);
			}
			RLERC_TICK(2);
			// F3. geometry of the new crossings, conservative top clip, pointer-map gather (into the registers F2 freed)
			fn = nvalid;
			fhave = false;
#if RLERC_DDA_ASM == 3
			float4 ra, rb;                                             // state before / after crossing gl
			if (nvalid > 0)
			{
				ra.x = __shfl_up_sync(FULL, cap.cx, 1); ra.y = __shfl_up_sync(FULL, cap.cy, 1); ra.z = __shfl_up_sync(FULL, cap.cz, 1);
				if (gl == 0) { ra.x = cap.kx; ra.y = cap.ky; ra.z = cap.kz; }
				rb.x = cap.cx; rb.w = __int_as_float(cap.cmip);
			}
#endif
			if (gl < nvalid)
			{
#if RLERC_DDA_ASM != 3
				const float4 ra = rec[gl], rb = rec[gl + 1];           // state before / after crossing gl
#endif
				const float db = fabsf(ra.x), dn = fabsf(rb.x);
				const int ib = __float_as_int(ra.x) < 0 ? 1 : 0;        // index_before: sign bit of the record
				fg.cmip = __float_as_int(rb.w);
				const int fix_x = (1 - ib) * fixx, fix_z = ib * fixz;    // Cuda_Render.h:418-419
				const float ddelta = dn - db;
				const float vsx = ray_x * db, vsz = ray_z * db;
				const int voxel_x = f2i(vpx + ra.y) + fix_x;             // Cuda_Render.h:429-430
				const int voxel_z = f2i(vpz + ra.z) + fix_z;
				const int gx = P.level[fg.cmip].sx, gz = P.level[fg.cmip].sz;
				// CLIPREGION (Cuda_Render.h:432-437): finite scene, columns outside the grid are skipped
				const bool outside = (P.flags & 1) && (voxel_x < 0 || voxel_z < 0 || (voxel_x >> fg.cmip) > gx - 1 || (voxel_z >> fg.cmip) > gz - 1);
				const int vx = (voxel_x >> fg.cmip) & (gx - 1);          // Cuda_Render.h:441-442
				const int vz = (voxel_z >> fg.cmip) & (gz - 1);
				fg.cidx = vx + vz * gx;
				const float corx = ray_x * ddelta, corz = ray_z * ddelta;
				fg.pz = cos_x * vsz + sin_x * mountain;                  // Cuda_Render.h:459-464
				fg.py = vertical ? (cos_x * mountain - sin_x * vsz) : vsx;
				fg.py *= rx2mr;
				fg.czz = cos_x * corz;                                   // Cuda_Render.h:483-486
				fg.cyy = vertical ? (-sin_x * corz) : corx;
				fg.cyy *= rx2mr;
				// The horizon only rises.  For pz > 0 a column culled now stays culled; for pz <= 0 (or NaN)
				// the test can flip, so keep those.
				fhave = !outside && (!(fg.pz * res_y2 + fg.py <= fg.pz * (float)Hs.ycmin) || !(fg.pz > 0));   // Cuda_Render.h:467
				if (fhave)
				{
					const uint2 ent = __ldg(P.level[fg.cmip].map + fg.cidx);         // Cuda_Render.h:474-478
					fe0 = ent.x; fe1 = ent.y;
				}
			}
			__syncwarp();
			RLERC_TICK(3);
		}
		RLERC_TICK(0);

		// ---- C1. take the next batch of live columns off the queue, request their run words ---------------------
		Stage s1;
		Geo g1;
		{
			const int n1 = qcount < 32 ? qcount : 32;
			s1.nvalid = n1; s1.have = gl < n1;
			s1.e0 = s1.e1 = 0;
			#pragma unroll
			for (int k = 0; k < 4; k++) s1.rw[k] = 0;
			g1.pz = g1.py = g1.czz = g1.cyy = 0; g1.cmip = 0; g1.cidx = 0;
			if (s1.have)
			{
				const uint32_t* q = queue + ((qhead + gl) & (RLERC_QCAP - 1));
				g1.pz = __uint_as_float(q[0 * RLERC_QCAP]); g1.py = __uint_as_float(q[1 * RLERC_QCAP]);
				g1.czz = __uint_as_float(q[2 * RLERC_QCAP]); g1.cyy = __uint_as_float(q[3 * RLERC_QCAP]);
				g1.cmip = (int)q[4 * RLERC_QCAP]; g1.cidx = (int)q[5 * RLERC_QCAP];
				s1.e0 = q[6 * RLERC_QCAP]; s1.e1 = q[7 * RLERC_QCAP];
				const int sl = (int)(s1.e1 & 0xffffu);
				// element i0 of the slab stream is run 0; runs 0..7 are fetched as aligned 32-bit words
				const unsigned i0 = 2u + s1.e0;
				const uint32_t* w32 = reinterpret_cast<const uint32_t*>(P.level[g1.cmip].slabs);
				const uint32_t* p = w32 + ((i0 + (i0 & 1u)) >> 1);
				const int odd = (int)(i0 & 1u);                            // odd: words hold runs (1,2) (3,4) (5,6) (7,8)
				s1.rw[0] = (sl > 1) ? __ldg(p) : 0u;
				s1.rw[1] = (sl > 2 + odd) ? __ldg(p + 1) : 0u;
				s1.rw[2] = (sl > 4 + odd) ? __ldg(p + 2) : 0u;
				s1.rw[3] = (sl > 6 + odd) ? __ldg(p + 3) : 0u;
			}
			qhead = (qhead + n1) & (RLERC_QCAP - 1);
			qcount -= n1;
		}
		__syncwarp();
		RLERC_TICK(4);

		// ---- C2. project the runs of batch s0 (their words were requested one round ago) -------------------------
		if (s0.nvalid > 0)
		{
			const int ycmin = Hs.ycmin;
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

			RLERC_TICK(5);
			if (PROF) prof[7] += 1ll << 32;
			// ---- B / B0 / S. consume batch s0 (traverse_common.cuh) ----------------------------------------------
			const bool finished = consume_batch<IDS, PROF>(P, R, Hs, Cn, s0, g0, slen, nr, longcol, flags, proj, shade, job);
			RLERC_TICK(6);
			if (finished) break;
		}
		else if (s1.nvalid == 0 && dda_done && fn == 0 && qcount == 0) break;   // drained (z > z_far, Cuda_Render.h:367)

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
}

template <bool IDS, bool PROF>
static void launch_f(const TraverseParams& p, cudaStream_t st)
{
	const int wpb = RLERC_F_WPB;
	const int rays = (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
	if (rays <= 0) return;
	const int blocks = (rays + wpb - 1) / wpb;
	const size_t smem = (size_t)wpb * ((RLERC_F_REC + RLERC_F_QUEUE + 16 + RLERC_PS_WORDS + p.mask_words + 3) & ~3) * sizeof(uint32_t);
	// dynamic shared memory above 48 KB is an opt-in per kernel AND per device
	static size_t configured_on[64] = { 0 };
	int dev = 0;
	cudaGetDevice(&dev);
	size_t& configured = configured_on[dev & 63];
	if (smem > configured)
	{
		cudaFuncSetAttribute(k_traverse_f<IDS, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = smem;
	}
	k_traverse_f<IDS, PROF><<<blocks, wpb * 32, smem, st>>>(p, rays);
}


// ================================================================================================================
// k_traverse_p (variant 68; picked automatically for launches of at most 12 ray planes per SM, capi.cu pick_lanes):
// the same algorithm with the two loops of k_traverse_f on TWO warps per
// ray plane.  Warp F runs the FILTER (DDA, first-run test, pointer-map gather) and pushes the live columns into a ring
// in shared memory; warp C pops batches of 32, loads and projects their runs and runs consume_batch.  A ray plane's
// chain becomes max(filter, consume) instead of their sum.  The filter only ever drops provable no-ops under a horizon
// that can only rise, so a stale y_clip_min (published by C after every batch) lets more columns through but cannot
// change the result: the picture does not depend on timing.
// Protocol (volatile words in shared memory, one writer each): tail (F), head (C), ycmin (C), done (F), closed (C).
// F waits while the ring has no room for 32 more entries, C waits until 32 entries are there or F is done; both
// waits are bounded (a broken protocol traps after >= 1 s of polling: a launch error, neither a hang nor a wrong picture).  Waiting is __nanosleep(64)
// polling by the whole warp; on a full frame the polling loops are up to 30 % of the kernel's instructions (ncu source
// view).  Tried instead, all bit-exact, all slower (1080p frame 0, full frame / uncontended chain, ms; shipped: 0.86 /
// 0.36): back-off 128 ns .. 2 us 0.85 / 0.37; back-off 512 ns .. 8 us 0.94 / 0.46; mbarrier.try_wait as "sleep until
// the partner signals" (one arrive per tail / head update; the phase bookkeeping falls behind when nobody waits, so
// waits run into their time limit) 1.09 / 0.41; lane 0 polling alone with one vector load of the control words, the
// other lanes parked at the broadcast shuffle 1.16 / 0.48.
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
	const int per_plane = (RLERC_F_REC + RLERC_P_RING + 8 + 16 + RLERC_PS_WORDS + P.mask_words + 3) & ~3;
	uint32_t* wbase = smem + (size_t)pl * per_plane;
	float4* rec = reinterpret_cast<float4*>(wbase);
	uint32_t* ring = wbase + RLERC_F_REC;                            // [8][RCAP]
	volatile int* ctl = reinterpret_cast<volatile int*>(wbase + RLERC_F_REC + RLERC_P_RING);   // tail, head, ycmin, done, closed
	DrawJob* job = reinterpret_cast<DrawJob*>(wbase + RLERC_F_REC + RLERC_P_RING + 8);
	int2* proj = reinterpret_cast<int2*>(wbase + RLERC_F_REC + RLERC_P_RING + 8 + 16);
	uint32_t* shade = wbase + RLERC_F_REC + RLERC_P_RING + 8 + 16 + RLERC_RW * 64;
	uint32_t* ymask = wbase + RLERC_F_REC + RLERC_P_RING + 8 + 16 + RLERC_PS_WORDS;

	const int res_y = P.res_y;
	const float res_y2 = (float)(res_y / 2);
	uint32_t* row = P.warp + (size_t)x * res_y;
	RayInit ri;
	ray_init(P, x, ri);
	const float ray_x = ri.ray_x, ray_z = ri.ray_z, rx2mr = ri.rx2mr;
	const bool vertical = ri.vertical;
	const float sin_x = P.sin_x, cos_x = P.cos_x;
	const float vpx = P.viewpos[0], mountain = P.viewpos[1], vpz = P.viewpos[2];
	const float pz_add = sin_x;
	const float py_add = (vertical ? cos_x : 0.0f) * rx2mr;
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
		if (!active) return;
		// ================= warp F: DDA -> geometry + gather -> first-run test -> ring ==========================
		int fixx, fixz;
		DdaQ Q;
		{
			Dda dd;
			dda_init(P, ray_x, ray_z, dd);
			fixx = dd.fixx; fixz = dd.fixz;
			Q.d0 = dd.d0; Q.x0 = dd.i0x; Q.y0 = dd.i0y; Q.nd1 = -dd.d1; Q.x1 = dd.i1x; Q.y1 = dd.i1y;
			Q.gd0 = dd.gd0; Q.gx0 = dd.g0x; Q.gy0 = dd.g0y; Q.ngd1 = -dd.gd1; Q.gx1 = dd.g1x; Q.gy1 = dd.g1y;
		}
		Q.mip = 0; Q.zi = 0; Q.dzi = 1; Q.mapswitch = P.mapswitch0;
		if (gl == 0) rec[0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		int prev_n = 0;
		__syncwarp();
		const int zfar_i = P.z_far;
		const int last_map = P.nummaps - 1;
		for (float yms = mountain; yms > 512.0f; yms = yms * 0.5f) ddaq_lod_switch(Q, last_map);
		Geo fg;
		fg.pz = fg.py = fg.czz = fg.cyy = 0; fg.cmip = 0; fg.cidx = 0;
		unsigned fe0 = 0, fe1 = 0;
		bool fhave = false;
		int fn = 0;
		bool dda_done = false;
		int tail = 0;
		bool closed = false;
		while (!dda_done || fn > 0)
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
			int nvalid = 0;
			if (!dda_done)
			{
				nvalid = ddaq_batch(Q, rec, prev_n, last_map, zfar_i, gl == 0);
				prev_n = nvalid;
				if (nvalid < G) dda_done = true;
			}
			__syncwarp();
			if (fn > 0)
			{
				bool live = false;
				if (gl < fn && fhave)
				{
					const int slen = (int)(fe1 & 0xffffu);
					const unsigned first = fe1 >> 16;
					const int solid = (int)(first >> 10), skip = (int)(first & 1023u);
					if (slen == 0) live = false;
					else if (solid == 0) live = true;
					else
					{
						const float ft = (float)(skip << fg.cmip);
						float zz1 = fg.pz, yy1 = fg.py;
						if (mountain + ft >= 0) { zz1 += fg.czz; yy1 += fg.cyy; }
						const float z1 = zz1 + pz_add * ft;
						if (z1 <= 0) live = true;
						else
						{
							const float y1 = yy1 + py_add * ft;
							live = f2i(res_y2 + y1 / z1) > ycmin;
						}
					}
				}
				const unsigned lb = __ballot_sync(FULL, live);
				if (live)
				{
					uint32_t* q = ring + ((tail + __popc(lb & lt_mask)) & (RLERC_P_RCAP - 1));
					q[0 * RLERC_P_RCAP] = __float_as_uint(fg.pz); q[1 * RLERC_P_RCAP] = __float_as_uint(fg.py);
					q[2 * RLERC_P_RCAP] = __float_as_uint(fg.czz); q[3 * RLERC_P_RCAP] = __float_as_uint(fg.cyy);
					q[4 * RLERC_P_RCAP] = (uint32_t)fg.cmip; q[5 * RLERC_P_RCAP] = (uint32_t)fg.cidx;
					q[6 * RLERC_P_RCAP] = fe0; q[7 * RLERC_P_RCAP] = fe1;
				}
				tail += __popc(lb);
				__threadfence_block();
				__syncwarp();
				if (gl == 0 && lb) ctl[0] = tail;
			}
			fn = nvalid;
			fhave = false;
			if (gl < nvalid)
			{
				const float4 ra = rec[gl], rb = rec[gl + 1];
				const float db = fabsf(ra.x), dn = fabsf(rb.x);
				const int ib = __float_as_int(ra.x) < 0 ? 1 : 0;
				fg.cmip = __float_as_int(rb.w);
				const int fix_x = (1 - ib) * fixx, fix_z = ib * fixz;
				const float ddelta = dn - db;
				const float vsx = ray_x * db, vsz = ray_z * db;
				const int voxel_x = f2i(vpx + ra.y) + fix_x;
				const int voxel_z = f2i(vpz + ra.z) + fix_z;
				const int gx = P.level[fg.cmip].sx, gz = P.level[fg.cmip].sz;
				const bool outside = (P.flags & 1) && (voxel_x < 0 || voxel_z < 0 || (voxel_x >> fg.cmip) > gx - 1 || (voxel_z >> fg.cmip) > gz - 1);
				const int vx = (voxel_x >> fg.cmip) & (gx - 1);
				const int vz = (voxel_z >> fg.cmip) & (gz - 1);
				fg.cidx = vx + vz * gx;
				const float corx = ray_x * ddelta, corz = ray_z * ddelta;
				fg.pz = cos_x * vsz + sin_x * mountain;
				fg.py = vertical ? (cos_x * mountain - sin_x * vsz) : vsx;
				fg.py *= rx2mr;
				fg.czz = cos_x * corz;
				fg.cyy = vertical ? (-sin_x * corz) : corx;
				fg.cyy *= rx2mr;
				fhave = !outside && (!(fg.pz * res_y2 + fg.py <= fg.pz * (float)ycmin) || !(fg.pz > 0));
				if (fhave)
				{
					const uint2 ent = __ldg(P.level[fg.cmip].map + fg.cidx);
					fe0 = ent.x; fe1 = ent.y;
				}
			}
			__syncwarp();
		}
		__threadfence_block();
		__syncwarp();
		if (gl == 0 && !closed) { ctl[0] = tail; __threadfence_block(); ctl[3] = 1; }
		return;
	}

	// ===================== warp C: ring -> run words -> projection -> consume_batch ============================
#if RLERC_P_REG_F
	asm volatile("setmaxnreg.inc.sync.aligned.u32 " RLERC_STR(RLERC_P_REG_C) ";");
#endif
	if (!active) return;
	HorizonState Hs;
	Hs.ycmin = ri.ycmin; Hs.ycmax = ri.ycmax; Hs.hiw = 0;
	Counters Cn;
	memset(&Cn, 0, sizeof(Cn));
	RayCtx R;
	R.row = row; R.ymask = ymask; R.ids = nullptr;
	R.res_y2 = res_y2; R.pz_add = pz_add; R.py_add = py_add; R.mountain = mountain; R.gl = gl;
	R.stat = nullptr;
	R.hc_on = (P.flags & 2) ? 1 : 0;
	R.hc = f2i((4095.0f - mountain) + P.viewpos[1]);
	Geo g0;
	g0.pz = g0.py = g0.czz = g0.cyy = 0; g0.cmip = 0; g0.cidx = 0;
	Stage s0;
	s0.nvalid = 0; s0.have = false; s0.e0 = s0.e1 = 0;
	#pragma unroll
	for (int k = 0; k < 4; k++) s0.rw[k] = 0;
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
		{
			const int n1 = avail < 32 ? avail : 32;
			s1.nvalid = n1; s1.have = gl < n1;
			s1.e0 = s1.e1 = 0;
			#pragma unroll
			for (int k = 0; k < 4; k++) s1.rw[k] = 0;
			g1.pz = g1.py = g1.czz = g1.cyy = 0; g1.cmip = 0; g1.cidx = 0;
			if (s1.have)
			{
				const uint32_t* q = ring + ((head + gl) & (RLERC_P_RCAP - 1));
				g1.pz = __uint_as_float(q[0 * RLERC_P_RCAP]); g1.py = __uint_as_float(q[1 * RLERC_P_RCAP]);
				g1.czz = __uint_as_float(q[2 * RLERC_P_RCAP]); g1.cyy = __uint_as_float(q[3 * RLERC_P_RCAP]);
				g1.cmip = (int)q[4 * RLERC_P_RCAP]; g1.cidx = (int)q[5 * RLERC_P_RCAP];
				s1.e0 = q[6 * RLERC_P_RCAP]; s1.e1 = q[7 * RLERC_P_RCAP];
				const int sl = (int)(s1.e1 & 0xffffu);
				const unsigned i0 = 2u + s1.e0;
				const uint32_t* w32 = reinterpret_cast<const uint32_t*>(P.level[g1.cmip].slabs);
				const uint32_t* p = w32 + ((i0 + (i0 & 1u)) >> 1);
				const int odd = (int)(i0 & 1u);
				s1.rw[0] = (sl > 1) ? __ldg(p) : 0u;
				s1.rw[1] = (sl > 2 + odd) ? __ldg(p + 1) : 0u;
				s1.rw[2] = (sl > 4 + odd) ? __ldg(p + 2) : 0u;
				s1.rw[3] = (sl > 6 + odd) ? __ldg(p + 3) : 0u;
			}
			head += n1;
			__syncwarp();
			if (gl == 0 && n1) ctl[1] = head;                          // the slots may be overwritten from now on
		}
		// C2. project the runs of batch s0
		if (s0.nvalid > 0)
		{
			const int ycmin = Hs.ycmin;
			int slen = 0, nr = 0;
			bool longcol = false;
			unsigned flags = 0;
			if (s0.have)
			{
				{
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
					const int top = (blen + skip) << g0.cmip;
					const int bot = top + (solid << g0.cmip);
					blen += skip + solid;
					if (solid == 0) continue;
					const float ft = (float)top, fb = (float)bot;
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
					if (sy2 <= ycmin) { nr = r + 1; longcol = false; break; }
				}
			}
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
}

void launch_traverse_pair(const TraverseParams& p, cudaStream_t st)
{
	const int rays = (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
	if (rays <= 0) return;
	const int blocks = (rays + RLERC_P_PLANES - 1) / RLERC_P_PLANES;
	const size_t smem = (size_t)RLERC_P_PLANES * ((RLERC_F_REC + RLERC_P_RING + 8 + 16 + RLERC_PS_WORDS + p.mask_words + 3) & ~3) * sizeof(uint32_t);
	static size_t configured_on[64] = { 0 };
	int dev = 0;
	cudaGetDevice(&dev);
	size_t& configured = configured_on[dev & 63];
	if (smem > configured)
	{
		cudaFuncSetAttribute(k_traverse_p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = smem;
	}
	k_traverse_p<<<blocks, RLERC_P_PLANES * 64, smem, st>>>(p, rays);
}

void launch_traverse_filter(const TraverseParams& p, bool ids, cudaStream_t st)
{
	if (p.dda_mode == 99) launch_f<false, true>(p, st);          // tools/ray_profile.py
	else if (ids) launch_f<true, false>(p, st);
	else launch_f<false, false>(p, st);
}

} // namespace rlerc
