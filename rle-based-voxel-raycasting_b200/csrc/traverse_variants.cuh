// DDA helpers of the VARIANT kernels only (traverse_warp.cu, traverse_chunk.cu; built with `make VARIANTS=1`): the
// warp-uniform serial batch and the closed-form / merge-path DDA.  The production kernels (traverse_filter.cu) run
// the DDA lane-parallel from saved states instead.
#pragma once
#include "traverse_common.cuh"
#include "dda_closed.cuh"

namespace rlerc {

#define RLERC_DDA_WORDS (66 * 4 + 72)   // shared words per warp for the DDA hand-over (merge path: 66 float4 + 72 float)

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


// Out-of-line serial batch (rare paths of the closed-form and merge-path builds)
static __device__ __noinline__ int dda_serial_batch(DdaState& S, float4* rec, int last_map, int zfar_i)
{
	return dda_serial_batch_inl(S, rec, last_map, zfar_i);
}

// ---- closed-form DDA (dda_closed.cuh): lane-parallel, no serial recurrence -------------------------------------
// Lanes 0..5 OWN one DDA variable each (0-2: dds_dist0, isect0.x, isect0.y of the x-track; 3-5: the z-track) as a
// bit pattern plus its linear regime.  Per batch every owner publishes a plan of its variable's next >= 33 steps
// (8 words), every lane evaluates firing distance `lane` of both tracks into a table, finds its crossing with a
// merge-path search (dda_merge_search) and evaluates the fired track's position from the plans.  What the warp
// carries between batches besides the owners' variables is uniform: z, dz, mapswitch, mip and the last record.
struct DdaUni { int mip, zi, dzi, mapswitch; float csd, cpx, cpy; };   // csd: signed distance of the last crossing

// Serial fallback of the closed-form build: first crossings of a ray plane (a variable doubles more than once per
// batch), NaN rays.  Rebuilds the uniform float state from the owner lanes, runs the serial batch, hands it back.
static __device__ __noinline__ int dda_closed_fallback(DdaVar& var, DdaUni& U, float4* rec, int last_map, int zfar_i, int gl)
{
	const unsigned FULL = 0xffffffffu;
	DdaState T;
	T.d0 = __int_as_float(__shfl_sync(FULL, var.b, 0)); T.i0x = __int_as_float(__shfl_sync(FULL, var.b, 1)); T.i0y = __int_as_float(__shfl_sync(FULL, var.b, 2));
	T.d1 = __int_as_float(__shfl_sync(FULL, var.b, 3)); T.i1x = __int_as_float(__shfl_sync(FULL, var.b, 4)); T.i1y = __int_as_float(__shfl_sync(FULL, var.b, 5));
	T.gd0 = __int_as_float(__shfl_sync(FULL, var.gb, 0)); T.g0x = __int_as_float(__shfl_sync(FULL, var.gb, 1)); T.g0y = __int_as_float(__shfl_sync(FULL, var.gb, 2));
	T.gd1 = __int_as_float(__shfl_sync(FULL, var.gb, 3)); T.g1x = __int_as_float(__shfl_sync(FULL, var.gb, 4)); T.g1y = __int_as_float(__shfl_sync(FULL, var.gb, 5));
	T.dist_now = fabsf(U.csd); T.index = __float_as_int(U.csd) < 0 ? 1 : 0; T.posx = U.cpx; T.posy = U.cpy;
	T.mip = U.mip; T.zi = U.zi; T.dzi = U.dzi; T.mapswitch = U.mapswitch;
	const int n = dda_serial_batch_inl(T, rec, last_map, zfar_i);
	__syncwarp();
	const float v = gl == 0 ? T.d0 : gl == 1 ? T.i0x : gl == 2 ? T.i0y : gl == 3 ? T.d1 : gl == 4 ? T.i1x : T.i1y;
	const float g = gl == 0 ? T.gd0 : gl == 1 ? T.g0x : gl == 2 ? T.g0y : gl == 3 ? T.gd1 : gl == 4 ? T.g1x : T.g1y;
	if (gl < 6) dda_var_init(var, v, g);
	U.mip = T.mip; U.zi = T.zi; U.dzi = T.dzi; U.mapswitch = T.mapswitch;
	U.csd = T.index ? -T.dist_now : T.dist_now; U.cpx = T.posx; U.cpy = T.posy;
	return n;
}

// One batch.  sm: >= 112 shared words of this warp (48 plan + 64 table); rec: the serial hand-over records (fallback).
// Returns the crossings made (0 with ended = true at z_far); ra / rb = record before / after this lane's crossing.
// closed_ok = false (NaN ray) sends every batch through the fallback.
__device__ __forceinline__ int dda_closed_batch(DdaVar& var, DdaUni& U, int* sm, float4* rec, int last_map, int zfar_i, int gl,
                                                bool closed_ok, float4& ra, float4& rb, bool& ended)
{
	const unsigned FULL = 0xffffffffu;
	int want = 0;
	bool use_serial = !closed_ok;
	DdaPlan p;
	int F2 = 0, L2 = 0;
	if (closed_ok)
	{
		while (U.zi > U.mapswitch)                                      // Cuda_Render.h:343-365
		{
			if (U.mip < last_map) U.mip++;
			if (gl < 6) dda_var_double(var);
			U.mapswitch *= 2; U.dzi *= 2;
		}
		const int lod_free = (U.mapswitch - U.zi) / U.dzi + 1;          // crossings before z > mapswitch
		const int far_free = (zfar_i - U.zi) / U.dzi;                   // crossings with z + dz <= z_far
		if (far_free <= 0) { ended = true; return 0; }
		want = 32;
		want = want < lod_free ? want : lod_free;
		want = want < far_free ? want : far_free;
		int* plan = sm;            // [6][8]
		int* T0 = sm + 48;         // [32]
		int* T1 = sm + 80;         // [32]
		if (gl < 6)
		{
			dda_var_plan(var, 33, p, F2, L2);
			// a track reaches as far as the shortest of its three variables
			p.V = __reduce_min_sync(gl < 3 ? 0x07u : 0x38u, p.V);
			int4* q = reinterpret_cast<int4*>(plan + gl * 8);
			q[0] = make_int4(p.b0, p.c1, p.D1, p.L1);
			q[1] = make_int4(p.bx, p.c2, p.D2, p.V);
		}
		__syncwarp();
		int V0, V1;
		{
			const int4 a0 = reinterpret_cast<const int4*>(plan)[0], a1 = reinterpret_cast<const int4*>(plan)[1];
			const int4 b0 = reinterpret_cast<const int4*>(plan)[6], b1 = reinterpret_cast<const int4*>(plan)[7];
			DdaPlan pa, pb;
			pa.b0 = a0.x; pa.c1 = a0.y; pa.D1 = a0.z; pa.L1 = a0.w; pa.bx = a1.x; pa.c2 = a1.y; pa.D2 = a1.z; pa.V = a1.w;
			pb.b0 = b0.x; pb.c1 = b0.y; pb.D1 = b0.z; pb.L1 = b0.w; pb.bx = b1.x; pb.c2 = b1.y; pb.D2 = b1.z; pb.V = b1.w;
			V0 = pa.V; V1 = pb.V;
			T0[gl] = gl <= V0 ? dda_plan_eval(pa, gl) : RLERC_DDA_INF_BITS;
			T1[gl] = gl <= V1 ? dda_plan_eval(pb, gl) : RLERC_DDA_INF_BITS;
		}
		__syncwarp();
		const int i = dda_merge_search(T0, T1, gl), j = gl - i;
		const unsigned bad = __ballot_sync(FULL, i > V0 || j > V1);
		const int first_bad = bad ? (__ffs(bad) - 1) : 32;
		if (first_bad < want) use_serial = true;
		else
		{
			const int d0 = T0[i], d1 = T1[j];
			const bool t1 = d1 < d0;                                       // Cuda_Render.h:398 (non-negative floats order like their bits)
			const int idx = t1 ? j : i;
			const int4* q = reinterpret_cast<const int4*>(plan) + (t1 ? 8 : 2);
			const int4 x0 = q[0], x1 = q[1], y0 = q[2], y1 = q[3];
			DdaPlan px, py;
			px.b0 = x0.x; px.c1 = x0.y; px.D1 = x0.z; px.L1 = x0.w; px.bx = x1.x; px.c2 = x1.y; px.D2 = x1.z; px.V = x1.w;
			py.b0 = y0.x; py.c1 = y0.y; py.D1 = y0.z; py.L1 = y0.w; py.bx = y1.x; py.c2 = y1.y; py.D2 = y1.z; py.V = y1.w;
			const float d = __int_as_float(t1 ? d1 : d0);
			rb = make_float4(t1 ? -d : d, __int_as_float(dda_plan_eval(px, idx)), __int_as_float(dda_plan_eval(py, idx)), __int_as_float(U.mip));
			ra.x = __shfl_up_sync(FULL, rb.x, 1); ra.y = __shfl_up_sync(FULL, rb.y, 1); ra.z = __shfl_up_sync(FULL, rb.z, 1);
			ra.w = 0.0f;
			if (gl == 0) ra = make_float4(U.csd, U.cpx, U.cpy, 0.0f);
			const int last = want - 1;
			U.csd = __shfl_sync(FULL, rb.x, last); U.cpx = __shfl_sync(FULL, rb.y, last); U.cpy = __shfl_sync(FULL, rb.z, last);
			const int n0 = __shfl_sync(FULL, i + (t1 ? 0 : 1), last);
			if (gl < 6) dda_var_advance(var, p, F2, L2, gl < 3 ? n0 : want - n0);
			U.zi += want * U.dzi;
			__syncwarp();
			return want;
		}
	}
	// serial fallback (also advances through LOD switches and stops at z_far by itself)
	const int n = dda_closed_fallback(var, U, rec, last_map, zfar_i, gl);
	if (n < 32 && (zfar_i - U.zi) / U.dzi <= 0) ended = true;
	ra = rec[gl]; rb = rec[gl + 1];
	__syncwarp();
	return n;
}

} // namespace rlerc
