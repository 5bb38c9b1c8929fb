// Device helpers shared by the traversal kernels (kernels.cu, traverse_warp.cu).
#pragma once
#include <stdint.h>
#include <limits.h>
#include "kernels.cuh"

namespace rlerc {

// Stores into the warped ray buffer.  RLERC_STREAM_STORES: cache-streaming (evict-first) stores, so that the 56-225 MB
// the traversal writes per frame do not push the compressed scene (gathered on the ray planes' critical chain) out of L2.
#ifndef RLERC_STREAM_STORES
#define RLERC_STREAM_STORES 0
#endif
__device__ __forceinline__ void st_warp(uint32_t* p, uint32_t v)
{
#if RLERC_STREAM_STORES
	__stcs(p, v);
#else
	*p = v;
#endif
}


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
inline int owned_count(int count, int block, int n, int rank)
{
	if (n <= 1) return count;
	const int cyc = block * n;
	int owned = (count / cyc) * block;
	int rem = count % cyc - rank * block;
	if (rem > block) rem = block;
	if (rem > 0) owned += rem;
	return owned;
}


// ---- per-ray set-up shared by all traversal kernels ------------------------------------------
struct RayInit {
	float ray_x, ray_z;      // ray direction in the x-z plane        (Cuda_Render.h:165-166)
	float rx2mr;             // res_x2_mul_reverse                    (Cuda_Render.h:208-210)
	int ycmin, ycmax;        // y_clip_min / y_clip_max of the ray row (Cuda_Render.h:228-248)
	bool vertical;           // ml_direction_y                        (Cuda_Render.h:171)
	bool skip;               // one of the early returns at Cuda_Render.h:226,250 fired
};

// A (Cuda_Render.h:128-172) + B (Cuda_Render.h:203-250) of SURVEY.md §3.3.
__device__ __forceinline__ void ray_init(const TraverseParams& P, int x, RayInit& o)
{
	const int res_x = P.res_x, res_y = P.res_y;
	const float res_x2 = (float)(res_x / 2);             // Cuda_Render.h:107 (integer division)
	float s2x, s2y, e2x, e2y;
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
		o.ray_x = P.cos_my * dx + P.sin_my * dz;
		o.ray_z = P.cos_my * dz - P.sin_my * dx;
		s2x = p1x; s2y = p1y; e2x = p2x; e2y = p2y;
		o.vertical = (q < 2);
	}
	bool reverse = false;
	if (o.vertical) { if (o.ray_z <= 0) reverse = true; }
	else
	{
		if (o.ray_x <= 0) { if (P.sin_x > 0) reverse = true; }
		if (o.ray_x > 0) { if (P.sin_x < 0) reverse = true; }
	}
	float rx2mr = reverse ? -res_x2 : res_x2;
	if (o.vertical) rx2mr = -rx2mr;
	o.rx2mr = rx2mr;
	const int p_add = reverse ? 1 : -2;
	int q1x = f2i((float)res_x * s2x) + p_add;
	int q1y = f2i((float)res_y * s2y) + p_add;
	int q2x = f2i((float)res_x * e2x) - p_add;
	int q2y = f2i((float)res_y * e2y) - p_add;
	if (q1x < 0) q1x = 0; if (q1x >= res_x) q1x = res_x - 1;
	if (q1y < 0) q1y = 0; if (q1y >= res_y) q1y = res_y - 1;
	if (q2x < 0) q2x = 0; if (q2x >= res_x) q2x = res_x - 1;
	if (q2y < 0) q2y = 0; if (q2y >= res_y) q2y = res_y - 1;
	bool skip = (q1y == q2y);                           // Cuda_Render.h:226
	int ycmin = res_x - 1 - q1x;
	int ycmax = res_x - 1 - q2x;
	if (o.vertical) { ycmin = res_y - 1 - q1y; ycmax = res_y - 1 - q2y; }
	if (reverse) { ycmin = res_y - 1 - ycmin; ycmax = res_y - 1 - ycmax; }
	if (ycmin > ycmax) { const int t = ycmin; ycmin = ycmax; ycmax = t; }
	if (ycmin >= ycmax) skip = true;                    // Cuda_Render.h:250
	o.ycmin = ycmin; o.ycmax = ycmax; o.skip = skip;
}

// Texels outside the clip range are never written by the reference (stale from the previous
// frame) yet the unwarp samples a few of them; they are defined as 0 here so
// that a frame does not depend on history.  Whole row when the ray is skipped.
template <int G>
__device__ __forceinline__ void clear_outside(uint32_t* row, int res_y, const RayInit& r, int gl)
{
	const int lo = r.skip ? res_y : r.ycmin, hi = r.skip ? res_y - 1 : r.ycmax;
	for (int y = gl; y < lo; y += G) st_warp(row + y, 0u);
	for (int y = hi + 1 + gl; y < res_y; y += G) st_warp(row + y, 0u);
}

// D (Cuda_Render.h:270-305): 2-D DDA over the x-z pointer map
struct Dda {
	float g0x, g0y, g1x, g1y;    // grad0, grad1
	float i0x, i0y, i1x, i1y;    // isect0, isect1
	float gd0, gd1, d0, d1;      // grad_dist0/1, dds_dist0/1
	int fixx, fixz;              // fix.x, fix.y as integers (0 or -1)
};

__device__ __forceinline__ void dda_init(const TraverseParams& P, float ray_x, float ray_z, Dda& d)
{
	const float vpx = P.viewpos[0], vpz = P.viewpos[2];
	const float drx = ray_x * P.cos_y + ray_z * P.sin_y;
	const float dry = ray_x * P.sin_y - ray_z * P.cos_y;
	float fx = vpx - (float)f2i(vpx);
	float fy = vpz - (float)f2i(vpz);
	float sgx = -1, sgy = -1;
	d.fixx = -1; d.fixz = -1;
	if (drx >= 0) { d.fixx = 0; sgx = 1; fx = 1 - fx; }
	if (dry >= 0) { d.fixz = 0; sgy = 1; fy = 1 - fy; }
	d.g0y = dry / fabsf(drx); d.g0x = sgx;
	d.g1x = drx / fabsf(dry); d.g1y = sgy;
	d.i0x = d.g0x * fx; d.i0y = d.g0y * fx;
	d.i1x = d.g1x * fy; d.i1y = d.g1y * fy;
	d.gd0 = sqrtf(d.g0x * d.g0x + d.g0y * d.g0y);
	d.gd1 = sqrtf(d.g1x * d.g1x + d.g1y * d.g1y);
	d.d0 = sqrtf(d.i0x * d.i0x + d.i0y * d.i0y);
	d.d1 = sqrtf(d.i1x * d.i1x + d.i1y * d.i1y);
}

} // namespace rlerc
