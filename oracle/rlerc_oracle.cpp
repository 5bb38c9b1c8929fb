/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  The oracle ("port") of the hot path.
 *
 * A plain, serial, CPU restatement of the reference's algorithm for the frame loop, one
 * function per reference function, each citing the file:line it follows ("R/" =
 * RLE-Raycaster/ in the reference checkout).  It exists to (a) say what "correct" means for
 * the pieces of the path the reference only has as GPU code (the GLSL unwarp pass has no
 * executable reference here), (b) expose per-pixel hit identity (column, mip, voxel index),
 * which the reference computes but never stores, and (c) provide the work counters of the
 * byte model.
 *
 * PARITY PINNING.  The reference ships no tests, golden vectors or fixtures for this path
 *  — what pins this file is the reference ITSELF compiled for the host
 * (oracle/_ref, built by oracle/Makefile from /root/reference): tests/test_oracle_vs_ref.py
 * requires orc_render / orc_get_ray_map / orc_build_map to be bit-identical to it on a
 * camera x scene grid, and tests/golden/ holds hashes generated from oracle/_ref by
 * tests/golden/make_golden.py.  orc_unwarp restates GLSL and is pinned only by those
 * hashes of its own output ("parity unpinned" for the shading arithmetic).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Arithmetic: strict IEEE single precision, no FMA contraction (-ffp-contract=off),
 * float->int by C truncation == x86 cvttss2si (0x80000000 on overflow/NaN) — the semantics
 * the host-compiled reference has, which DESIGN.md §3 declares canonical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 { float x, y, z; };
struct OMap4 { int32_t sx, sy, sz, slabs_size; uint32_t* map; uint16_t* slabs; };
/* R/src/RayMap.h:16-54 on LP64 */
struct ORayMap {
	V3 vanishing_point_2d; int32_t map_line_count, map_line_limit; V3 rotation, position;
	float border, clip_min, clip_max; OMap4 map4_gpu[16]; int32_t nummaps, maxres, res[4];
	V3 p4, p_2d[8], p_no[8]; float to3d[4][4]; float p_ofs_min[4], p_ofs_max[4];
};
static_assert(sizeof(ORayMap) == 896, "RayMap_GPU layout");

inline V3 mk(float x, float y, float z) { V3 r = { x, y, z }; return r; }
inline V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }

/* float -> int exactly as the host-compiled reference does it (cvttss2si). Written out so the
 * behaviour does not depend on the compiler's treatment of out-of-range casts. */
inline int f2i(float f)
{
	if (!(std::fabs(f) < 2147483648.0f)) return INT32_MIN;
	return (int)f;
}

/* R/src/Cuda_Render.h:39-52 */
float LineScale(V3 input, V3 center, float clip_max, float clip_min)
{
	float scale_x = 1, scale_y = 1;
	if (center.x > 1) scale_x = (1 - input.x) / (center.x - input.x);
	if (center.x < 0) scale_x = input.x / (input.x - center.x);
	if (center.y > clip_max) scale_y = (clip_max - input.y) / (center.y - input.y);
	if (center.y < clip_min) scale_y = (-clip_min + input.y) / (input.y - center.y);
	return (scale_x < scale_y) ? scale_x : scale_y;
}

/* R/src/Cuda_Render.h:54-65 */
void ClipLine(V3& p1, V3& p2, float clip_max, float clip_min)
{
	float scale = LineScale(p1, p2, clip_max, clip_min);
	const V3 c2 = add(p1, mul(sub(p2, p1), scale));
	scale = LineScale(p2, p1, clip_max, clip_min);
	const V3 c1 = add(p2, mul(sub(p1, p2), scale));
	p1 = c1; p2 = c2;
}

/* R/src/Cuda_Render.h:67-73 */
V3 MatMul(const float m[4][4], V3 v)
{
	return mk(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z + m[3][0],
	          m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z + m[3][1],
	          m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z + m[3][2]);
}

struct Counters {   /* R/src/Cuda_Render.h:14-22 + byte-model terms */
	long long elems_total, elems_processed, voxels_processed, elems_rendered, pixels;
	long long cols_fetched, run_iters, cols_nonempty, cleared, dda_steps;
};

/* One ray plane: R/src/Cuda_Render.h:96-737 with the active configuration of R/src/core.h
 * (BUDDHA, FLOATING_HORIZON, XFLOATING_HORIZON, SHAREMEMCLIP; everything else off, and
 * without DETAIL_BENCH, i.e. with the early return at :370).  y_cache: private occlusion
 * bitmask of >= res_y bits, zeroed in full by the caller.  ids: optional uint32[res_y][2]. */
void render_line(const ORayMap& rm, int x, uint32_t* y_cache, int res_x, int res_y, int MIP_DISTANCE, int RAYS_DISTANCE,
                 uint32_t* ofs_rgb_start, uint32_t* ids, Counters& cnt, int flags)
{
	const V3 viewpos = rm.position, viewrot = rm.rotation;
	const float res_x2 = res_x / 2;                      /* :107-108 integer division */
	const float res_y2 = res_y / 2;

	/* :128-172 ray set-up */
	float ml_ray_x, ml_ray_z; V3 ml_start2d, ml_end2d; bool ml_direction_y;
	{
		int rays[4];
		rays[0] = rm.res[0]; rays[1] = rm.res[1] + rays[0]; rays[2] = rm.res[2] + rays[1]; rays[3] = rm.res[3] + rays[2];
		int quadrant = 0;
		if (x >= rays[2]) quadrant = 3; else if (x >= rays[1]) quadrant = 2; else if (x >= rays[0]) quadrant = 1;
		float quadrant_ofs = x;
		if (quadrant >= 1) quadrant_ofs -= rays[quadrant - 1];
		const float quadrant_num = rm.res[quadrant];
		const float a = quadrant_ofs / quadrant_num;
		const int j = quadrant;
		V3 p1 = rm.p_2d[5];
		V3 p2 = add(rm.p_no[j * 2], mul(sub(rm.p_no[j * 2 + 1], rm.p_no[j * 2]), a));
		ClipLine(p1, p2, rm.clip_max, rm.clip_min);
		const V3 p1_3d = MatMul(rm.to3d, mul(p1, 4.0f));
		const V3 p2_3d = MatMul(rm.to3d, mul(p2, 4.0f));
		V3 delta = sub(mul(add(p1_3d, p2_3d), 0.5f), rm.p4);
		delta.y = 0;
		/* normalize(): v * rsqrtf(dot(v,v)) with the host fallback rsqrtf = 1.0f/sqrtf (R/inc/cutil_math.h:58-61,1184-1188) */
		const float inv = 1.0f / sqrtf(delta.x * delta.x + delta.y * delta.y + delta.z * delta.z);
		delta = mul(delta, inv);
		/* vec3f_rot_y(viewrot.y) :75-80 */
		const float a_ = viewrot.y;
		const float xx = std::cos(-a_) * delta.x + std::sin(-a_) * delta.z;
		const float zz = std::cos(-a_) * delta.z - std::sin(-a_) * delta.x;
		ml_ray_x = xx; ml_ray_z = zz;
		ml_start2d = p1; ml_end2d = p2;
		ml_direction_y = 1 - (j >> 1);
	}

	int mip_lvl = 0;
	int y_clip_min = 0, y_clip_max = res_y - 1;
	const int z_far = RAYS_DISTANCE;
	float dz = 1 << mip_lvl;
	int mapswitch = MIP_DISTANCE;

	const float sin_x = std::sin(rm.rotation.x), cos_x = std::cos(rm.rotation.x);   /* :188-191 */
	const float sin_y = std::sin(rm.rotation.y), cos_y = std::cos(rm.rotation.y);
	const float ray_x = ml_ray_x, ray_z = ml_ray_z;
	const bool vertical = ml_direction_y;

	bool reverse = false;                                                              /* :203-210 */
	if (vertical) if (ray_z <= 0) reverse = true;
	if (!vertical) if (ray_x <= 0) if (sin_x > 0) reverse = true;
	if (!vertical) if (ray_x > 0) if (sin_x < 0) reverse = true;
	float res_x2_mul_reverse = reverse ? -res_x2 : res_x2;
	if (vertical) res_x2_mul_reverse = -res_x2_mul_reverse;

	{	/* :215-250 screen-space clipping */
		const int p_add = reverse ? 1 : -2;
		int p1x = f2i(float(float(res_x) * ml_start2d.x)) + p_add;
		int p1y = f2i(float(float(res_y) * ml_start2d.y)) + p_add;
		int p2x = f2i(float(float(res_x) * ml_end2d.x)) - p_add;
		int p2y = f2i(float(float(res_y) * ml_end2d.y)) - p_add;
		if (p1x < 0) p1x = 0;
		if (p1x >= res_x) p1x = res_x - 1;
		if (p1y < 0) p1y = 0;
		if (p1y >= res_y) p1y = res_y - 1;
		if (p2x < 0) p2x = 0;
		if (p2x >= res_x) p2x = res_x - 1;
		if (p2y < 0) p2y = 0;
		if (p2y >= res_y) p2y = res_y - 1;
		if (p1y == p2y) return;
		y_clip_min = res_x - 1 - p1x;
		y_clip_max = res_x - 1 - p2x;
		if (vertical) { y_clip_min = res_y - 1 - p1y; y_clip_max = res_y - 1 - p2y; }
		if (reverse) { y_clip_min = res_y - 1 - y_clip_min; y_clip_max = res_y - 1 - y_clip_max; }
		if (y_clip_min > y_clip_max) { const int tmp = y_clip_min; y_clip_min = y_clip_max; y_clip_max = tmp; }
		if (y_clip_min >= y_clip_max) return;
	}

	for (int n = y_clip_min; n <= y_clip_max; n++) ofs_rgb_start[n] = 0xff8844;          /* :255-261 */
	cnt.cleared += y_clip_max - y_clip_min + 1;

	/* :270-305 DDA set-up */
	float dirx = ray_x * cos_y + ray_z * sin_y;
	float diry = ray_x * sin_y - ray_z * cos_y;
	float fixx = -1, fixy = -1, signx = -1, signy = -1;
	float fracx = viewpos.x - f2i(viewpos.x);
	float fracy = viewpos.z - f2i(viewpos.z);
	if (dirx >= 0) { fixx = 0; signx = 1; fracx = 1 - fracx; }
	if (diry >= 0) { fixy = 0; signy = 1; fracy = 1 - fracy; }
	float grad0x = signx, grad0y = diry / std::fabs(dirx);
	float grad1x = dirx / std::fabs(diry), grad1y = signy;
	float isect0x = grad0x * fracx, isect0y = grad0y * fracx;
	float isect1x = grad1x * fracy, isect1y = grad1y * fracy;
	float grad_dist0 = sqrtf(grad0x * grad0x + grad0y * grad0y);
	float grad_dist1 = sqrtf(grad1x * grad1x + grad1y * grad1y);
	float dds_dist0 = sqrtf(isect0x * isect0x + isect0y * isect0y);
	float dds_dist1 = sqrtf(isect1x * isect1x + isect1y * isect1y);
	float pos_before_x = 0, pos_before_y = 0, dds_dist_before = 0;
	float pos_x = 0, pos_y = 0, dds_dist_now = 0;
	int index = 0, index_before = 0;

	int rle4_gridx = rm.map4_gpu[mip_lvl].sx, rle4_gridz = rm.map4_gpu[mip_lvl].sz;
	const float pos3d_z_add = sin_x;
	float pos3d_y_add = vertical ? cos_x : 0;
	pos3d_y_add *= res_x2_mul_reverse;
	const uint32_t* map_ptr = rm.map4_gpu[mip_lvl].map;
	const uint16_t* slab_ptr = rm.map4_gpu[mip_lvl].slabs;
	float z = 0;
	float y_map_switch = viewpos.y;
	mapswitch = mapswitch * (0.25 * (4 - std::abs(viewrot.x)));                          /* :335, double arithmetic */

	while (true)
	{
		while (z > mapswitch || (y_map_switch > 512.0))                                  /* :343-365 */
		{
			y_map_switch = y_map_switch * 0.5;
			if (mip_lvl < rm.nummaps - 1)
			{
				mip_lvl++;
				rle4_gridx >>= 1; rle4_gridz >>= 1;
				map_ptr = rm.map4_gpu[mip_lvl].map;
				slab_ptr = rm.map4_gpu[mip_lvl].slabs;
			}
			grad0x *= 2; grad0y *= 2; grad1x *= 2; grad1y *= 2;
			grad_dist0 *= 2; grad_dist1 *= 2;
			mapswitch *= 2;
			dz *= 2;
		}
		z += dz;
		if (z > z_far) return;
		if (y_clip_min >= y_clip_max) return;                                            /* :370 */
		cnt.dda_steps++;

		dds_dist_before = dds_dist_now;                                                  /* :376-414 */
		pos_before_x = pos_x; pos_before_y = pos_y;
		index_before = index;
		if (dds_dist1 < dds_dist0)
		{
			dds_dist_now = dds_dist1; index = 1; dds_dist1 += grad_dist1;
			pos_x = isect1x; pos_y = isect1y;
			isect1x += grad1x; isect1y += grad1y;
		}
		else
		{
			dds_dist_now = dds_dist0; index = 0;
			pos_x = isect0x; pos_y = isect0y;
			dds_dist0 += grad_dist0;
			isect0x += grad0x; isect0y += grad0y;
		}
		const int fix_x = (1 - index_before) * fixx, fix_z = (index_before) * fixy;   /* :418-419 */
		const float dds_dist_delta = dds_dist_now - dds_dist_before;
		const float view_space_x = ray_x * dds_dist_before, view_space_z = ray_z * dds_dist_before;
		const int voxel_x = f2i(viewpos.x + pos_before_x) + fix_x;                       /* :429-430 */
		const int voxel_z = f2i(viewpos.z + pos_before_y) + fix_z;
		if (flags & 1)                                                                   /* CLIPREGION :432-437 */
		{
			if (voxel_x < 0) continue;
			if (voxel_z < 0) continue;
			if ((voxel_x >> mip_lvl) > rle4_gridx - 1) continue;
			if ((voxel_z >> mip_lvl) > rle4_gridz - 1) continue;
		}
		const int vx = (voxel_x >> mip_lvl) & (rle4_gridx - 1);                          /* :441-442 */
		const int vz = (voxel_z >> mip_lvl) & (rle4_gridz - 1);
		const float mountain = viewpos.y;
		const float correct_x = ray_x * dds_dist_delta, correct_z = ray_z * dds_dist_delta;
		const float pos3d_z = cos_x * view_space_z + sin_x * mountain;                   /* :459-464 */
		float pos3d_y = vertical ? cos_x * mountain - sin_x * view_space_z : view_space_x;
		pos3d_y *= res_x2_mul_reverse;
		if (pos3d_z * res_y2 + pos3d_y <= pos3d_z * y_clip_min) continue;               /* :467 top clip */

		const uint32_t slab_offset = map_ptr[(size_t)(vx + vz * rle4_gridx) * 2];        /* :474-478 */
		const uint32_t len_first = map_ptr[(size_t)(vx + vz * rle4_gridx) * 2 + 1];
		const uint16_t slen = len_first;
		const float corr_zz = cos_x * correct_z;                                         /* :483-486 */
		float corr_yy = vertical ? -sin_x * correct_z : correct_x;
		corr_yy *= res_x2_mul_reverse;
		uint16_t sti_ = len_first >> 16;
		const uint16_t* slabs = slab_ptr + 2 + slab_offset;                               /* :498-499 */
		const uint16_t* const slabs_first = slabs;
		const uint16_t* send = slabs + slen;
		float tex = 0;
		int sti_general = 0, sti_skip = 0;
		cnt.cols_fetched++; cnt.elems_total += slen; if (slen) cnt.cols_nonempty++;

		for (; slabs < send; ++slabs)                                                     /* :512-734 */
		{
			cnt.run_iters++;
			if (slabs > slabs_first) sti_ = *slabs;                                       /* :517 (index compare, not truncated pointers) */
			sti_skip = (sti_ >> 10);
			sti_general += (sti_ & 1023) << mip_lvl;
			if (sti_skip == 0) continue;
			const int texture = tex;
			tex += sti_skip;
			sti_skip <<= mip_lvl;
			const float sti_general_sti_skip = sti_general;
			sti_general += sti_skip;
			cnt.elems_processed++; cnt.voxels_processed += sti_skip;

			float correct_zz1 = pos3d_z, correct_yy1 = pos3d_y;
			if (mountain + sti_general_sti_skip >= 0) { correct_zz1 += corr_zz; correct_yy1 += corr_yy; }
			const float pos3d_z1 = correct_zz1 + pos3d_z_add * sti_general_sti_skip; if (pos3d_z1 <= 0) continue;
			const float pos3d_y1 = correct_yy1 + pos3d_y_add * sti_general_sti_skip;
			int scr_y2 = f2i(res_y2 + pos3d_y1 / pos3d_z1);                               /* :542 */
			if (scr_y2 <= y_clip_min) break;

			float correct_zz2 = pos3d_z, correct_yy2 = pos3d_y;
			if (mountain + sti_general < 0) { correct_zz2 += corr_zz; correct_yy2 += corr_yy; }
			const float pos3d_z2 = correct_zz2 + pos3d_z_add * sti_general; if (pos3d_z2 <= 0) continue;
			const float pos3d_y2 = correct_yy2 + pos3d_y_add * sti_general;
			int scr_y1 = f2i(res_y2 + pos3d_y2 / pos3d_z2 - 1); if (scr_y1 >= y_clip_max) continue;   /* :560 */

			if (scr_y2 >= y_clip_max) { scr_y2 = y_clip_max; y_clip_max = scr_y1; }       /* :564-580 */
			if (scr_y1 <= y_clip_min)
			{
				scr_y1 = y_clip_min;
				y_clip_min = scr_y2;
				while ((y_clip_max > y_clip_min) && (y_cache[y_clip_min >> 5] & (1u << (y_clip_min & 31)))) ++y_clip_min;
			}
			int y = scr_y1;
			while ((y < scr_y2) && (y_cache[y >> 5] & (1u << (y & 31)))) ++y;              /* :639-640 */
			if (y >= scr_y2) continue;

			const float pos3d_z1r = pos3d_z + pos3d_z_add * sti_general_sti_skip;          /* :645-660 */
			const float pos3d_y1r = pos3d_y + pos3d_y_add * sti_general_sti_skip;
			const float pos3d_z2r = pos3d_z + pos3d_z_add * sti_general;
			const float pos3d_y2r = pos3d_y + pos3d_y_add * sti_general;
			const float scr_y2r = res_y2 + pos3d_y1r / pos3d_z1r;
			const float scr_y1r = res_y2 + pos3d_y2r / pos3d_z2r;
			const float u1z = (tex) / pos3d_z2r;
			float u2dz = (texture) / pos3d_z1r - u1z;
			const float onez1 = 1 / pos3d_z2r;
			float onedz2 = 1 / pos3d_z1r - onez1;
			u2dz /= scr_y2r - scr_y1r;
			onedz2 /= scr_y2r - scr_y1r;
			cnt.elems_rendered++;

			const int height_color = f2i(4095 - mountain + viewpos.y);                    /* HEIGHT_COLOR :675 (float arithmetic, int target) */
			const float mult = y + 1 - scr_y1r;                                            /* :678-680 */
			float uz = u1z + u2dz * mult;
			float onez = onez1 + onedz2 * mult;
			for (; y < scr_y2; ++y, uz += u2dz, onez += onedz2)                            /* :687-733 */
			{
				const int y5 = y >> 5;
				const uint32_t y31 = 1u << (y & 31);
				if (y_cache[y5] & y31) continue;
				/* :707 in the host build binds to the int min/max of R/inc/cutil_math.h:48-56 */
				int ui = f2i(float(uz / onez));
				const int lo = f2i(float(texture)), hi = f2i(float(tex - 1.0));
				ui = ui > lo ? ui : lo;
				ui = ui < hi ? ui : hi;
				const uint32_t u = ui;
				const uint32_t real_z = f2i(float(1 / onez)) & 0xfffe;
				uint32_t color16 = send[u];
				if (flags & 2)                                                               /* HEIGHT_COLOR :716-722 */
				{
					const uint16_t colorpal = color16 & 0xff00;
					int v = (int)(((color16 & 0xff) * (uint32_t)height_color) >> 12);       /* uint * int is unsigned */
					v = v > 0 ? v : 0;                                                       /* int max / min of R/inc/cutil_math.h:48-56 */
					v = v < 255 ? v : 255;
					color16 = (uint32_t)v | colorpal;
				}
				ofs_rgb_start[y] = color16 + (real_z << 16);
				if (ids) { ids[y * 2] = (uint32_t)(vx + vz * rle4_gridx); ids[y * 2 + 1] = ((uint32_t)mip_lvl << 16) | u; }
				cnt.pixels++;
				y_cache[y5] |= y31;
			}
		}
	}
}

/* GLSL step(edge, x) */
inline float stepf(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }
inline uint32_t quant8(float c)
{
	c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
	return (uint32_t)(int)(c * 255.0f + 0.5f);
}

} /* namespace */

extern "C" {

/* RLE4::load's pointer-map rebuild, R/src/Rle4.cpp:284-314. map: uint32[sx*sz*2]. */
int orc_build_map(const uint16_t* slabs, unsigned long long slabs_size, int sx, int sz, uint32_t* map)
{
	memset(map, 0, (size_t)sx * sz * 8);
	unsigned long long ofs = 0;
	int x = 0, z = 0;
	while (1)
	{
		if (ofs + 2 > slabs_size) return -1;
		map[((size_t)x + (size_t)z * sx) * 2 + 0] = (uint32_t)ofs;
		const uint32_t count = slabs[ofs];
		uint32_t firstrle = 0;
		if (ofs + 2 < slabs_size) firstrle = slabs[ofs + 2];
		map[((size_t)x + (size_t)z * sx) * 2 + 1] = count + (firstrle << 16);
		x = (x + 1) % sx;
		if (x == 0) { z = (z + 1) % sz; if (z == 0) break; }
		ofs += slabs[ofs] + slabs[ofs + 1] + 2;
	}
	return 0;
}

/* All ray planes [ray_begin, ray_end) of one frame, as cudaRender does (R/src/Cuda_Main.cu:150-181).
 * raymap: RayMap_GPU bytes with HOST pointers in map4_gpu.  warp: uint32[rays][res]; ids: optional
 * uint32[rays][res][2]; counters: optional long long[10]. */
int orc_render_flags(const void* raymap, int res, int mip_distance, int z_far, uint32_t* warp, uint32_t* ids,
                     long long* counters, int ray_begin, int ray_end, int threads, int flags);

int orc_render(const void* raymap, int res, int mip_distance, int z_far, uint32_t* warp, uint32_t* ids,
               long long* counters, int ray_begin, int ray_end, int threads)
{
	return orc_render_flags(raymap, res, mip_distance, z_far, warp, ids, counters, ray_begin, ray_end, threads, 0);
}

/* flags: 1 = CLIPREGION, 2 = HEIGHT_COLOR (R/src/core.h:18,22) */
int orc_render_flags(const void* raymap, int res, int mip_distance, int z_far, uint32_t* warp, uint32_t* ids,
                     long long* counters, int ray_begin, int ray_end, int threads, int flags)
{
	const ORayMap& rm = *(const ORayMap*)raymap;
	if (ray_end < 0 || ray_end > rm.map_line_count) ray_end = rm.map_line_count;
	if (ray_begin < 0) ray_begin = 0;
	const int words = (res + 31) / 32 + 2;
	Counters total;
	memset(&total, 0, sizeof(total));
#ifdef _OPENMP
	if (threads > 0) omp_set_num_threads(threads);
#endif
	#pragma omp parallel
	{
		uint32_t* mask = (uint32_t*)malloc(words * sizeof(uint32_t));
		Counters c;
		memset(&c, 0, sizeof(c));
		#pragma omp for schedule(dynamic, 16)
		for (int x = ray_begin; x < ray_end; x++)
		{
			memset(mask, 0, words * sizeof(uint32_t));
			render_line(rm, x, mask, res, res, mip_distance, z_far, warp + (size_t)x * res,
			            ids ? ids + (size_t)x * res * 2 : 0, c, flags);
		}
		#pragma omp critical
		{
			long long* t = (long long*)&total; const long long* s = (const long long*)&c;
			for (int k = 0; k < 10; k++) t[k] += s[k];
		}
		free(mask);
	}
	if (counters) memcpy(counters, &total, sizeof(total));
	return 0;
}

/* Unwarp + shade: R/bin/shader/colorize_buddha_soft.frag:12-137 with the uniforms of
 * R/src/main.cpp:578-603; texture fetch GL_NEAREST + CLAMP_TO_EDGE (R/src/GL_Main.cpp:154-157)
 * from the warped buffer seen as an RGBA8 texture RS wide x RC tall; fragment centres at +0.5,
 * GL origin bottom-left.  Output RGBA8 [H][W][4] with row 0 = TOP of the window, the 8-bit
 * quantisation of the GL_RGBA8 FBO it renders into (R/src/GL_Main.h:151): round(clamp(c)*255). */
static int32_t* g_texel_out = 0;   /* optional int32[H][W][2] = {ray row iy, texel ix} chosen per pixel */
void orc_unwarp_set_texel_output(int32_t* p) { g_texel_out = p; }
/* 0: colorize_buddha_soft.frag (the shipped configuration); 1: colorize_buddha_soft_2xAA.frag (the -DANTIALIAS build,
 * R/src/main.cpp:510-522): the same texel geometry without the RAYS_CASTED_RES / RAYS_CASTED factor (frag:62), its own
 * shading (frag:79-139: 5:5:5 normal from the attribute, point light, two-term tone curve, vertical sky gradient). */
static int g_shader = 0;
void orc_unwarp_set_shader(int s) { g_shader = s; }

int orc_unwarp(const void* raymap, int W, int H, int RS, int RC, int rays_casted_res, const uint32_t* warp, uint8_t* rgba,
               int ray_begin, int ray_end)
{
	const ORayMap& rm = *(const ORayMap*)raymap;
	const float border_u = rm.border;
	const float vanish_x = 1 - rm.vanishing_point_2d.x;                                  /* main.cpp:579-584 */
	const float vanish_y = (1 - rm.vanishing_point_2d.y - border_u) * float(W) / float(H);
	const float ofs1 = 4 * float(rm.res[0]) / float(rays_casted_res);                    /* main.cpp:586-588 */
	const float ofs2 = 4 * float(rm.res[1]) / float(rays_casted_res) + ofs1;
	const float ofs3 = 4 * float(rm.res[2]) / float(rays_casted_res) + ofs2;
	const float ofs_add[4] = { -rm.p_ofs_min[0], -rm.p_ofs_min[1] + ofs1, -rm.p_ofs_min[2] + ofs2, -rm.p_ofs_min[3] + ofs3 };
	const float ratio = float(rays_casted_res) / float(RC);                              /* main.cpp:598-603 */
	const float rg = (rm.rotation.x > 0) ? 1.0f : 0.0f;                                  /* main.cpp:590 */
	const float RESX = (float)W, RESY = (float)H;
	#pragma omp parallel for schedule(static)
	for (int row = 0; row < H; row++)
	{
		const int pyg = H - 1 - row;
		for (int px = 0; px < W; px++)
		{
			const float fx = (float)px + 0.5f, fy = (float)pyg + 0.5f;
			const float scx = fx / RESX, scy = fy / RESY;                                /* frag:19-20 */
			const float border = (RESX - RESY) / (RESX * 2);                             /* frag:22 */
			const float scx1 = scx - vanish_x, scy1 = scy - vanish_y;                    /* frag:24-25 */
			const float upper = stepf(scy1, 0.0f), left = stepf(scx1, 0.0f);             /* frag:27-28 */
			const float ostep = stepf(std::fabs(scy1) - std::fabs(scx1) * RESX / RESY, 0.0f);
			const float seg_up = (1 - upper) * (1 - ostep), seg_dn = (upper) * (1 - ostep);   /* frag:31-34 */
			const float seg_rt = (1 - left) * (ostep), seg_lt = (left) * (ostep);
			const float o2 = (ostep * fx + (1 - ostep) * fy) / RESX;                     /* frag:36 */
			const float ang2 = scx1 * std::fabs(1 - upper - vanish_y) / scy1 + upper * (1 - vanish_x) + (1 - upper) * (vanish_x);
			float ang3 = scy1 * std::fabs(1 - left - vanish_x) / scx1 + left * (1 - vanish_y) + (1 - left) * (vanish_y);
			ang3 = ang3 * RESY / RESX + border;                                          /* frag:46 */
			const float x_pre = (ostep * ang3 + ang2 * (1 - ostep));                     /* frag:53 */
			float ty = seg_dn * (ofs_add[1] + x_pre) + seg_up * (ofs_add[0] + 1.0f - x_pre) +
			           seg_lt * (ofs_add[3] + x_pre) + seg_rt * (ofs_add[2] + 1.0f - x_pre);   /* frag:56-60 */
			ty = (g_shader == 1) ? ty * 0.25f : ty * ratio * 0.25f;                      /* frag:62 (2xAA frag:62: no ratio) */
			const float seg_up_x = rg * seg_up + (1.0f - rg) * seg_dn;                   /* frag:68-71 */
			const float seg_dn_x = rg * seg_dn + (1.0f - rg) * seg_up;
			const float seg_rt_x = rg * seg_rt + (1.0f - rg) * seg_lt;
			const float seg_lt_x = rg * seg_lt + (1.0f - rg) * seg_rt;
			const float tx = (seg_up_x) * (o2 + border) + (seg_dn_x) * (1.0f - (o2 + border)) +
			                 (seg_rt_x) * (o2) + (seg_lt_x) * (1.0f - o2);               /* frag:73-77 */
			int ix = f2i(floorf(tx * (float)RS)), iy = f2i(floorf(ty * (float)RC));
			ix = ix < 0 ? 0 : (ix >= RS ? RS - 1 : ix);
			iy = iy < 0 ? 0 : (iy >= RC ? RC - 1 : iy);
			uint8_t* o = rgba + ((size_t)row * W + px) * 4;
			if (g_texel_out) { g_texel_out[((size_t)row * W + px) * 2] = iy; g_texel_out[((size_t)row * W + px) * 2 + 1] = ix; }
			if (ray_end >= 0 && (iy < ray_begin || iy >= ray_end)) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
			const uint32_t t = warp[(size_t)iy * RS + ix];
			const float cr = (float)(t & 255u) / 255.0f, cg = (float)((t >> 8) & 255u) / 255.0f;
			const float cb = (float)((t >> 16) & 255u) / 255.0f, ca = (float)(t >> 24) / 255.0f;
			float r, g, b, fragz = 0.0f;
			if (g_shader == 1)                                                           /* colorize_buddha_soft_2xAA.frag:79-139 */
			{
				const int x1 = f2i(cr * 255.0f);                                         /* 2xAA frag:84-88: int(c.r*255.0) truncates */
				const int x2 = f2i(cg * 255.0f) * 256 + x1;
				const float col16b = float(x2 & 31) / 31.0f, col16g = float((x2 >> 5) & 31) / 31.0f, col16r = float((x2 >> 10) & 31) / 31.0f;
				if (cb != 1.0f)
				{
					const float z = (cb * (1.0f / 256.0f) + ca);                          /* :103-106 */
					fragz = 0.001f / z;
					const float pos3dx = z * (scx * 2.0f - 1.0f), pos3dy = z * (scy * 2.0f - 1.0f);
					const float nx = 2.0f * col16r - 1.0f, ny = 2.0f * col16g - 1.0f, nz = 2.0f * col16b - 1.0f;   /* :108-113 */
					float lx = pos3dx - 10.0f, ly = pos3dy + 5.0f, lz = z;               /* :115-121 */
					const float inv = 1.0f / sqrtf(lx * lx + ly * ly + lz * lz);         /* normalize = v * (1/sqrt(dot)) */
					lx *= inv; ly *= inv; lz *= inv;
					const float light = nx * lx + ny * ly + nz * lz;                     /* :123 */
					r = light * 0.9f + light * light * 0.5f;                              /* :131 */
					g = light * 0.6f + light * light * 0.4f;
					b = light * 0.3f + light * light * 0.3f;
				}
				else { r = 0.3f * (1.0f - scy) + 0.8f * scy; g = r; b = r; }              /* :134 */
				r = r * 1.1f; g = g * 1.1f; b = b * 1.1f;                                 /* :136 */
				o[0] = (uint8_t)quant8(r); o[1] = (uint8_t)quant8(g); o[2] = (uint8_t)quant8(b); o[3] = (uint8_t)quant8(fragz);
				continue;
			}
			if (cb != 1.0f)                                                              /* frag:89-121 */
			{
				const float zz = (cb * (1.0f / 256.0f) + ca);
				fragz = 0.001f / zz;
				const float light = (1.0f - cg) * 1.0f + (0.0f + cr) * 0.3f - 0.5f;
				const float pw = 1.2f * powf(light > 0.0f ? light : 0.0f, 4.0f);
				r = light * 1.3f + pw * 1.2f;
				g = light * 0.9f + pw * 1.2f;
				b = light * 0.7f + pw * 1.2f;
			}
			else { r = 178.0f / 255.0f; g = 204.0f / 255.0f; b = 1.0f; }                 /* frag:125-126 */
			o[0] = (uint8_t)quant8(r); o[1] = (uint8_t)quant8(g); o[2] = (uint8_t)quant8(b); o[3] = (uint8_t)quant8(fragz);
		}
	}
	return 0;
}


/* Depth-aware smoothing: R/bin/shader/soft.frag:1-75 on the pass-1 image as it sits in the lower-left
 * W x H corner of the reference's square GL_RGBA8 FBO texture (GL_LINEAR, CLAMP_TO_EDGE, R/src/GL_Main.h:151-164;
 * texCoord spans 0..W/fbo, 0..H/fbo, R/src/main.cpp:606-607,639-643).  Texels outside the window are never
 * rendered by the reference and read as 0.  in/out: RGBA8 [H][W][4], row 0 = top.  Parity unpinned (GLSL). */
namespace {
struct F4 { float x, y, z, w; };
inline F4 soft_texel(const uint8_t* in, int W, int H, int i, int j)
{
	F4 r = { 0, 0, 0, 0 };
	if (i >= W || j >= H) return r;
	const uint8_t* p = in + ((size_t)(H - 1 - j) * W + i) * 4;
	r.x = (float)p[0] / 255.0f; r.y = (float)p[1] / 255.0f; r.z = (float)p[2] / 255.0f; r.w = (float)p[3] / 255.0f;
	return r;
}
inline F4 soft_fetch(const uint8_t* in, int W, int H, int fbo, float x, float y)
{
	const float u = x * (float)fbo - 0.5f, v = y * (float)fbo - 0.5f;
	const float fu0 = floorf(u), fv0 = floorf(v);
	const float fu = u - fu0, fv = v - fv0;
	int i0 = f2i(fu0), j0 = f2i(fv0), i1 = i0 + 1, j1 = j0 + 1;
	const int hi = fbo - 1;
	i0 = i0 < 0 ? 0 : (i0 > hi ? hi : i0); i1 = i1 < 0 ? 0 : (i1 > hi ? hi : i1);
	j0 = j0 < 0 ? 0 : (j0 > hi ? hi : j0); j1 = j1 < 0 ? 0 : (j1 > hi ? hi : j1);
	const F4 a = soft_texel(in, W, H, i0, j0), b = soft_texel(in, W, H, i1, j0), c = soft_texel(in, W, H, i0, j1), d = soft_texel(in, W, H, i1, j1);
	F4 r;
	r.x = (a.x * (1.0f - fu) + b.x * fu) * (1.0f - fv) + (c.x * (1.0f - fu) + d.x * fu) * fv;
	r.y = (a.y * (1.0f - fu) + b.y * fu) * (1.0f - fv) + (c.y * (1.0f - fu) + d.y * fu) * fv;
	r.z = (a.z * (1.0f - fu) + b.z * fu) * (1.0f - fv) + (c.z * (1.0f - fu) + d.z * fu) * fv;
	r.w = (a.w * (1.0f - fu) + b.w * fu) * (1.0f - fv) + (c.w * (1.0f - fu) + d.w * fu) * fv;
	return r;
}
} /* namespace */

int orc_soft(int W, int H, const uint8_t* in, uint8_t* out)
{
	int fbo = 2048;                                                /* FBO fbo1(2048,2048), R/src/main.cpp:540 */
	while (fbo < W || fbo < H) fbo *= 2;
	float tap_x[7], tap_y[7];
	{	/* soft.frag:16: for (float a = 0; a < 3.1415*2.0; a += 3.1415*1.9/6.0), float like GLSL: 7 taps */
		float a = 0.0f;
		const float step = (3.1415f * 1.9f) / 6.0f, lim = 3.1415f * 2.0f;
		for (int k = 0; k < 7 && a < lim; k++, a += step) { tap_x[k] = std::sin(a) * 0.005f; tap_y[k] = std::cos(a) * 0.005f; }
	}
	/* soft.frag:38-39: for (float a = -1.0; a < 1.0; a += 2.0/5.0): five values, accumulated in float */
	float ofs[8]; int nofs = 0;
	for (float a = -1.0f; a < 1.0f; a += 2.0f / 5.0f) ofs[nofs++] = a;
	#pragma omp parallel for schedule(static)
	for (int row = 0; row < H; row++)
	for (int px = 0; px < W; px++)
	{
		const int j = H - 1 - row;
		const float tx = ((float)px + 0.5f) / (float)fbo, ty = ((float)j + 0.5f) / (float)fbo;
		F4 col = soft_fetch(in, W, H, fbo, tx, ty);                                   /* :8 */
		float radmax = col.w;                                                        /* :11-24 */
		for (int k = 0; k < 7; k++)
		{
			const float w = soft_fetch(in, W, H, fbo, tap_x[k] + tx, tap_y[k] + ty).w;
			radmax = radmax > w ? radmax : w;
		}
		const float rad = 0.0023f * radmax;                                          /* :28 */
		if (rad > 0.00008f)                                                          /* :31 */
		{
			F4 avg = col; float n = 1.0f;
			for (int ia = 0; ia < nofs; ia++)
			for (int ib = 0; ib < nofs; ib++)
			{
				const F4 cin = soft_fetch(in, W, H, fbo, tx + ofs[ia] * rad, ty + ofs[ib] * rad);
				if (cin.w >= radmax * 0.7f) { avg.x += cin.x; avg.y += cin.y; avg.z += cin.z; avg.w += cin.w; n += 1.0f; }
			}
			if (n > 6.0f) { const float s = 1.0f / n; col.x = avg.x * s; col.y = avg.y * s; col.z = avg.z * s; col.w = avg.w * s; }   /* :58-66 */
		}
		uint8_t* o = out + ((size_t)row * W + px) * 4;
		o[0] = (uint8_t)quant8(col.x); o[1] = (uint8_t)quant8(col.y); o[2] = (uint8_t)quant8(col.z); o[3] = (uint8_t)quant8(col.w);
	}
	return 0;
}

/* RayMap::get_ray_map, R/src/RayMap.h:98-402 (+ Nebula matrix44 helpers R/inc/mathlib/_matrix44.h:
 * rotate_x/y :529-560, translate :583-588, invert_simpler :431-441, m*v :863-869).
 * out: RayMap_GPU bytes; map4_gpu/nummaps are left untouched. */
void orc_get_ray_map(const float pos[3], const float rot[3], float border, int rays_casted_res, void* out)
{
	ORayMap& rm = *(ORayMap*)out;
	rm.rotation = mk(rot[0], rot[1], rot[2]);
	rm.position = mk(pos[0], pos[1], pos[2]);
	rm.border = border; rm.clip_min = border; rm.clip_max = 1 - border; rm.map_line_limit = rays_casted_res;
	V3 p[6] = { mk(1, 1, 1), mk(-1, 1, 1), mk(-1, -1, 1), mk(1, -1, 1), mk(0, 0, 0), mk(0, 0, 0) };
	float m[4][4] = { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };
	{ const float c = std::cos(rot[0]), s = std::sin(rot[0]);
	  for (int i = 0; i < 4; i++) { const float mi1 = m[i][1], mi2 = m[i][2]; m[i][1] = mi1 * c + mi2 * -s; m[i][2] = mi1 * s + mi2 * c; } }
	{ const float c = std::cos(rot[1]), s = std::sin(rot[1]);
	  for (int i = 0; i < 4; i++) { const float mi0 = m[i][0], mi2 = m[i][2]; m[i][0] = mi0 * c + mi2 * s; m[i][2] = mi0 * -s + mi2 * c; } }
	m[3][0] += 3; m[3][1] += 2; m[3][2] += 0;
	for (int i = 0; i < 5; i++) p[i] = MatMul(m, p[i]);
	const V3 down = mk(0, -1, 0);
	const V3 view = sub(mk((p[0].x + p[2].x) / 2, (p[0].y + p[2].y) / 2, (p[0].z + p[2].z) / 2), p[4]);
	float angle;
	{	/* vec3f::angle, R/src/VecMath.h:71-81 */
		const float dot = view.x * down.x + view.y * down.y + view.z * down.z;
		float len = sqrtf(view.x * view.x + view.y * view.y + view.z * view.z) * sqrtf(down.x * down.x + down.y * down.y + down.z * down.z);
		if (len == 0) len = 0.00001f;
		float input = dot / len;
		if (input < -1) input = -1;
		if (input > 1) input = 1;
		angle = std::acos(input);
	}
	const float alpha = float(M_PI) / 2 - angle;
	const float scale = 1 / std::sin(alpha);
	p[5] = add(p[4], mul(down, scale));
	const V3 e1 = sub(p[1], p[0]), e2 = sub(p[3], p[0]);
	const V3 nrm = mk(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
	V3 d[4] = { e1, e2, nrm, p[0] };
	for (int i = 0; i < 3; i++) d[i] = mul(d[i], 1 / (d[i].x * d[i].x + d[i].y * d[i].y + d[i].z * d[i].z));
	float t2[4][4];
	for (int i = 0; i < 4; i++) { t2[i][0] = d[i].x; t2[i][1] = d[i].y; t2[i][2] = d[i].z; t2[i][3] = 1.0f; }
	memcpy(rm.to3d, t2, sizeof(t2));
	{	/* invert_simpler */
		float t;
		t = t2[0][1]; t2[0][1] = t2[1][0]; t2[1][0] = t;
		t = t2[0][2]; t2[0][2] = t2[2][0]; t2[2][0] = t;
		t = t2[1][2]; t2[1][2] = t2[2][1]; t2[2][1] = t;
		const float m30 = -(t2[0][0] * t2[3][0] + t2[1][0] * t2[3][1] + t2[2][0] * t2[3][2]);
		const float m31 = -(t2[0][1] * t2[3][0] + t2[1][1] * t2[3][1] + t2[2][1] * t2[3][2]);
		t2[3][2] = -(t2[0][2] * t2[3][0] + t2[1][2] * t2[3][1] + t2[2][2] * t2[3][2]);
		t2[3][1] = m31; t2[3][0] = m30;
	}
	for (int i = 0; i < 6; i++) rm.p_2d[i] = MatMul(t2, p[i]);
	rm.vanishing_point_2d = rm.p_2d[5];
	const int maxres = rays_casted_res / 4;
	rm.maxres = maxres;
	const int safety = 2;
	const float ys_min = border, ys_max = 1 - border;
	const V3 plist[4] = { mk(-1, -1, 0), mk(1, -1, 0), mk(1, 1, 0), mk(-1, 1, 0) };
	const V3 plist2[4] = { mk(0, ys_min, 0), mk(1, ys_min, 0), mk(1, ys_max, 0), mk(0, ys_max, 0) };
	rm.res[0] = rm.res[1] = rm.res[2] = rm.res[3] = 0;
	rm.p_2d[5].z = 0;
	const V3 v = rm.p_2d[5];
	V3* n = rm.p_no;
	#define DOT(a, b) ((b).x * (a).x + (b).y * (a).y + (b).z * (a).z)   /* vec3f::dot(a): a.x*x + ... with this = first */
	#define AIM(c, num, den) add(v, mul(sub((c), v), std::abs((num) / (den))))
	#define SNAP(lo, hi, fld, q)                                                            \
		n[lo].fld = float(int(maxres * n[lo].fld) - safety) / maxres;                       \
		n[hi].fld = float(int(maxres * n[hi].fld) + safety) / maxres;                       \
		rm.res[q] = maxres * std::abs(n[lo].fld - n[hi].fld);                               \
		if (n[lo].fld - n[hi].fld > 0) rm.res[q] = 0;                                       \
		if (rm.res[q] > (maxres * 3)) rm.res[q] = (maxres * 3);                             \
		rm.p_ofs_min[q] = n[lo].fld; rm.p_ofs_max[q] = n[hi].fld;
	if (v.y > border)                                                                       /* :209-240 */
	{
		n[0] = add(v, mul(plist[0], std::abs(v.y - border)));
		n[1] = add(v, mul(plist[1], std::abs(v.y - border)));
		const V3 in = n[1];
		if (v.x > 1) { if (n[1].x > 1) n[1].x = 1; }
		else if (DOT(sub(n[0], v), sub(plist2[2], v)) > 0) n[1] = AIM(plist2[2], plist2[0].y - v.y, plist2[2].y - v.y);
		if (v.x < 0) { if (n[0].x < 0) n[0].x = 0; }
		else if (DOT(sub(in, v), sub(plist2[3], v)) > 0) n[0] = AIM(plist2[3], plist2[1].y - v.y, plist2[3].y - v.y);
		n[0].y = n[1].y = plist2[0].y;
		SNAP(0, 1, x, 0)
	}
	if (v.y < 1 - border)                                                                   /* :243-277 */
	{
		n[2] = add(v, mul(plist[3], std::abs(v.y - 1 + border)));
		n[3] = add(v, mul(plist[2], std::abs(v.y - 1 + border)));
		const V3 in = n[2];
		if (v.x < 0) { if (n[2].x < 0) n[2].x = 0; }
		else if (DOT(sub(n[3], v), sub(plist2[0], v)) > 0) n[2] = AIM(plist2[0], plist2[2].y - v.y, plist2[0].y - v.y);
		if (v.x > 1) { if (n[3].x > 1) n[3].x = 1; }
		else { const V3 delta = sub(plist2[1], v); if (DOT(sub(in, v), delta) > 0) n[3] = add(v, mul(delta, std::abs((plist2[3].y - v.y) / delta.y))); }
		n[2].y = n[3].y = plist2[2].y;
		SNAP(2, 3, x, 1)
	}
	if (v.x > 0)                                                                            /* :280-313 */
	{
		n[4] = add(v, mul(plist[0], std::abs(v.x)));
		n[5] = add(v, mul(plist[3], std::abs(v.x)));
		const V3 in = n[5];
		if (v.y > 1 - border) { if (n[5].y > 1 - border) n[5].y = 1 - border; }
		else if (DOT(sub(n[4], v), sub(plist2[2], v)) > 0) n[5] = AIM(plist2[2], plist2[3].x - v.x, plist2[2].x - v.x);
		if (v.y < border) { if (n[4].y < border) n[4].y = border; }
		else if (DOT(sub(in, v), sub(plist2[1], v)) > 0) n[4] = AIM(plist2[1], plist2[0].x - v.x, plist2[1].x - v.x);
		n[4].x = n[5].x = 0;
		SNAP(4, 5, y, 2)
	}
	if (v.x < 1)                                                                            /* :316-352 */
	{
		n[6] = add(v, mul(plist[1], std::abs(1 - v.x)));
		n[7] = add(v, mul(plist[2], std::abs(1 - v.x)));
		const V3 in = n[7];
		if (v.y > 1 - border) { if (n[7].y > 1 - border) n[7].y = 1 - border; }
		else if (DOT(sub(n[6], v), sub(plist2[3], v)) > 0) n[7] = AIM(plist2[3], plist2[2].x - v.x, plist2[3].x - v.x);
		if (v.y < border) { if (n[6].y < border) n[6].y = border; }
		else if (DOT(sub(in, v), sub(plist2[0], v)) > 0) n[6] = AIM(plist2[0], plist2[1].x - v.x, plist2[0].x - v.x);
		n[6].x = n[7].x = 1;
		SNAP(6, 7, y, 3)
	}
	#undef DOT
	#undef AIM
	#undef SNAP
	rm.p4 = p[4];
	rm.map_line_count = rm.res[0] + rm.res[1] + rm.res[2] + rm.res[3];
}

int orc_max_threads()
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

} /* extern "C" */
