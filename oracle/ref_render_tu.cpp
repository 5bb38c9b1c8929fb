/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Glue that compiles the reference's own traversal kernel body as host C++
 * (SURVEY.md §8c "TU-A").  Nothing here restates the algorithm: the body is
 * #included in place from /root/reference (read-only, never copied):
 *     RLE-Raycaster/src/RayMap.h        (struct RayMap_GPU)
 *     RLE-Raycaster/src/Cuda_Render.h   (struct Render, Render::render_line :96-737)
 *     RLE-Raycaster/inc/cutil_math.h    (float3 operators + host fallbacks)
 * The only product of this file is oracle/_ref/libref_render[_bench].so which
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * arm may load.  The product path never links or loads it.
 *
 * Compile-time constants of R/src/core.h:3-10 are made run-time here by
 * re-pointing the macros at globals AFTER core.h has been read once (it is
 * #pragma once): SCREEN_SIZE_X feeds MIP_DISTANCE (Cuda_Render.h:182) and
 * RAYS_CASTED_RES; RAYS_DISTANCE feeds z_far (Cuda_Render.h:180).
 */
#define IN_CUDA_ENV
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <cuda_runtime.h>
#include "cutil_math.h"
/* keep core.h:64-70 from macro-defining uint/ushort (cutil_math.h typedefs them) */
#define uint uint
#define ushort ushort
#include "Core.h"

/* -DPERPIXELFORWARD (R/src/core.h:31): the per-pixel skip list REPLACES the shared-memory occlusion mask.  With both
 * on, the pixel loop's `continue` at the mask test skips the loop's own ++y (Cuda_Render.h:682-731) and the loop
 * never ends, so that build is SHAREMEMCLIP off (core.h:32 defines it unconditionally). */
#ifdef PERPIXELFORWARD
#undef SHAREMEMCLIP
#endif

static int g_ref_screen_size_x = 1024;
static int g_ref_rays_distance = 80000;
#undef SCREEN_SIZE_X
#define SCREEN_SIZE_X g_ref_screen_size_x
#undef RAYS_DISTANCE
#define RAYS_DISTANCE g_ref_rays_distance
#undef RAYS_CASTED
#ifdef DETAIL_BENCH
#define RAYS_CASTED 32768 /* struct Render::perf[] capacity (Cuda_Render.h:22) */
#else
#define RAYS_CASTED 4
#endif

#include "RayMap.h"
#include "Cuda_Render.h"

#ifdef _OPENMP
#include <omp.h>
#endif

static Render g_render;

extern "C" {

int ref_sizeof_raymap() { return (int)sizeof(RayMap_GPU); }
int ref_sizeof_map4() { return (int)sizeof(Map4); }
int ref_has_detail_bench()
{
#ifdef DETAIL_BENCH
	return 1;
#else
	return 0;
#endif
}
int ref_max_threads()
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* Render every ray plane of one frame into `warp` (uint32[rays][res_y]) exactly as
 * cudaRender does (R/src/Cuda_Main.cu:150-181): one render_line call per ray index,
 * viewpos/viewrot taken from the ray map (Cuda_Main.cu:229-230).
 * Differences from the device launch, all required to make the body well-defined
 * on a host (SURVEY.md §5 "Race detection"):
 *   - every ray gets a private occlusion mask of >= res_y bits, zeroed in full
 *     (the reference zeroes 31 words only, Cuda_Render.h:263);
 *   - the dead store ofs_cache_start[0]=0 (Cuda_Render.h:112) is pointed at a
 *     per-thread dummy word.
 * ray_begin/ray_end select a slice of the ray index range. threads<=0: OpenMP default.
 * perf_out (may be NULL; needs the DETAIL_BENCH build): int[5] sums of
 * elems_total, elems_processed, voxels_processed, elems_rendered, pixels. */
int ref_render_frame(const void* raymap_gpu, int res_x, int res_y, int mip_distance, int z_far,
                     uint32_t* warp, int ray_begin, int ray_end, int threads, long long* perf_out)
{
	g_ref_screen_size_x = mip_distance;
	g_ref_rays_distance = z_far;
	g_render.set_raymap((RayMap_GPU*)raymap_gpu);
	g_render.res_x = res_x;
	g_render.res_y = res_y;
	g_render.data_rgb = (int*)warp;
	int count = g_render.ray_map.map_line_count;
	if (ray_end > count) ray_end = count;
	if (ray_begin < 0) ray_begin = 0;
#ifdef DETAIL_BENCH
	if (count > RAYS_CASTED) return -2;
	memset(g_render.perf, 0, sizeof(g_render.perf));
#endif
	/* render_line itself clears words 0..30 (Cuda_Render.h:263), whatever res_y is */
	const int mask_words = ((res_y + 31) / 32 > 31 ? (res_y + 31) / 32 : 31) + 2;
	vec3f pos = g_render.ray_map.position;
	vec3f rot = g_render.ray_map.rotation;
#ifdef _OPENMP
	if (threads > 0) omp_set_num_threads(threads);
#endif
	#pragma omp parallel
	{
		unsigned int* mask = (unsigned int*)malloc(mask_words * sizeof(unsigned int));
		uint dummy[4];
#ifdef PERPIXELFORWARD
		/* that build really uses the skip row (ushort[res_y], cleared by render_line over the clip range) and still
		 * writes ((uint*)skip)[x*res_y] = 0: one private buffer per thread that holds both */
		uint* skipbuf = (uint*)calloc((size_t)(ray_end > 0 ? ray_end : 1) * res_y + res_y + 4, sizeof(uint));
#endif
		#pragma omp for schedule(dynamic, 16)
		for (int x = ray_begin; x < ray_end; x++)
		{
			memset(mask, 0, mask_words * sizeof(unsigned int));
#ifdef PERPIXELFORWARD
			memset(skipbuf, 0, (size_t)res_y * sizeof(ushort));
			ushort* skip = (ushort*)skipbuf;
#else
			/* render_line writes ((uint*)skip)[x*res_y] = 0 */
			ushort* skip = (ushort*)(dummy - (long long)x * res_y);
#endif
			g_render.render_line(x, mask, pos, rot, res_x, res_y, skip);
		}
#ifdef PERPIXELFORWARD
		free(skipbuf);
#endif
		free(mask);
	}
#ifdef DETAIL_BENCH
	if (perf_out)
	{
		for (int k = 0; k < 5; k++) perf_out[k] = 0;
		for (int x = ray_begin; x < ray_end; x++)
		{
			perf_out[0] += g_render.perf[x].elems_total;
			perf_out[1] += g_render.perf[x].elems_processed;
			perf_out[2] += g_render.perf[x].voxels_processed;
			perf_out[3] += g_render.perf[x].elems_rendered;
			perf_out[4] += g_render.perf[x].pixels;
		}
	}
#else
	(void)perf_out;
#endif
	return 0;
}

} /* extern "C" */
