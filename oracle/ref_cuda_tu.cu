/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * The reference's own CUDA kernel, unmodified, compiled for sm_100a: the same-GPU comparator of SURVEY.md §8d
 * ("the reference cudaRender recompiled as-is") that bench.py reports as roofline.reference_kernel_on_this_gpu.
 * Nothing here restates the algorithm: the kernel body is #included in place from /root/reference (read-only):
 *     RLE-Raycaster/src/RayMap.h, RLE-Raycaster/src/Cuda_Render.h   (struct Render, Render::render_line :96-737)
 *     RLE-Raycaster/inc/cutil_math.h
 *     RLE-Raycaster/src/Cuda_Main.cu:150-181   (__global__ cudaRender) — that file as a whole needs cutil and GL interop,
 *         so oracle/Makefile extracts exactly those lines into oracle/_ref/ref_cudaRender.inc (a build artefact, git-ignored)
 * with the compile-time configuration of R/src/core.h as shipped (BASELINE config 1: 1024 x 768 window, RENDER_SIZE 1024,
 * RAYS_CASTED 4096, 128 threads per block, 16300 bytes of shared memory).  The launch below is cuda_main_render2's
 * (R/src/Cuda_Main.cu:183-271) without the GL buffer mapping: Render block copied host -> device, grid (2, calls / 128).
 * Known defect kept as it is: the 31-word shared mask is too small for res_y = 1024, neighbouring threads share a
 * word for rows 992..1023, so a few texels there differ from run to run.
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
#include "RayMap.h"
#include "Cuda_Render.h"
#include "ref_cudaRender.inc"

extern "C" {

int refcuda_sizeof_render() { return (int)sizeof(Render); }
int refcuda_render_size() { return RENDER_SIZE; }
int refcuda_rays_casted() { return RAYS_CASTED; }

/* raymap: RayMap_GPU whose map4_gpu[] hold DEVICE pointers (R/src/main.cpp:277-278); d_out: device buffer of
 * RAYS_CASTED * RENDER_SIZE ints.  Runs the frame `repeats` times; ms_kernel = best cudaRender time (CUDA events),
 * ms_call = best time of the whole cuda_main_render2 sequence (82 KB Render block upload + kernel + synchronise). */
int refcuda_frame(const void* raymap, uint32_t* d_out, int repeats, float* ms_kernel, float* ms_call)
{
	static Render render;
	Render* render_gpu = 0;
	ushort* skipmap_gpu = 0;
	if (cudaMalloc((void**)&render_gpu, sizeof(Render)) != cudaSuccess) return -1;
	/* ofs_cache_start[0] = 0 (Cuda_Render.h:110-112) writes at byte 6 * x * res_y: give it room */
	if (cudaMalloc((void**)&skipmap_gpu, (size_t)RAYS_CASTED * RENDER_SIZE * 8) != cudaSuccess) return -1;
	cudaFuncSetAttribute(cudaRender, cudaFuncAttributeMaxDynamicSharedMemorySize, 16300);
	cudaEvent_t e0, e1, e2;
	cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
	float best_k = 1e30f, best_c = 1e30f;
	for (int it = 0; it < repeats; it++)
	{
		RayMap_GPU* rm = (RayMap_GPU*)raymap;
		int lines_to_raycast = rm->map_line_count;                         /* Cuda_Main.cu:194-196 */
		int thread_calls = ((rm->map_line_count / 2) | (THREAD_COUNT - 1)) + 1;
		if (lines_to_raycast > RAYS_CASTED) lines_to_raycast = RAYS_CASTED;
		dim3 threads(THREAD_COUNT, 1, 1);                                   /* Cuda_Main.cu:201-202 */
		dim3 grid(2, thread_calls / (threads.x), 1);
		cudaEventRecord(e0);
		render.set_target(RENDER_SIZE, RENDER_SIZE, (int*)d_out);            /* Cuda_Main.cu:204-205 */
		render.set_raymap(rm);
		cudaMemcpy(render_gpu, &render, sizeof(Render), cudaMemcpyHostToDevice);   /* Cuda_Main.cu:218 */
		cudaEventRecord(e1);
		cudaRender<<<grid, threads, 16300>>>(render_gpu, render.ray_map.map_line_count, render.ray_map.position,
		                                      render.ray_map.rotation, render.res_x, render.res_y, skipmap_gpu);   /* :226-235 */
		cudaEventRecord(e2);
		if (cudaDeviceSynchronize() != cudaSuccess) { printf("refcuda: %s\n", cudaGetErrorString(cudaGetLastError())); return -2; }
		float k = 0, c = 0;
		cudaEventElapsedTime(&k, e1, e2);
		cudaEventElapsedTime(&c, e0, e2);
		if (k < best_k) best_k = k;
		if (c < best_c) best_c = c;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
	cudaFree(render_gpu); cudaFree(skipmap_gpu);
	if (ms_kernel) *ms_kernel = best_k;
	if (ms_call) *ms_call = best_c;
	return 0;
}

} /* extern "C" */
