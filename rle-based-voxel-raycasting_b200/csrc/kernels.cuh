// Device-side parameter blocks shared between the launcher (capi.cu) and kernels.cu.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace rlerc {

struct LevelDev {
	const uint2*    map;    // [sz][sx] {slab offset (ushort index), n_runs | first_run<<16}
	const uint16_t* slabs;
	int sx, sz;
};

// The LOD / z schedule of a frame (Cuda_Render.h:343-367): z counts crossings in steps of dz, dz doubles whenever
// z passes mapswitch (which doubles too), the ray plane ends at z_far.  All integer and the same for every ray plane
// of a frame, so the host tabulates its phases (constant dz) once: crossing k of any ray plane has a known z, dz, mip.
struct LodSched {
	int nphase;              // phases until z_far (<= 32)
	int k_total;             // crossings until z_far
	int ph_k[33];            // first crossing of phase p; ph_k[nphase] = k_total
	int ph_z[32];            // z before that crossing
	int ph_nsw[32];          // LOD switches made before it: dz = 1 << nsw, mapswitch = mapswitch0 << nsw, mip = min(nsw, nummaps-1)
};

// Everything the traversal kernel needs for one frame (~1 KB), passed as a
// __grid_constant__ kernel parameter: replaces the 82,832-byte `Render` block the reference
// copies host->device every frame (R/src/Cuda_Main.cu:218).
struct TraverseParams {
	LevelDev level[16];
	int nummaps;
	int res[4];              // rays per quadrant           (RayMap_GPU::res)
	float vp[3];             // p_2d[5], vanishing point    (Cuda_Render.h:149)
	float p_no[8][3];        // quadrant border points      (Cuda_Render.h:150)
	float clip_min, clip_max;
	float to3d[4][4];
	float p4[3];
	float viewpos[3];
	// trig of the camera angles, evaluated once on the host with the libm the oracle uses
	float sin_x, cos_x, sin_y, cos_y;   // sin/cos(rotation.x), sin/cos(rotation.y)  (Cuda_Render.h:188-191)
	float sin_my, cos_my;               // sin/cos(-rotation.y)                      (Cuda_Render.h:77-78)
	int res_x, res_y;
	int mapswitch0;          // int(MIP_DISTANCE * (0.25*(4-abs(rot.x)))), double math (Cuda_Render.h:335)
	int z_far;
	int ray_begin, ray_end;  // slice of the ray index range rendered by this launch
	// interleaved multi-GPU slices: ray r belongs to this launch iff (r / slice_block) % slice_n == slice_rank
	int slice_block, slice_n, slice_rank;
	int mask_words;          // per-ray occlusion bitmask size in 32-bit words
	int flags;               // RLERC_FLAG_CLIPREGION | RLERC_FLAG_HEIGHT_COLOR (include/rlerc.h; k_traverse_f only)
	uint32_t* warp;          // [rays_casted][res_y]
	uint32_t* ids;           // optional [rays_casted][res_y][2]
	unsigned long long* counters; // optional [10]
	// decoupled DDA producer (traverse_warp.cu); dda_ring == nullptr: every warp runs its own DDA
	float4* dda_ring;        // [rays][RING_DEPTH][33]
	int* dda_head;           // [rays] batches produced
	int* dda_tail;           // [rays] batches consumed, -1 = ray plane finished
	int* dda_err;            // protocol time-out flag
	int dda_mode;            // 0: serial DDA in every warp, otherwise merge-path DDA (ignored when dda_ring is set)
	// production kernels (traverse_filter.cu): DDA state of every ray plane of the launch at every 32nd crossing
	LodSched lod;
	float2* dda_states;      // [launch rays][ceil(k_total / 32)][3], written by k_dda_states
	// k_dda_states runs CONCURRENTLY with the traversal kernel (programmatic dependent launch): per ray plane of the launch,
	// (epoch << 32) | number of batches whose states are published
	unsigned long long* dda_progress;
	unsigned int dda_epoch;  // this launch's epoch: anything else in dda_progress is left over from an earlier launch
};

struct UnwarpParams {
	const uint32_t* warp;
	uint8_t* rgba;           // [H][W][4], row 0 = top
	int W, H;                // window
	int RS, RC;              // warped buffer: RS texels wide (along a ray), RC rows (rays)
	float vanish_x, vanish_y;
	float border;            // (RESX - RESY) / (RESX * 2), frag:22
	float ofs_add[4];
	float ratio;             // RAYS_CASTED_RES / RAYS_CASTED
	int rot_x_gt0;
	int row_begin, row_end;
	int ray_begin, ray_end;  // slice mode (ray_end < 0: off)
	int slice_block, slice_n, slice_rank;   // interleaved slice mode (slice_n <= 1: off)
	// multi-GPU pull mode (csrc/group.cu): peer_n == slice_n > 1, warp_peer[g] = GPU g's warped buffer (peer mapping);
	// every pixel of the rows is produced, its texel loaded from the GPU that traversed its ray plane
	int peer_n;
	const uint32_t* warp_peer[8];
	const uint32_t* shade_rgb;   // [65536] colour of every attribute value (k_shade_tables)
	const uint8_t* shade_alpha;  // [65536] smoothing weight of every depth value
	int shader;              // 0: colorize_buddha_soft.frag, 1: colorize_buddha_soft_2xAA.frag (RLERC_FLAG_SHADER_2XAA)
	int generic;             // evaluate the shader's texel arithmetic statement by statement (vanishing point beyond 1e6, kernels.cu)
};

struct SoftParams {
	const uint8_t* in;       // RGBA8 [H][W][4], row 0 = top, A = quantised 0.001/z (k_unwarp output)
	uint8_t* out;            // RGBA8 [H][W][4]
	int W, H;
	int fbo;                 // edge of the square FBO texture the reference renders pass 1 into (2048)
	float tap_x[7], tap_y[7];   // depth probe offsets sin(a)*0.005, cos(a)*0.005 (soft.frag:16-18)
};

void launch_soft(const SoftParams& p, cudaStream_t st);
void launch_traverse(const TraverseParams& p, int lanes_per_ray, bool ids, cudaStream_t st);
void launch_traverse_warp(const TraverseParams& p, bool ids, cudaStream_t st);   // traverse_warp.cu
void launch_traverse_filter(const TraverseParams& p, bool ids, cudaStream_t st); // traverse_filter.cu
void launch_traverse_pair(const TraverseParams& p, cudaStream_t st);              // traverse_filter.cu (k_traverse_p, variant 68)
void launch_traverse_quad(const TraverseParams& p, cudaStream_t st);              // traverse_filter.cu (k_traverse_q, variant 69)
void launch_dda_states(const TraverseParams& p, cudaStream_t st);                 // traverse_filter.cu (k_dda_states, before either of the two)
size_t traverse_dda_state_bytes(const TraverseParams& p);
size_t traverse_dda_progress_bytes(const TraverseParams& p);
void launch_traverse_chunk(const TraverseParams& p, bool ids, int wpb, cudaStream_t st); // traverse_chunk.cu
size_t traverse_ring_bytes(int rays);
void launch_unwarp(const UnwarpParams& p, cudaStream_t st, bool texels = false);
void launch_shade_tables(uint32_t* rgb, uint8_t* alpha, cudaStream_t st);
void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t st);
void launch_read_u4(const void* p, size_t bytes, int iters, uint32_t* sink, cudaStream_t st);

} // namespace rlerc
