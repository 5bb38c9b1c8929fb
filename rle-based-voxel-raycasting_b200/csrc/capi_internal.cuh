// Internals of the C ABI shared by capi.cu and group.cu: the context, its frame slots, the launch helpers.
#pragma once
#include <stdint.h>
#include <vector>
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "rlerc_internal.h"

#define CK(call)                                                                        \
	do {                                                                                \
		cudaError_t e_ = (call);                                                        \
		if (e_ != cudaSuccess) {                                                        \
			set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
			return RLERC_ERR_CUDA;                                                      \
		}                                                                               \
	} while (0)

struct FrameSlot {            // one in-flight frame of the pipelined path
	uint32_t* d_warp = nullptr;
	uint8_t* d_rgba = nullptr;
	float2* d_states = nullptr;      // DDA states of the frame's ray planes (k_dda_states -> k_traverse_f / _p)
	size_t warp_bytes = 0, rgba_bytes = 0, states_bytes = 0;
	cudaEvent_t done = nullptr;
	cudaStream_t stream = nullptr;   // each in-flight frame renders on its own stream: the tail of one
	                                 // frame's traversal (a few long ray planes) overlaps the next frame
	uint8_t* host_dst = nullptr;
	bool busy = false;
};

struct rlerc_streamed;             // stream.cu: a streamed replica (virtual address ranges + resident chunks)

struct rlerc_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;      // traversal + unwarp
	cudaStream_t copy_stream = nullptr; // D2H of finished frames
	// device scene replica
	int nummaps = 0;
	rlerc::LevelDev level[RLERC_MAX_MAPS];
	int level_sy[RLERC_MAX_MAPS];
	uint64_t level_slabs[RLERC_MAX_MAPS];
	std::vector<void*> scene_allocs;    // empty when the replica is borrowed (rlerc_scene_share) or streamed
	rlerc_streamed* stream_state = nullptr;   // rlerc_scene_upload_streamed
	// frame resources
	uint32_t* d_warp = nullptr;
	size_t warp_bytes = 0;
	uint8_t* d_rgba = nullptr;
	size_t rgba_bytes = 0;
	float2* d_states = nullptr;         // DDA states of the ray planes of the frame on c->stream
	size_t states_bytes = 0;
	float2** cur_states = nullptr;      // the states buffer render_impl uses (rlerc_frame_submit points it at the slot's)
	size_t* cur_states_bytes = nullptr;
	unsigned int dda_epoch = 0;         // epoch of the last traversal launch (k_dda_states -> traversal hand-over)
	const char* last_kernel = "";       // traversal kernel of the last launch
	uint32_t* d_ids_scratch = nullptr;
	unsigned long long* d_counters = nullptr;
	uint32_t* d_shade_rgb = nullptr;    // shading tables of k_unwarp (kernels.cu k_shade_tables), 320 KB
	uint8_t* d_shade_alpha = nullptr;
	int lanes = 0;                      // 0 = auto (pick_lanes)
	int sm_count = 148;
	int dda_mode = 0;                   // 0 serial (default: fastest measured), 2 merge path (k_traverse_w)
	int producer = 0;                   // decoupled DDA producer blocks (k_traverse_w): optional, off by default
	float4* d_ring = nullptr;
	size_t ring_bytes = 0;
	int* d_ring_ctl = nullptr;          // head[rays] | tail[rays] | err
	size_t ring_ctl_bytes = 0;
	bool timing = false;
	bool pipelined = false;             // several frames in flight (rlerc_frame_submit, groups): throughput-bound, never the paired kernel
	bool own_stream = true;
	cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
	bool ev_valid[2] = { false, false };
	// pipeline
	static const int kSlots = 8;        // frames in flight in rlerc_frame_submit (each on its own stream)
	FrameSlot slot[kSlots];
	int next_ticket = 0;
};


namespace rlerc {

int set_dev(rlerc_ctx* c);
int ensure(void** p, size_t* have, size_t need, cudaStream_t st, bool zero = true);
int check_cfg(const rlerc_frame_config* cfg);
int fill_traverse(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, int ray_begin, int ray_end, uint32_t* d_warp, TraverseParams& P);
void stream_free(rlerc_ctx* c);    // stream.cu
int fill_unwarp(const rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, const uint32_t* d_warp, uint8_t* d_rgba, UnwarpParams& U);
int render_impl(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                int ray_begin, int ray_end, uint32_t* d_warp, uint32_t* d_ids, bool ids,
                int slice_block = 1, int slice_n = 1, int slice_rank = 0, uint32_t* prof_out = nullptr);
int unwarp_impl(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, const uint32_t* d_warp,
                uint8_t* d_rgba, int row_begin, int row_end, int ray_begin, int ray_end,
                int slice_block = 1, int slice_n = 1, int slice_rank = 0);

} // namespace rlerc
