// Multi-GPU frames (include/rlerc.h "multi-GPU"; nothing in the reference corresponds: it is single-GPU).
//
// The ray planes of a frame are independent (R/src/Cuda_Render.h:109: a ray plane writes only its own row of the warped
// buffer), so N GPUs, each holding a full replica of the compressed scene, traverse interleaved blocks of ray planes.
// What remains is compositing, and that is done here WITHOUT a collective moving pixel data:
//
//   member r:  k_dda_states + traversal of its ray planes -> its own warped buffer (rows of the other members stay untouched)
//              barrier A            every member's rows of this frame are written
//              k_unwarp, rows [r H/N, (r+1) H/N) of the window: the texel of a pixel is loaded from the warped buffer of
//                                   the GPU that traversed its ray plane (peer mapping: NVLink loads, 32-byte sectors), the
//                                   shaded pixel is stored either into member `dst`'s frame image (peer mapping: 128-bit
//                                   NVLink stores; the finished frame is on one GPU) or into the member's own image (each
//                                   GPU then copies its band to the host over its own PCIe link: no rank-0 funnel)
//              barrier B            every band is written / every member is done reading: the slot may be reused
//
// Both the per-pixel work and the bytes that cross NVLink are 1/N per GPU (a reduce of whole RGBA images, the round-1
// compositor, costs every GPU a full-window unwarp and a full-image reduction).
//
// Barriers are flags in peer-mapped device memory: a one-warp kernel stores the slot's generation number into every
// member's flag array (st.release.sys) and spins until all members' numbers have arrived in its own (ld.acquire.sys,
// __nanosleep between polls, a trap after ~4 s instead of a hang).  Waiting occupies one warp, never an SM's worth of
// blocks, so a member that waits cannot starve the kernels its peers are waiting for; every wait is for a kernel that
// was launched EARLIER in the common frame order, hence no cycle.  `depth` frames are in flight, each slot with its own
// stream, warped buffer, image and DDA states.
//
// Members may live in one process (rlerc_create_multi: peer access, plain pointers) or one per process (one process per
// GPU under torchrun: CUDA IPC handles, exchanged by the caller through rlerc_group_export / rlerc_group_connect).
#include <stdio.h>
#include <string.h>
#include <unistd.h>
#include <vector>
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "rlerc_internal.h"
#include "capi_internal.cuh"

using namespace rlerc;

#define RLERC_GROUP_MAX 8
#define RLERC_GROUP_MAGIC 0x726c6772u

namespace {

struct GroupBlob {                      // what a member tells the others (RLERC_GROUP_BLOB_BYTES)
	uint32_t magic;
	int32_t rank, pid, device, depth;
	uint32_t pad0;
	uint64_t warp_ptr, img_ptr, flags_ptr;          // addresses in the owner's process
	uint64_t warp_slot_bytes, img_slot_bytes;
	cudaIpcMemHandle_t warp_h, img_h, flags_h;
	uint64_t views_ptr;                              // 0: the member takes no view batches (rlerc_group_enable_views)
	cudaIpcMemHandle_t views_h;
};
static_assert(sizeof(GroupBlob) <= RLERC_GROUP_BLOB_BYTES, "blob size");

struct GroupSlot {
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	cudaEvent_t t0 = nullptr, t1 = nullptr;          // frame timing (rlerc_group_last_ms)
	float2* d_states = nullptr;
	size_t states_bytes = 0;
	bool busy = false;
	bool timed = false;
};

struct BarrierParams {
	uint32_t* peer[RLERC_GROUP_MAX];    // every member's flag array (own included)
	const uint32_t* local;
	int n, rank, idx;                   // idx: first word of this (slot, kind) in a flag array
	uint32_t gen;
};

__global__ void __launch_bounds__(32) k_group_barrier(const BarrierParams B)
{
	const int p = threadIdx.x;
	if (p >= B.n) return;
	// everything this GPU did before (stream order) is visible system-wide before the flag is
	__threadfence_system();
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(B.peer[p] + B.idx + B.rank), "r"(B.gen) : "memory");
	const uint32_t* f = B.local + B.idx + p;
	for (long long spins = 0;; spins++)
	{
		uint32_t v;
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
		if ((int32_t)(v - B.gen) >= 0) break;
		__nanosleep(spins < 64 ? 100 : 1000);
		if (spins > (4ll << 20)) __trap();          // ~4 s: a member died; fail loudly (launch error), never hang
	}
}

} // namespace

struct rlerc_group {
	rlerc_ctx* c = nullptr;
	int rank = 0, n = 1, depth = 1, block = 32;
	rlerc_frame_config cfg;
	uint32_t* warp = nullptr;           // [depth][rays_casted][render_size]
	uint8_t* img = nullptr;             // [depth][height][width][4]
	uint32_t* flags = nullptr;          // [depth][2][RLERC_GROUP_MAX]
	uint8_t* views = nullptr;           // [depth][nranks][height][width][4]: where the members' finished views arrive (rlerc_group_enable_views)
	uint8_t* views_peer[RLERC_GROUP_MAX];
	size_t warp_slot_bytes = 0, img_slot_bytes = 0;
	const uint32_t* warp_peer[RLERC_GROUP_MAX];
	uint8_t* img_peer[RLERC_GROUP_MAX];
	uint32_t* flags_peer[RLERC_GROUP_MAX];
	void* opened[RLERC_GROUP_MAX][4];   // IPC mappings to close
	bool connected = false;
	std::vector<GroupSlot> slot;
	int next_ticket = 0;
};

#define CKG(call)                                                                       \
	do {                                                                                \
		cudaError_t e_ = (call);                                                        \
		if (e_ != cudaSuccess) {                                                        \
			set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
			return RLERC_ERR_CUDA;                                                      \
		}                                                                               \
	} while (0)

static void band_of(const rlerc_group* g, int r, int* row_begin, int* row_end)
{
	const int rows = (g->cfg.height + g->n - 1) / g->n;
	int b = r * rows, e = b + rows;
	if (b > g->cfg.height) b = g->cfg.height;
	if (e > g->cfg.height) e = g->cfg.height;
	*row_begin = b; *row_end = e;
}

extern "C" {

int rlerc_group_create(rlerc_ctx* c, int rank, int nranks, int depth, int slice_block, const rlerc_frame_config* cfg, rlerc_group** out)
{
	if (!c || !out) { set_error("rlerc_group_create: null argument"); return RLERC_ERR_ARG; }
	*out = nullptr;
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if (nranks < 1 || nranks > RLERC_GROUP_MAX || rank < 0 || rank >= nranks || depth < 1 || depth > 64 || slice_block < 1)
	{
		set_error("rlerc_group_create: bad group shape (rank %d of %d, depth %d, block %d; at most %d members)", rank, nranks, depth, slice_block, RLERC_GROUP_MAX);
		return RLERC_ERR_ARG;
	}
	if ((rc = set_dev(c))) return rc;
	rlerc_group* g = new rlerc_group();
	g->c = c; g->rank = rank; g->n = nranks; g->depth = depth; g->block = slice_block; g->cfg = *cfg;
	memset(g->warp_peer, 0, sizeof(g->warp_peer)); memset(g->img_peer, 0, sizeof(g->img_peer));
	memset(g->flags_peer, 0, sizeof(g->flags_peer)); memset(g->opened, 0, sizeof(g->opened));
	memset(g->views_peer, 0, sizeof(g->views_peer));
	g->warp_slot_bytes = ((size_t)cfg->rays_casted * cfg->render_size * 4 + 255) & ~(size_t)255;
	g->img_slot_bytes = ((size_t)cfg->width * cfg->height * 4 + 255) & ~(size_t)255;
	const size_t flag_bytes = (size_t)depth * 2 * RLERC_GROUP_MAX * sizeof(uint32_t);
	cudaError_t e;
	if ((e = cudaMalloc((void**)&g->warp, g->warp_slot_bytes * depth)) != cudaSuccess ||
	    (e = cudaMalloc((void**)&g->img, g->img_slot_bytes * depth)) != cudaSuccess ||
	    (e = cudaMalloc((void**)&g->flags, flag_bytes)) != cudaSuccess)
	{
		set_error("rlerc_group_create: cudaMalloc failed: %s", cudaGetErrorString(e));
		rlerc_group_destroy(g);
		return e == cudaErrorMemoryAllocation ? RLERC_ERR_NOMEM : RLERC_ERR_CUDA;
	}
	CKG(cudaMemset(g->warp, 0, g->warp_slot_bytes * depth));
	CKG(cudaMemset(g->img, 0, g->img_slot_bytes * depth));
	CKG(cudaMemset(g->flags, 0, flag_bytes));
	CKG(cudaDeviceSynchronize());
	g->slot.resize(depth);
	for (int k = 0; k < depth; k++)
	{
		CKG(cudaStreamCreateWithFlags(&g->slot[k].stream, cudaStreamNonBlocking));
		CKG(cudaEventCreateWithFlags(&g->slot[k].done, cudaEventDisableTiming));
		CKG(cudaEventCreate(&g->slot[k].t0));
		CKG(cudaEventCreate(&g->slot[k].t1));
	}
	// a group of one needs no connection
	g->warp_peer[rank] = g->warp; g->img_peer[rank] = g->img; g->flags_peer[rank] = g->flags;
	g->connected = nranks == 1;
	*out = g;
	return RLERC_OK;
}

void rlerc_group_destroy(rlerc_group* g)
{
	if (!g) return;
	if (g->c) { cudaSetDevice(g->c->device); cudaDeviceSynchronize(); }
	for (int p = 0; p < RLERC_GROUP_MAX; p++)
		for (int k = 0; k < 4; k++)
			if (g->opened[p][k]) cudaIpcCloseMemHandle(g->opened[p][k]);
	for (auto& s : g->slot)
	{
		if (s.d_states) cudaFree(s.d_states);
		if (s.done) cudaEventDestroy(s.done);
		if (s.t0) cudaEventDestroy(s.t0);
		if (s.t1) cudaEventDestroy(s.t1);
		if (s.stream) cudaStreamDestroy(s.stream);
	}
	if (g->warp) cudaFree(g->warp);
	if (g->img) cudaFree(g->img);
	if (g->flags) cudaFree(g->flags);
	if (g->views) cudaFree(g->views);
	delete g;
}

int rlerc_group_export(rlerc_group* g, void* blob)
{
	if (!g || !blob) { set_error("rlerc_group_export: null argument"); return RLERC_ERR_ARG; }
	int rc = set_dev(g->c);
	if (rc) return rc;
	GroupBlob b;
	memset(&b, 0, sizeof(b));
	b.magic = RLERC_GROUP_MAGIC; b.rank = g->rank; b.pid = (int32_t)getpid(); b.device = g->c->device; b.depth = g->depth;
	b.warp_ptr = (uint64_t)g->warp; b.img_ptr = (uint64_t)g->img; b.flags_ptr = (uint64_t)g->flags;
	b.warp_slot_bytes = g->warp_slot_bytes; b.img_slot_bytes = g->img_slot_bytes;
	CKG(cudaIpcGetMemHandle(&b.warp_h, g->warp));
	CKG(cudaIpcGetMemHandle(&b.img_h, g->img));
	CKG(cudaIpcGetMemHandle(&b.flags_h, g->flags));
	if (g->views) { b.views_ptr = (uint64_t)g->views; CKG(cudaIpcGetMemHandle(&b.views_h, g->views)); }
	memset(blob, 0, RLERC_GROUP_BLOB_BYTES);
	memcpy(blob, &b, sizeof(b));
	return RLERC_OK;
}

int rlerc_group_connect(rlerc_group* g, const void* blobs)
{
	if (!g || !blobs) { set_error("rlerc_group_connect: null argument"); return RLERC_ERR_ARG; }
	int rc = set_dev(g->c);
	if (rc) return rc;
	for (int p = 0; p < g->n; p++)
	{
		if (p == g->rank) continue;
		GroupBlob b;
		memcpy(&b, (const char*)blobs + (size_t)p * RLERC_GROUP_BLOB_BYTES, sizeof(b));
		if (b.magic != RLERC_GROUP_MAGIC || b.rank != p || b.depth != g->depth ||
		    b.warp_slot_bytes != g->warp_slot_bytes || b.img_slot_bytes != g->img_slot_bytes)
		{
			set_error("rlerc_group_connect: member %d's description does not match this group (rank %d, depth %d)", p, b.rank, b.depth);
			return RLERC_ERR_ARG;
		}
		if (b.pid == (int32_t)getpid())
		{
			// same process: plain pointers + peer access
			if (b.device != g->c->device)
			{
				int can = 0;
				CKG(cudaDeviceCanAccessPeer(&can, g->c->device, b.device));
				if (!can) { set_error("rlerc_group_connect: device %d cannot access device %d (no peer path)", g->c->device, b.device); return RLERC_ERR_CUDA; }
				cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(e)); return RLERC_ERR_CUDA; }
				cudaGetLastError();
			}
			g->warp_peer[p] = (const uint32_t*)b.warp_ptr; g->img_peer[p] = (uint8_t*)b.img_ptr; g->flags_peer[p] = (uint32_t*)b.flags_ptr;
			g->views_peer[p] = (uint8_t*)b.views_ptr;
		}
		else
		{
			void *w = nullptr, *i = nullptr, *f = nullptr;
			CKG(cudaIpcOpenMemHandle(&w, b.warp_h, cudaIpcMemLazyEnablePeerAccess));
			g->opened[p][0] = w;
			CKG(cudaIpcOpenMemHandle(&i, b.img_h, cudaIpcMemLazyEnablePeerAccess));
			g->opened[p][1] = i;
			CKG(cudaIpcOpenMemHandle(&f, b.flags_h, cudaIpcMemLazyEnablePeerAccess));
			g->opened[p][2] = f;
			g->warp_peer[p] = (const uint32_t*)w; g->img_peer[p] = (uint8_t*)i; g->flags_peer[p] = (uint32_t*)f;
			if (b.views_ptr)
			{
				void* v = nullptr;
				CKG(cudaIpcOpenMemHandle(&v, b.views_h, cudaIpcMemLazyEnablePeerAccess));
				g->opened[p][3] = v;
				g->views_peer[p] = (uint8_t*)v;
			}
		}
	}
	g->connected = true;
	return RLERC_OK;
}

static int group_barrier(rlerc_group* g, int slot, int kind, uint32_t gen, cudaStream_t st)
{
	BarrierParams B;
	memset(&B, 0, sizeof(B));
	for (int p = 0; p < g->n; p++) B.peer[p] = g->flags_peer[p];
	B.local = g->flags; B.n = g->n; B.rank = g->rank;
	B.idx = (slot * 2 + kind) * RLERC_GROUP_MAX;
	B.gen = gen;
	k_group_barrier<<<1, 32, 0, st>>>(B);
	return RLERC_OK;
}

int rlerc_group_submit(rlerc_group* g, const rlerc_raymap* rm, int dst_rank, uint8_t* host_rgba)
{
	if (!g || !rm) { set_error("rlerc_group_submit: null argument"); return RLERC_ERR_ARG; }
	if (!g->connected) { set_error("rlerc_group_submit: the group is not connected (rlerc_group_connect)"); return RLERC_ERR_STATE; }
	if (dst_rank >= g->n) { set_error("rlerc_group_submit: destination %d is not a member", dst_rank); return RLERC_ERR_ARG; }
	rlerc_ctx* c = g->c;
	int rc = set_dev(c);
	if (rc) return rc;
	const int ticket = g->next_ticket;
	if (dst_rank == RLERC_GROUP_DST_ROTATE) dst_rank = ticket % g->n;      // frame t is assembled on member t mod n
	const int k = ticket % g->depth;
	const uint32_t gen = (uint32_t)(ticket / g->depth) + 1u;
	GroupSlot& S = g->slot[k];
	if (S.busy) { CKG(cudaEventSynchronize(S.done)); S.busy = false; }
	uint32_t* warp = (uint32_t*)((char*)g->warp + (size_t)k * g->warp_slot_bytes);
	uint8_t* img = g->img + (size_t)k * g->img_slot_bytes;
	// the context's launch helpers run on the slot's stream with the slot's DDA states
	cudaStream_t const main_stream = c->stream;
	const bool was_pipelined = c->pipelined;
	c->stream = S.stream;
	c->cur_states = &S.d_states; c->cur_states_bytes = &S.states_bytes;
	c->pipelined = g->depth > 1;
	if (c->timing) { cudaEventRecord(S.t0, S.stream); }
	rc = render_impl(c, rm, &g->cfg, 0, -1, warp, nullptr, false, g->block, g->n, g->rank);
	if (!rc && g->n > 1) rc = group_barrier(g, k, 0, gen, S.stream);
	int rb = 0, re = g->cfg.height;
	if (!rc)
	{
		UnwarpParams U;
		fill_unwarp(c, rm, &g->cfg, warp, img, U);
		if (g->n > 1)
		{
			band_of(g, g->rank, &rb, &re);
			U.row_begin = rb; U.row_end = re;
			U.slice_block = g->block; U.slice_n = g->n; U.slice_rank = g->rank;
			U.peer_n = g->n;
			for (int p = 0; p < g->n; p++) U.warp_peer[p] = (const uint32_t*)((const char*)g->warp_peer[p] + (size_t)k * g->warp_slot_bytes);
			if (dst_rank >= 0) U.rgba = g->img_peer[dst_rank] + (size_t)k * g->img_slot_bytes;
		}
		if (c->timing) cudaEventRecord(c->ev[2], S.stream);
		launch_unwarp(U, S.stream);
		if (c->timing) { cudaEventRecord(c->ev[3], S.stream); c->ev_valid[1] = true; }
		if (g->n > 1) rc = group_barrier(g, k, 1, gen, S.stream);
	}
	if (c->timing) { cudaEventRecord(S.t1, S.stream); S.timed = true; }
	c->stream = main_stream;
	c->cur_states = nullptr; c->cur_states_bytes = nullptr;
	c->pipelined = was_pipelined;
	if (rc) return rc;
	CKG(cudaGetLastError());
	if (host_rgba)
	{
		const size_t rowb = (size_t)g->cfg.width * 4;
		if (g->n == 1 || dst_rank == g->rank)
			CKG(cudaMemcpyAsync(host_rgba, img, rowb * g->cfg.height, cudaMemcpyDeviceToHost, S.stream));
		else if (dst_rank < 0 && re > rb)
			CKG(cudaMemcpyAsync(host_rgba + rowb * rb, img + rowb * rb, rowb * (re - rb), cudaMemcpyDeviceToHost, S.stream));
	}
	CKG(cudaEventRecord(S.done, S.stream));
	S.busy = true;
	g->next_ticket++;
	return ticket;
}

int rlerc_group_enable_views(rlerc_group* g)
{
	if (!g) return RLERC_ERR_ARG;
	if (g->connected && g->n > 1) { set_error("rlerc_group_enable_views: call it before rlerc_group_export / rlerc_group_connect"); return RLERC_ERR_STATE; }
	if (g->views) return RLERC_OK;
	int rc = set_dev(g->c);
	if (rc) return rc;
	const size_t bytes = g->img_slot_bytes * g->n * g->depth;
	cudaError_t e = cudaMalloc((void**)&g->views, bytes);
	if (e != cudaSuccess) { set_error("rlerc_group_enable_views: cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); return e == cudaErrorMemoryAllocation ? RLERC_ERR_NOMEM : RLERC_ERR_CUDA; }
	CKG(cudaMemset(g->views, 0, bytes));
	g->views_peer[g->rank] = g->views;
	return RLERC_OK;
}

int rlerc_group_submit_view(rlerc_group* g, const rlerc_raymap* rm, int dst_rank, uint8_t* host_rgba)
{
	if (!g || !rm) { set_error("rlerc_group_submit_view: null argument"); return RLERC_ERR_ARG; }
	if (!g->connected) { set_error("rlerc_group_submit_view: the group is not connected"); return RLERC_ERR_STATE; }
	if (dst_rank >= g->n) { set_error("rlerc_group_submit_view: destination %d is not a member", dst_rank); return RLERC_ERR_ARG; }
	if (dst_rank >= 0 && !g->views_peer[dst_rank]) { set_error("rlerc_group_submit_view: member %d takes no views (rlerc_group_enable_views before the exchange)", dst_rank); return RLERC_ERR_STATE; }
	rlerc_ctx* c = g->c;
	int rc = set_dev(c);
	if (rc) return rc;
	const int ticket = g->next_ticket;
	const int k = ticket % g->depth;
	const uint32_t gen = (uint32_t)(ticket / g->depth) + 1u;
	GroupSlot& S = g->slot[k];
	if (S.busy) { CKG(cudaEventSynchronize(S.done)); S.busy = false; }
	uint32_t* warp = (uint32_t*)((char*)g->warp + (size_t)k * g->warp_slot_bytes);
	uint8_t* img = g->img + (size_t)k * g->img_slot_bytes;
	cudaStream_t const main_stream = c->stream;
	const bool was_pipelined = c->pipelined;
	c->stream = S.stream;
	c->cur_states = &S.d_states; c->cur_states_bytes = &S.states_bytes;
	c->pipelined = g->depth > 1;
	if (c->timing) cudaEventRecord(S.t0, S.stream);
	// this member's OWN camera: the whole frame on this GPU
	rc = render_impl(c, rm, &g->cfg, 0, -1, warp, nullptr, false);
	if (!rc) rc = unwarp_impl(c, rm, &g->cfg, warp, img, 0, -1, 0, -1);
	const size_t frame_bytes = (size_t)g->cfg.width * g->cfg.height * 4;
	// deliver the finished view into slot [k][rank] of the destination over NVLink (peer mapping), then tell everybody
	if (!rc && dst_rank >= 0 && g->n > 1)
	{
		uint8_t* dst = g->views_peer[dst_rank] + ((size_t)k * g->n + g->rank) * g->img_slot_bytes;
		cudaError_t e = cudaMemcpyAsync(dst, img, frame_bytes, cudaMemcpyDefault, S.stream);
		if (e != cudaSuccess) { set_error("rlerc_group_submit_view: peer copy failed: %s", cudaGetErrorString(e)); rc = RLERC_ERR_CUDA; }
	}
	if (!rc && g->n > 1) rc = group_barrier(g, k, 1, gen, S.stream);
	if (c->timing) { cudaEventRecord(S.t1, S.stream); S.timed = true; }
	c->stream = main_stream;
	c->cur_states = nullptr; c->cur_states_bytes = nullptr;
	c->pipelined = was_pipelined;
	if (rc) return rc;
	CKG(cudaGetLastError());
	if (host_rgba) CKG(cudaMemcpyAsync(host_rgba, img, frame_bytes, cudaMemcpyDeviceToHost, S.stream));
	CKG(cudaEventRecord(S.done, S.stream));
	S.busy = true;
	g->next_ticket++;
	return ticket;
}

int rlerc_group_views(rlerc_group* g, int ticket, uint8_t** d_views, size_t* view_stride)
{
	if (!g || !d_views || ticket < 0 || !g->views) { set_error("rlerc_group_views: bad argument or views not enabled"); return RLERC_ERR_ARG; }
	*d_views = g->views + (size_t)(ticket % g->depth) * g->n * g->img_slot_bytes;
	if (view_stride) *view_stride = g->img_slot_bytes;
	return RLERC_OK;
}

int rlerc_group_wait(rlerc_group* g, int ticket)
{
	if (!g || ticket < 0 || ticket >= g->next_ticket) { set_error("rlerc_group_wait: bad ticket"); return RLERC_ERR_ARG; }
	if (ticket < g->next_ticket - g->depth) return RLERC_OK;      // its slot has been recycled: finished long ago
	int rc = set_dev(g->c);
	if (rc) return rc;
	GroupSlot& S = g->slot[ticket % g->depth];
	if (S.busy) { CKG(cudaEventSynchronize(S.done)); S.busy = false; }
	return RLERC_OK;
}

int rlerc_group_sync(rlerc_group* g)
{
	if (!g) return RLERC_ERR_ARG;
	int rc = set_dev(g->c);
	if (rc) return rc;
	for (auto& S : g->slot) { CKG(cudaStreamSynchronize(S.stream)); S.busy = false; }
	return RLERC_OK;
}

int rlerc_group_image(rlerc_group* g, int ticket, uint8_t** d_rgba, int* row_begin, int* row_end)
{
	if (!g || !d_rgba || ticket < 0) { set_error("rlerc_group_image: bad argument"); return RLERC_ERR_ARG; }
	*d_rgba = g->img + (size_t)(ticket % g->depth) * g->img_slot_bytes;
	int rb = 0, re = g->cfg.height;
	if (g->n > 1) band_of(g, g->rank, &rb, &re);
	if (row_begin) *row_begin = rb;
	if (row_end) *row_end = re;
	return RLERC_OK;
}

void* rlerc_group_stream(rlerc_group* g, int ticket)
{
	return (g && ticket >= 0) ? (void*)g->slot[ticket % g->depth].stream : nullptr;
}

int rlerc_group_last_ms(rlerc_group* g, int ticket, float* ms)
{
	if (!g || !ms || ticket < 0) return RLERC_ERR_ARG;
	int rc = set_dev(g->c);
	if (rc) return rc;
	GroupSlot& S = g->slot[ticket % g->depth];
	*ms = -1.0f;
	if (S.timed) { CKG(cudaEventSynchronize(S.t1)); CKG(cudaEventElapsedTime(ms, S.t0, S.t1)); }
	return RLERC_OK;
}

} // extern "C"

// ---- all members in ONE process: the multi-GPU frame behind the C ABI (rlerc_create(devices, n)) --------
// One worker thread per GPU issues that member's launches (five per frame): a single host thread issuing 5 N launches per
// frame is the bottleneck from four GPUs up (measured: 8 GPUs, 1080p: 0.6 ms of launches per frame).
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>

namespace {
struct MultiJob { rlerc_raymap rm; uint8_t* host; };
struct MultiWorker {
	rlerc_group* g = nullptr;
	std::thread th;
	std::mutex mu;
	std::condition_variable cv;
	std::deque<MultiJob> q;
	int done = 0;                    // frames this member has enqueued on its GPU
	int err = 0;
	std::string err_text;
	bool stop = false;
	void run()
	{
		for (;;)
		{
			MultiJob job;
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&] { return stop || !q.empty(); });
				if (q.empty()) return;
				job = q.front();
				q.pop_front();
			}
			const int t = rlerc_group_submit(g, &job.rm, -1, job.host);     // every GPU copies its own band into the host frame
			{
				std::lock_guard<std::mutex> lk(mu);
				if (t < 0 && !err) { err = t; err_text = rlerc_last_error(); }
				done++;
			}
			cv.notify_all();
		}
	}
};
} // namespace

struct rlerc_multi {
	std::vector<rlerc_ctx*> ctx;
	std::vector<MultiWorker*> worker;
	rlerc_frame_config cfg;
	bool have_cfg = false;
	int depth = 8, block = 32;
	int next_ticket = 0;
};

static void multi_drop_groups(rlerc_multi* m)
{
	for (MultiWorker* w : m->worker)
	{
		{ std::lock_guard<std::mutex> lk(w->mu); w->stop = true; }
		w->cv.notify_all();
		if (w->th.joinable()) w->th.join();
		rlerc_group_destroy(w->g);
		delete w;
	}
	m->worker.clear();
	m->have_cfg = false;
	m->next_ticket = 0;
}

static int multi_ensure_groups(rlerc_multi* m, const rlerc_frame_config* cfg)
{
	if (m->have_cfg && memcmp(&m->cfg, cfg, sizeof(*cfg)) == 0) return RLERC_OK;
	multi_drop_groups(m);
	const int n = (int)m->ctx.size();
	int rc = RLERC_OK;
	std::vector<rlerc_group*> grp;
	for (int r = 0; r < n && !rc; r++)
	{
		rlerc_group* g = nullptr;
		rc = rlerc_group_create(m->ctx[r], r, n, m->depth, m->block, cfg, &g);
		if (!rc) grp.push_back(g);
	}
	std::vector<char> blobs((size_t)n * RLERC_GROUP_BLOB_BYTES);
	for (int r = 0; r < n && !rc; r++) rc = rlerc_group_export(grp[r], blobs.data() + (size_t)r * RLERC_GROUP_BLOB_BYTES);
	for (int r = 0; r < n && !rc; r++) rc = rlerc_group_connect(grp[r], blobs.data());
	if (rc) { for (rlerc_group* g : grp) rlerc_group_destroy(g); return rc; }
	for (int r = 0; r < n; r++)
	{
		MultiWorker* w = new MultiWorker();
		w->g = grp[r];
		w->th = std::thread([w] { w->run(); });
		m->worker.push_back(w);
	}
	m->cfg = *cfg; m->have_cfg = true;
	return RLERC_OK;
}

extern "C" {

int rlerc_create_multi(const int* devices, int n, rlerc_multi** out)
{
	if (!devices || !out || n < 1 || n > RLERC_GROUP_MAX) { set_error("rlerc_create_multi: 1..%d devices", RLERC_GROUP_MAX); return RLERC_ERR_ARG; }
	*out = nullptr;
	for (int i = 0; i < n; i++)
		for (int j = 0; j < i; j++)
			if (devices[i] == devices[j]) { set_error("rlerc_create_multi: device %d listed twice", devices[i]); return RLERC_ERR_ARG; }
	rlerc_multi* m = new rlerc_multi();
	for (int i = 0; i < n; i++)
	{
		rlerc_ctx* c = nullptr;
		int rc = rlerc_create(devices[i], &c);
		if (rc) { rlerc_multi_destroy(m); return rc; }
		m->ctx.push_back(c);
	}
	*out = m;
	return RLERC_OK;
}

void rlerc_multi_destroy(rlerc_multi* m)
{
	if (!m) return;
	multi_drop_groups(m);
	for (rlerc_ctx* c : m->ctx) rlerc_destroy(c);
	delete m;
}

int rlerc_multi_count(const rlerc_multi* m) { return m ? (int)m->ctx.size() : 0; }

rlerc_ctx* rlerc_multi_ctx(rlerc_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

int rlerc_multi_set_depth(rlerc_multi* m, int depth, int slice_block)
{
	if (!m || depth < 1 || depth > 16 || slice_block < 1) { set_error("rlerc_multi_set_depth: bad argument"); return RLERC_ERR_ARG; }
	if (depth != m->depth || slice_block != m->block) multi_drop_groups(m);
	m->depth = depth; m->block = slice_block;
	return RLERC_OK;
}

int rlerc_multi_scene_upload(rlerc_multi* m, const rlerc_scene* s)
{
	if (!m || !s) { set_error("rlerc_multi_scene_upload: null argument"); return RLERC_ERR_ARG; }
	for (rlerc_ctx* c : m->ctx)
	{
		int rc = rlerc_scene_upload(c, s);       // a full replica in every GPU's HBM
		if (rc) return rc;
	}
	return RLERC_OK;
}

int rlerc_multi_frame_submit(rlerc_multi* m, const float pos[3], const float rot[3], const rlerc_frame_config* cfg, uint8_t* host_rgba)
{
	if (!m || !pos || !rot || !host_rgba) { set_error("rlerc_multi_frame_submit: null argument"); return RLERC_ERR_ARG; }
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if ((rc = multi_ensure_groups(m, cfg))) return rc;
	MultiJob job;
	memset(&job.rm, 0, sizeof(job.rm));
	if ((rc = rlerc_frame_setup(pos, rot, cfg, &job.rm))) return rc;
	job.host = host_rgba;
	// the same frame, in the same order, to every member's worker; the ticket is the frame's number
	for (MultiWorker* w : m->worker)
	{
		{ std::lock_guard<std::mutex> lk(w->mu); w->q.push_back(job); }
		w->cv.notify_all();
	}
	return m->next_ticket++;
}

int rlerc_multi_frame_wait(rlerc_multi* m, int ticket)
{
	if (!m || ticket < 0 || ticket >= m->next_ticket) { set_error("rlerc_multi_frame_wait: bad ticket"); return RLERC_ERR_ARG; }
	for (MultiWorker* w : m->worker)
	{
		{
			std::unique_lock<std::mutex> lk(w->mu);
			w->cv.wait(lk, [&] { return w->done > ticket; });
			if (w->err) { set_error("%s", w->err_text.c_str()); return w->err; }
		}
		int rc = rlerc_group_wait(w->g, ticket);
		if (rc) return rc;
	}
	return RLERC_OK;
}

int rlerc_multi_render_frame(rlerc_multi* m, const float pos[3], const float rot[3], const rlerc_frame_config* cfg, uint8_t* host_rgba)
{
	const int t = rlerc_multi_frame_submit(m, pos, rot, cfg, host_rgba);
	if (t < 0) return t;
	return rlerc_multi_frame_wait(m, t);
}

} // extern "C"
