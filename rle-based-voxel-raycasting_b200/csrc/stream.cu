// LOD streaming: mip-level residency by distance (include/rlerc.h "LOD streaming"; SURVEY.md §8f rank 4, BASELINE config 5).
//
// The reference keeps every level of the scene resident (RLE4::all_to_gpu, R/src/Rle4.cpp:432-438), and so does
// rlerc_scene_upload.  But a frame cannot touch all of it: the traversal's LOD schedule (Cuda_Render.h:343-367) is the same
// integer sequence for every ray plane — z counts crossings in steps of dz, dz doubles whenever z passes mapswitch — and a
// crossing at step dz moves a ray by at most dz level-0 voxels, so while a ray plane reads mip level m it is within
// z_end(m) voxels of the camera.  Level m therefore needs only the z-rows within that reach of the camera's row (all of
// them once the reach wraps around the grid: the coarse levels, which are small), and both of a level's arrays are laid
// out z-row by z-row (pointer map [sz][sx]; slabs in the same column order), so the needed part of each is one or two
// contiguous byte ranges.
//
// A streamed replica reserves VIRTUAL address space for every level (cuMemAddressReserve) and maps physical chunks of
// RLERC_STREAM_CHUNK bytes only where a frame needs them (cuMemCreate / cuMemMap, driver API through
// cudaGetDriverEntryPoint: the library does not link libcuda).  The kernels see the same pointers and the same bytes as
// with a full replica — nothing changes on the hot path, frames are bit-identical — and a frame that reached outside
// what rlerc_stream_prepare made resident would fault (a CUDA error from the next call), never draw a wrong picture.
// rlerc_stream_prepare(camera) maps + uploads what is missing for the reach plus a look-ahead margin and unmaps what is
// beyond twice the margin.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "rlerc_internal.h"
#include "capi_internal.cuh"

using namespace rlerc;

#ifndef RLERC_STREAM_CHUNK
#define RLERC_STREAM_CHUNK (4u << 20)
#endif

namespace {

struct Drv {                                  // the driver entry points of CUDA's virtual memory management
	CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
	CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
	CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
	CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
	CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
	CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
	CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
	CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
	bool ok = false;
};

Drv& drv()
{
	static Drv d;
	static bool tried = false;
	if (!tried)
	{
		tried = true;
		struct { const char* name; void** p; } syms[] = {
			{ "cuMemAddressReserve", (void**)&d.memAddressReserve }, { "cuMemAddressFree", (void**)&d.memAddressFree },
			{ "cuMemCreate", (void**)&d.memCreate }, { "cuMemRelease", (void**)&d.memRelease },
			{ "cuMemMap", (void**)&d.memMap }, { "cuMemUnmap", (void**)&d.memUnmap },
			{ "cuMemSetAccess", (void**)&d.memSetAccess }, { "cuMemGetAllocationGranularity", (void**)&d.memGetAllocationGranularity } };
		d.ok = true;
		for (auto& s : syms)
		{
			cudaDriverEntryPointQueryResult st;
			if (cudaGetDriverEntryPoint(s.name, s.p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*s.p) d.ok = false;
		}
		cudaGetLastError();
	}
	return d;
}

struct Range {                                // one array of one level in virtual memory
	CUdeviceptr va = 0;
	size_t va_bytes = 0;                      // reserved (multiple of the chunk size)
	size_t data_bytes = 0;                    // bytes that exist on the host
	const char* host = nullptr;
	std::vector<CUmemGenericAllocationHandle> chunk;    // 0 = not resident
};

struct StreamLevel {
	Range map, slabs;
	int sx = 0, sz = 0;
	const uint32_t* host_map = nullptr;       // offsets of the columns (slab range of a row band)
	uint64_t slab_count = 0;
	bool all_slabs = false;                   // a level with inconsistent columns (scene.cpp validate_columns): its attribute gathers may reach anywhere behind a column
};

} // namespace

struct rlerc_streamed {
	std::vector<StreamLevel> level;
	size_t chunk = RLERC_STREAM_CHUNK;
	int device = 0;
	uint64_t resident = 0, total = 0;
	bool synced_for_unmap = false;            // per rlerc_stream_prepare call
};

namespace {

#define CKD(call)                                                                       \
	do {                                                                                \
		CUresult r_ = (call);                                                           \
		if (r_ != CUDA_SUCCESS) { set_error("%s failed: CUresult %d (%s:%d)", #call, (int)r_, __FILE__, __LINE__); return RLERC_ERR_CUDA; } \
	} while (0)

int range_reserve(rlerc_streamed* st, Range& r, const void* host, size_t bytes, size_t pad)
{
	r.host = (const char*)host; r.data_bytes = bytes;
	r.va_bytes = (bytes + pad + st->chunk - 1) / st->chunk * st->chunk;
	CKD(drv().memAddressReserve(&r.va, r.va_bytes, st->chunk, 0, 0));
	r.chunk.assign(r.va_bytes / st->chunk, 0);
	st->total += bytes;
	return RLERC_OK;
}

// make bytes [b0, b1) of the range resident; everything else may go if `evict`
int range_want(rlerc_streamed* st, Range& r, size_t b0, size_t b1, std::vector<char>& keep)
{
	if (b1 > r.va_bytes) b1 = r.va_bytes;
	for (size_t k = b0 / st->chunk; k * st->chunk < b1; k++) keep[k] = 1;
	return RLERC_OK;
}

int range_apply(rlerc_streamed* st, Range& r, const std::vector<char>& need, const std::vector<char>& keep, cudaStream_t s, rlerc_stream_stats* out)
{
	CUmemAllocationProp prop;
	memset(&prop, 0, sizeof(prop));
	prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
	prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	prop.location.id = st->device;
	CUmemAccessDesc acc;
	memset(&acc, 0, sizeof(acc));
	acc.location = prop.location;
	acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	for (size_t k = 0; k < r.chunk.size(); k++)
	{
		const size_t off = k * st->chunk;
		if (need[k] && !r.chunk[k])
		{
			CUmemGenericAllocationHandle h;
			CKD(drv().memCreate(&h, st->chunk, &prop, 0));
			CUresult m = drv().memMap(r.va + off, st->chunk, 0, h, 0);
			if (m != CUDA_SUCCESS) { drv().memRelease(h); set_error("cuMemMap failed: CUresult %d", (int)m); return RLERC_ERR_CUDA; }
			CKD(drv().memSetAccess(r.va + off, st->chunk, &acc, 1));
			r.chunk[k] = h;
			const size_t n = off < r.data_bytes ? std::min(st->chunk, r.data_bytes - off) : 0;
			if (n) { cudaError_t e = cudaMemcpyAsync((void*)(r.va + off), r.host + off, n, cudaMemcpyHostToDevice, s); if (e != cudaSuccess) { set_error("upload of a streamed chunk failed: %s", cudaGetErrorString(e)); return RLERC_ERR_CUDA; } }
			if (n < st->chunk) cudaMemsetAsync((void*)(r.va + off + n), 0, st->chunk - n, s);       // padding behind the data
			st->resident += st->chunk;
			if (out) { out->uploaded_bytes += n; out->chunks_mapped++; }
		}
		else if (!need[k] && !keep[k] && r.chunk[k])
		{
			// nothing on this device may still be reading the chunk: frames in flight on other streams (rlerc_frame_submit,
			// groups) included, so the first eviction of a call waits for the whole device
			if (!st->synced_for_unmap) { cudaDeviceSynchronize(); st->synced_for_unmap = true; }
			CKD(drv().memUnmap(r.va + off, st->chunk));
			CKD(drv().memRelease(r.chunk[k]));
			r.chunk[k] = 0;
			st->resident -= st->chunk;
			if (out) { out->evicted_bytes += st->chunk; out->chunks_unmapped++; }
		}
	}
	return RLERC_OK;
}

void range_free(rlerc_streamed* st, Range& r)
{
	if (!drv().ok) return;
	for (size_t k = 0; k < r.chunk.size(); k++)
		if (r.chunk[k]) { drv().memUnmap(r.va + k * st->chunk, st->chunk); drv().memRelease(r.chunk[k]); r.chunk[k] = 0; }
	if (r.va) drv().memAddressFree(r.va, r.va_bytes);
	r.va = 0;
}

// rows [lo, hi] (any integers) of a grid of sz rows with wrap-around -> marks chunks of both arrays of the level
void want_rows(rlerc_streamed* st, StreamLevel& L, long long lo, long long hi, std::vector<char>& wm, std::vector<char>& ws)
{
	const long long sz = L.sz;
	if (hi - lo + 1 >= sz) { lo = 0; hi = sz - 1; }
	// split at the wrap
	long long a = ((lo % sz) + sz) % sz, n = hi - lo + 1;
	while (n > 0)
	{
		const long long m = std::min(n, sz - a);                       // rows a .. a + m - 1
		const size_t rowb = (size_t)L.sx * 8;
		range_want(st, L.map, (size_t)a * rowb, (size_t)(a + m) * rowb, wm);
		const uint64_t s0 = L.all_slabs ? 0 : L.host_map[(size_t)a * L.sx * 2];
		const uint64_t s1 = (a + m < sz && !L.all_slabs) ? L.host_map[(size_t)(a + m) * L.sx * 2] : L.slab_count;
		// + the look-ahead / gather padding a column's reads may reach behind its own data (capi.cu rlerc_scene_upload)
		range_want(st, L.slabs, (size_t)s0 * 2, (size_t)s1 * 2 + 64 + 4096, ws);
		n -= m; a = 0;
	}
}

} // namespace

namespace rlerc {
void stream_free(rlerc_ctx* c)
{
	if (!c->stream_state) return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	for (auto& L : c->stream_state->level) { range_free(c->stream_state, L.map); range_free(c->stream_state, L.slabs); }
	delete c->stream_state;
	c->stream_state = nullptr;
	c->nummaps = 0;
}
}

extern "C" {

int rlerc_scene_upload_streamed(rlerc_ctx* c, const rlerc_scene* s)
{
	if (!c || !s || s->levels.empty()) { set_error("rlerc_scene_upload_streamed: bad argument"); return RLERC_ERR_ARG; }
	int rc = set_dev(c);
	if (rc) return rc;
	if (!drv().ok) { set_error("rlerc_scene_upload_streamed: the CUDA driver's virtual memory management entry points are not available"); return RLERC_ERR_CUDA; }
	const int n = (int)s->levels.size();
	for (int m = 0; m < n; m++)
	{
		const Level& lv = s->levels[m];
		if (lv.sx < 1 || lv.sz < 1 || (lv.sx & (lv.sx - 1)) || (lv.sz & (lv.sz - 1)) || (m > 0 && (lv.sx != (s->levels[0].sx >> m) || lv.sz != (s->levels[0].sz >> m))))
		{
			set_error("level %d: grid %d x %d is not a power of two / not level 0 halved", m, lv.sx, lv.sz);
			return RLERC_ERR_FORMAT;
		}
		if (lv.slabs.size() > 0xffffffffull || lv.map.size() != (size_t)lv.sx * lv.sz * 2) { set_error("level %d: malformed", m); return RLERC_ERR_FORMAT; }
	}
	// replaces whatever replica the context held
	stream_free(c);
	for (void* p : c->scene_allocs) cudaFree(p);
	c->scene_allocs.clear();
	rlerc_streamed* st = new rlerc_streamed();
	st->device = c->device;
	CUmemAllocationProp prop;
	memset(&prop, 0, sizeof(prop));
	prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
	prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	prop.location.id = c->device;
	size_t gran = 0;
	if (drv().memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) gran = 2u << 20;
	st->chunk = (RLERC_STREAM_CHUNK + gran - 1) / gran * gran;
	st->level.resize(n);
	c->stream_state = st;
	for (int m = 0; m < n; m++)
	{
		const Level& lv = s->levels[m];
		StreamLevel& L = st->level[m];
		L.sx = lv.sx; L.sz = lv.sz; L.host_map = lv.map.data(); L.slab_count = lv.slabs.size(); L.all_slabs = lv.gather_pad > 0;
		if ((rc = range_reserve(st, L.map, lv.map.data(), lv.map.size() * 4, 0))) { stream_free(c); return rc; }
		if ((rc = range_reserve(st, L.slabs, lv.slabs.data(), lv.slabs.size() * 2, 64 + (size_t)lv.gather_pad * 2 + 4096))) { stream_free(c); return rc; }
		c->level[m].map = (const uint2*)L.map.va;
		c->level[m].slabs = (const uint16_t*)L.slabs.va;
		c->level[m].sx = lv.sx; c->level[m].sz = lv.sz;
		c->level_sy[m] = lv.sy;
		c->level_slabs[m] = lv.slabs.size();
	}
	c->nummaps = n;
	return RLERC_OK;
}

int rlerc_stream_prepare(rlerc_ctx* c, const float pos[3], const float rot[3], const rlerc_frame_config* cfg, int margin_voxels, rlerc_stream_stats* out)
{
	if (!c || !pos || !rot) { set_error("rlerc_stream_prepare: null argument"); return RLERC_ERR_ARG; }
	if (!c->stream_state) { set_error("rlerc_stream_prepare: the context holds no streamed scene (rlerc_scene_upload_streamed)"); return RLERC_ERR_STATE; }
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if ((rc = set_dev(c))) return rc;
	if (margin_voxels < 0) margin_voxels = 0;
	rlerc_streamed* st = c->stream_state;
	if (out) memset(out, 0, sizeof(*out));
	// the frame's LOD schedule, exactly as the traversal will follow it
	rlerc_raymap rm;
	memset(&rm, 0, sizeof(rm));
	if ((rc = rlerc_frame_setup(pos, rot, cfg, &rm))) return rc;
	TraverseParams P;
	if ((rc = fill_traverse(c, &rm, cfg, 0, -1, nullptr, P))) return rc;
	const int n = c->nummaps, last = n - 1;
	std::vector<long long> reach(n, -1);
	for (int p = 0; p < P.lod.nphase; p++)
	{
		const int nsw = P.lod.ph_nsw[p];
		const int mip = nsw < last ? nsw : last;
		// z when the phase ends = an upper bound of the distance travelled (a crossing at step dz moves <= dz voxels);
		// + the step itself and the integer snapping of the column address (Cuda_Render.h:429-442)
		const long long z_end = (p + 1 < P.lod.nphase) ? P.lod.ph_z[p + 1] : P.z_far;
		const long long r = z_end + (2ll << nsw) + 2;
		if (r > reach[mip]) reach[mip] = r;
	}
	CK(cudaStreamSynchronize(c->stream));
	st->synced_for_unmap = false;
	const double cz = (double)pos[2];
	for (int m = 0; m < n; m++)
	{
		StreamLevel& L = st->level[m];
		std::vector<char> need_m(L.map.chunk.size(), 0), need_s(L.slabs.chunk.size(), 0), keep_m(need_m), keep_s(need_s);
		if (reach[m] >= 0)
		{
			const double scale = (double)(1ll << m);
			const long long lo = (long long)floor((cz - (double)reach[m] - margin_voxels) / scale) - 1;
			const long long hi = (long long)floor((cz + (double)reach[m] + margin_voxels) / scale) + 1;
			want_rows(st, L, lo, hi, need_m, need_s);
			const long long lo2 = (long long)floor((cz - (double)reach[m] - 2.0 * margin_voxels) / scale) - 1;
			const long long hi2 = (long long)floor((cz + (double)reach[m] + 2.0 * margin_voxels) / scale) + 1;
			want_rows(st, L, lo2, hi2, keep_m, keep_s);
		}
		if ((rc = range_apply(st, L.map, need_m, keep_m, c->stream, out))) return rc;
		if ((rc = range_apply(st, L.slabs, need_s, keep_s, c->stream, out))) return rc;
	}
	CK(cudaStreamSynchronize(c->stream));
	if (out) { out->resident_bytes = st->resident; out->total_bytes = st->total; }
	return RLERC_OK;
}

} // extern "C"
