// C ABI of the render path (include/rlerc.h): context, device scene replica, per-frame
// parameter blocks, kernel launches, host<->device plumbing, and the legacy entry points of
// the reference (cuda_main_render2 & co, R/src/Cuda_Main.cu:124-148,183-271).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#include <map>
#include <vector>
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "device_common.cuh"
#include "rlerc_internal.h"
#include "capi_internal.cuh"

using namespace rlerc;

namespace rlerc {

int set_dev(rlerc_ctx* c)
{
	CK(cudaSetDevice(c->device));
	return RLERC_OK;
}

void free_scene(rlerc_ctx* c)
{
	stream_free(c);
	for (void* p : c->scene_allocs) cudaFree(p);
	c->scene_allocs.clear();
	c->nummaps = 0;
}

// (Re)allocate a device buffer.  A new buffer is zero-filled ON THE STREAM THAT WILL USE IT: the streams here are
// cudaStreamNonBlocking, which a memset on the legacy default stream does not order with.  zero = false: the buffer is
// fully written before it is read (DDA states).
int ensure(void** p, size_t* have, size_t need, cudaStream_t st, bool zero)
{
	if (*have >= need && *p) return RLERC_OK;
	if (*p) cudaFree(*p);                                        // synchronises the device: nothing still uses the old buffer
	*p = nullptr; *have = 0;
	CK(cudaMalloc(p, need));
	if (zero) CK(cudaMemsetAsync(*p, 0, need, st));
	*have = need;
	return RLERC_OK;
}

// The frame's LOD / z schedule (kernels.cuh LodSched): the integer half of the traversal loop, Cuda_Render.h:343-367,
// followed exactly as the kernels' ddaq_run32 steps it.
int fill_lod_sched(TraverseParams& P)
{
	LodSched& L = P.lod;
	memset(&L, 0, sizeof(L));
	int nsw = 0;
	for (float yms = P.viewpos[1]; yms > 512.0f; yms = yms * 0.5f)      // Cuda_Render.h:343: y_map_switch halves until <= 512
		if (++nsw > 24) { set_error("camera height %g is out of range", (double)P.viewpos[1]); return RLERC_ERR_ARG; }
	long long zi = 0, k = 0;
	int np = 0;
	while (true)
	{
		long long ms = (long long)P.mapswitch0 << nsw;
		while (zi > ms) { nsw++; ms *= 2; }
		if (nsw > 29 || ms > 0x7fffffffll) { set_error("z_far %d needs more LOD doublings than 32-bit z allows (mip_distance %d)", P.z_far, P.mapswitch0); return RLERC_ERR_ARG; }
		const long long lod_free = ((ms - zi) >> nsw) + 1;          // crossings before z > mapswitch
		const long long far_free = ((long long)P.z_far - zi) >> nsw; // crossings with z + dz <= z_far
		if (far_free <= 0) break;
		const long long n = lod_free < far_free ? lod_free : far_free;
		if (np == 32) { set_error("LOD schedule has more than 32 phases"); return RLERC_ERR_ARG; }
		L.ph_k[np] = (int)k; L.ph_z[np] = (int)zi; L.ph_nsw[np] = nsw; np++;
		k += n; zi += n << nsw;
	}
	L.nphase = np; L.k_total = (int)k; L.ph_k[np] = (int)k;
	return RLERC_OK;
}

int check_cfg(const rlerc_frame_config* cfg)
{
	if (!cfg) { set_error("null frame config"); return RLERC_ERR_ARG; }
	if (cfg->width < 4 || cfg->height < 1 || cfg->render_size < 32 || cfg->render_size > 16384 ||
	    cfg->rays_casted < 4 || cfg->rays_casted_res < 4 || cfg->z_far < 1 || cfg->mip_distance < 1 ||
	    (cfg->flags & ~(RLERC_FLAG_CLIPREGION | RLERC_FLAG_HEIGHT_COLOR | RLERC_FLAG_SHADER_2XAA)) != 0)
	{
		set_error("frame config out of range");
		return RLERC_ERR_ARG;
	}
	return RLERC_OK;
}

// Host half of the traversal set-up: everything that depends only on the camera.
int fill_traverse(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                  int ray_begin, int ray_end, uint32_t* d_warp, TraverseParams& P)
{
	if (c->nummaps < 1) { set_error("no scene uploaded"); return RLERC_ERR_STATE; }
	memset(&P, 0, sizeof(P));
	for (int m = 0; m < c->nummaps; m++) P.level[m] = c->level[m];
	P.nummaps = c->nummaps;
	int count = rm->map_line_count;
	if (count > cfg->rays_casted) count = cfg->rays_casted;      // Cuda_Main.cu:196
	if (count < 0) count = 0;
	for (int i = 0; i < 4; i++) P.res[i] = rm->res[i];
	P.vp[0] = rm->p_2d[5].x; P.vp[1] = rm->p_2d[5].y; P.vp[2] = rm->p_2d[5].z;
	for (int i = 0; i < 8; i++) { P.p_no[i][0] = rm->p_no[i].x; P.p_no[i][1] = rm->p_no[i].y; P.p_no[i][2] = rm->p_no[i].z; }
	P.clip_min = rm->clip_min; P.clip_max = rm->clip_max;
	memcpy(P.to3d, rm->to3d, sizeof(P.to3d));
	P.p4[0] = rm->p4.x; P.p4[1] = rm->p4.y; P.p4[2] = rm->p4.z;
	P.viewpos[0] = rm->position.x; P.viewpos[1] = rm->position.y; P.viewpos[2] = rm->position.z;
	const float rx = rm->rotation.x, ry = rm->rotation.y;
	// same overloads the host-compiled reference resolves to: sin(float) -> sinf etc.
	P.sin_x = std::sin(rx); P.cos_x = std::cos(rx);
	P.sin_y = std::sin(ry); P.cos_y = std::cos(ry);
	P.sin_my = std::sin(-ry); P.cos_my = std::cos(-ry);
	P.res_x = cfg->render_size; P.res_y = cfg->render_size;
	// Cuda_Render.h:182,335: int mapswitch = MIP_DISTANCE; mapswitch = mapswitch * (0.25*(4-abs(viewrot.x)));
	{
		int mapswitch = cfg->mip_distance;
		mapswitch = mapswitch * (0.25 * (4 - std::abs(rx)));
		P.mapswitch0 = mapswitch;
		// mapswitch doubles until it passes z: from 0 (or below) it never would, the reference loops forever there
		if (mapswitch < 1) { set_error("mip_distance %d at pitch %g gives mapswitch %d < 1", cfg->mip_distance, (double)rx, mapswitch); return RLERC_ERR_ARG; }
	}
	if (!std::isfinite(rm->position.x) || !std::isfinite(rm->position.y) || !std::isfinite(rm->position.z))
	{
		set_error("camera position is not finite");
		return RLERC_ERR_ARG;
	}
	P.z_far = cfg->z_far;
	if (ray_end < 0 || ray_end > count) ray_end = count;
	if (ray_begin < 0) ray_begin = 0;
	P.ray_begin = ray_begin; P.ray_end = ray_end;
	P.mask_words = (cfg->render_size + 31) / 32 + 1;
	P.flags = cfg->flags;
	P.warp = d_warp;
	P.slice_block = 1; P.slice_n = 1; P.slice_rank = 0;
	return fill_lod_sched(P);
}

// Uniforms of the colorize pass exactly as main.cpp:578-603 computes them.
int fill_unwarp(const rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, const uint32_t* d_warp, uint8_t* d_rgba, UnwarpParams& U)
{
	memset(&U, 0, sizeof(U));
	U.shade_rgb = c->d_shade_rgb; U.shade_alpha = c->d_shade_alpha;
	U.warp = d_warp; U.rgba = d_rgba;
	U.W = cfg->width; U.H = cfg->height;
	U.RS = cfg->render_size; U.RC = cfg->rays_casted;
	const float border = rm->border;
	{
		const float RESX = (float)cfg->width, RESY = (float)cfg->height;
		U.border = (RESX - RESY) / (RESX * 2);                            // colorize_buddha_soft.frag:22
	}
	U.vanish_x = 1 - rm->vanishing_point_2d.x;
	U.vanish_y = (1 - rm->vanishing_point_2d.y - border) * float(cfg->width) / float(cfg->height);
	const float ofs1 = 4 * float(rm->res[0]) / float(cfg->rays_casted_res);
	const float ofs2 = 4 * float(rm->res[1]) / float(cfg->rays_casted_res) + ofs1;
	const float ofs3 = 4 * float(rm->res[2]) / float(cfg->rays_casted_res) + ofs2;
	U.ofs_add[0] = -rm->p_ofs_min[0];
	U.ofs_add[1] = -rm->p_ofs_min[1] + ofs1;
	U.ofs_add[2] = -rm->p_ofs_min[2] + ofs2;
	U.ofs_add[3] = -rm->p_ofs_min[3] + ofs3;
	U.ratio = float(cfg->rays_casted_res) / float(cfg->rays_casted);
	U.shader = (cfg->flags & RLERC_FLAG_SHADER_2XAA) ? 1 : 0;
	if (U.shader == 1) U.ratio = 1.0f;                                  // colorize_buddha_soft_2xAA.frag:62 has no ratio factor (x * 1.0f is exact)
	U.rot_x_gt0 = (rm->rotation.x > 0) ? 1 : 0;
	U.row_begin = 0; U.row_end = cfg->height;
	U.ray_begin = 0; U.ray_end = -1;
	U.slice_block = 1; U.slice_n = 1; U.slice_rank = 0;
	U.peer_n = 0;
	U.generic = (std::fabs(U.vanish_x) < 1e6f && std::fabs(U.vanish_y) < 1e6f) ? 0 : 1;
	return RLERC_OK;
}

// 0 = automatic: the production kernel k_traverse_f (one warp per ray plane), except for launches so small that every
// ray plane is resident at once with warps to spare (at most 5 per SM): such a launch is bound by the serial chain of its
// slowest ray planes, which the four-role kernel k_traverse_q (filter / project / resolve / shade warps per ray plane)
// shortens (measured on B200, round 2: 1/8 of a 1080p frame 0.42 -> 0.34 ms, 1/8 of a 4K frame 0.64 -> 0.56 ms, 1/16
// 0.66 -> 0.48 ms; a full frame is throughput-bound and 1.7x slower on it).  That is the multi-GPU slice mode at 8 GPUs,
// one frame at a time, and small windows.  Never when frames overlap (pipelined): then the GPU is throughput-bound.
// 65 = k_traverse_f always, 68 = k_traverse_p (two warps per ray plane), 69 = k_traverse_q always, 64 = k_traverse_w,
// 66/67 = k_traverse_c, else k_traverse<lanes>.
int pick_lanes(const rlerc_ctx* c, int rays, bool ids)
{
	if (c->lanes != 0) return c->lanes;
	if (!ids && !c->pipelined && c->dda_mode != 99 && rays > 0 && rays <= 5 * c->sm_count) return 69;
	return 0;
}

} // namespace rlerc

extern "C" {

int rlerc_create(int device, rlerc_ctx** out)
{
	if (!out) { set_error("rlerc_create: null out"); return RLERC_ERR_ARG; }
	*out = nullptr;
	// One stream per frame in flight (rlerc_frame_submit, groups): with the default of 8 hardware queues streams share
	// queues, and a kernel that waits (k_group_barrier) holds up unrelated kernels queued behind it.  Only effective if
	// the CUDA context does not exist yet; a host that created it earlier sets the variable itself.
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n < 1) { set_error("no CUDA device: %s", cudaGetErrorString(e)); return RLERC_ERR_CUDA; }
	if (device < 0 || device >= n) { set_error("device %d out of range (%d devices)", device, n); return RLERC_ERR_ARG; }
	rlerc_ctx* c = new rlerc_ctx();
	c->device = device;
	CK(cudaSetDevice(device));
	CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
	CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < 4; i++) CK(cudaEventCreate(&c->ev[i]));
	for (int i = 0; i < rlerc_ctx::kSlots; i++) CK(cudaEventCreateWithFlags(&c->slot[i].done, cudaEventDisableTiming));
	CK(cudaMalloc((void**)&c->d_counters, 32 * sizeof(unsigned long long)));
	CK(cudaMalloc((void**)&c->d_shade_rgb, 65536 * sizeof(uint32_t)));
	CK(cudaMalloc((void**)&c->d_shade_alpha, 65536));
	launch_shade_tables(c->d_shade_rgb, c->d_shade_alpha, c->stream);
	CK(cudaStreamSynchronize(c->stream));
	*out = c;
	return RLERC_OK;
}

void rlerc_destroy(rlerc_ctx* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	free_scene(c);
	if (c->d_warp) cudaFree(c->d_warp);
	if (c->d_rgba) cudaFree(c->d_rgba);
	if (c->d_counters) cudaFree(c->d_counters);
	if (c->d_shade_rgb) cudaFree(c->d_shade_rgb);
	if (c->d_shade_alpha) cudaFree(c->d_shade_alpha);
	if (c->d_states) cudaFree(c->d_states);
	if (c->d_ring) cudaFree(c->d_ring);
	if (c->d_ring_ctl) cudaFree(c->d_ring_ctl);
	for (int i = 0; i < rlerc_ctx::kSlots; i++)
	{
		if (c->slot[i].d_warp) cudaFree(c->slot[i].d_warp);
		if (c->slot[i].d_rgba) cudaFree(c->slot[i].d_rgba);
		if (c->slot[i].d_states) cudaFree(c->slot[i].d_states);
		if (c->slot[i].done) cudaEventDestroy(c->slot[i].done);
		if (c->slot[i].stream) cudaStreamDestroy(c->slot[i].stream);
	}
	for (int i = 0; i < 4; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	delete c;
}

int rlerc_scene_upload(rlerc_ctx* c, const rlerc_scene* s)
{
	if (!c || !s || s->levels.empty()) { set_error("rlerc_scene_upload: bad argument"); return RLERC_ERR_ARG; }
	int rc = set_dev(c);
	if (rc) return rc;
	// the traversal wraps coordinates with & (grid-1) and halves the grid per level
	// (Cuda_Render.h:351-352,441-442): validate instead of rendering garbage
	const int n = (int)s->levels.size();
	for (int m = 0; m < n; m++)
	{
		const Level& lv = s->levels[m];
		if (lv.sx < 1 || lv.sz < 1 || (lv.sx & (lv.sx - 1)) || (lv.sz & (lv.sz - 1)))
		{
			set_error("level %d: grid %d x %d is not a power of two", m, lv.sx, lv.sz);
			return RLERC_ERR_FORMAT;
		}
		if (m > 0 && (lv.sx != (s->levels[0].sx >> m) || lv.sz != (s->levels[0].sz >> m)))
		{
			set_error("level %d: grid %d x %d is not level 0 halved %d times", m, lv.sx, lv.sz, m);
			return RLERC_ERR_FORMAT;
		}
		if (lv.sy > 65535) { set_error("level %d: sy %d exceeds 65535", m, lv.sy); return RLERC_ERR_FORMAT; }
		if (lv.slabs.size() > 0xffffffffull) { set_error("level %d: slab stream exceeds 32-bit offsets", m); return RLERC_ERR_FORMAT; }
		if (lv.map.size() != (size_t)lv.sx * lv.sz * 2) { set_error("level %d: pointer map size mismatch", m); return RLERC_ERR_FORMAT; }
	}
	free_scene(c);
	for (int m = 0; m < n; m++)
	{
		const Level& lv = s->levels[m];
		void *dm = nullptr, *ds = nullptr;
		// +64 bytes of zero padding behind the slabs: a column's "first run" style look-ahead
		// and vector loads may read a few words past the last column; + what the attribute gathers of
		// columns with inconsistent counts can reach (scene.cpp validate_columns)
		const size_t mbytes = lv.map.size() * 4, sbytes = lv.slabs.size() * 2, pad = 64 + (size_t)lv.gather_pad * 2;
		CK(cudaMalloc(&dm, mbytes));
		c->scene_allocs.push_back(dm);
		CK(cudaMalloc(&ds, sbytes + pad));
		c->scene_allocs.push_back(ds);
		CK(cudaMemcpy(dm, lv.map.data(), mbytes, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(ds, lv.slabs.data(), sbytes, cudaMemcpyHostToDevice));
		CK(cudaMemset((char*)ds + sbytes, 0, pad));
		c->level[m].map = (const uint2*)dm;
		c->level[m].slabs = (const uint16_t*)ds;
		c->level[m].sx = lv.sx;
		c->level[m].sz = lv.sz;
		c->level_sy[m] = lv.sy;
		c->level_slabs[m] = lv.slabs.size();
	}
	c->nummaps = n;
	return RLERC_OK;
}

int rlerc_scene_share(rlerc_ctx* dst, const rlerc_ctx* src)
{
	if (!dst || !src || dst == src) { set_error("rlerc_scene_share: bad argument"); return RLERC_ERR_ARG; }
	if (dst->device != src->device) { set_error("rlerc_scene_share: contexts are on different devices (%d, %d)", dst->device, src->device); return RLERC_ERR_ARG; }
	if (src->nummaps < 1) { set_error("rlerc_scene_share: the source context has no scene"); return RLERC_ERR_STATE; }
	int rc = set_dev(dst);
	if (rc) return rc;
	free_scene(dst);
	for (int m = 0; m < src->nummaps; m++)
	{
		dst->level[m] = src->level[m];
		dst->level_sy[m] = src->level_sy[m];
		dst->level_slabs[m] = src->level_slabs[m];
	}
	dst->nummaps = src->nummaps;
	return RLERC_OK;
}

int rlerc_scene_device_maps(rlerc_ctx* c, rlerc_map4* out16, int* nummaps)
{
	if (!c || !out16) { set_error("rlerc_scene_device_maps: bad argument"); return RLERC_ERR_ARG; }
	memset(out16, 0, sizeof(rlerc_map4) * RLERC_MAX_MAPS);
	for (int m = 0; m < c->nummaps; m++)
	{
		out16[m].sx = c->level[m].sx; out16[m].sy = c->level_sy[m]; out16[m].sz = c->level[m].sz;
		out16[m].slabs_size = (int32_t)(uint32_t)c->level_slabs[m];
		out16[m].map = (uint32_t*)c->level[m].map;
		out16[m].slabs = (uint16_t*)c->level[m].slabs;
	}
	if (nummaps) *nummaps = c->nummaps;
	return RLERC_OK;
}

int rlerc_has_variants(void) { return RLERC_VARIANTS ? 1 : 0; }

int rlerc_set_lanes_per_ray(rlerc_ctx* c, int lanes)
{
	if (!c || !(lanes == 0 || lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8 || lanes == 16 || lanes == 32 || (lanes >= 64 && lanes <= 69)))
	{
		set_error("lanes per ray must be 0,1,2,4,8,16,32 or a kernel code 64..69");
		return RLERC_ERR_ARG;
	}
	if (!RLERC_VARIANTS && (lanes == 64 || lanes == 66 || lanes == 67))
	{
		set_error("kernel variant %d is not in this build (make VARIANTS=1)", lanes);
		return RLERC_ERR_ARG;
	}
	c->lanes = lanes;
	return RLERC_OK;
}

const char* rlerc_last_kernel(rlerc_ctx* c) { return c ? c->last_kernel : ""; }

int rlerc_set_dda_producer(rlerc_ctx* c, int on)
{
	if (!c) return RLERC_ERR_ARG;
	c->producer = on != 0;
	return RLERC_OK;
}

int rlerc_set_dda_mode(rlerc_ctx* c, int mode)
{
	if (!c || (mode != 0 && mode != 2 && mode != 3)) { set_error("dda mode must be 0 (serial), 2 (merge path) or 3 (closed form)"); return RLERC_ERR_ARG; }
	c->dda_mode = mode;
	return RLERC_OK;
}

int rlerc_set_timing(rlerc_ctx* c, int on)
{
	if (!c) return RLERC_ERR_ARG;
	c->timing = on != 0;
	return RLERC_OK;
}

int rlerc_warp_buffer(rlerc_ctx* c, const rlerc_frame_config* cfg, uint32_t** d_warp)
{
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if (!c || !d_warp) { set_error("rlerc_warp_buffer: null argument"); return RLERC_ERR_ARG; }
	if ((rc = set_dev(c))) return rc;
	rc = ensure((void**)&c->d_warp, &c->warp_bytes, (size_t)cfg->rays_casted * cfg->render_size * 4, c->stream);
	if (rc) return rc;
	*d_warp = c->d_warp;
	return RLERC_OK;
}

} // extern "C"
int rlerc::render_impl(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                       int ray_begin, int ray_end, uint32_t* d_warp, uint32_t* d_ids, bool ids,
                       int slice_block, int slice_n, int slice_rank, uint32_t* prof_out)
{
	if (!c || !rm) { set_error("rlerc_render: null argument"); return RLERC_ERR_ARG; }
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if ((rc = set_dev(c))) return rc;
	if (!d_warp)
	{
		rc = ensure((void**)&c->d_warp, &c->warp_bytes, (size_t)cfg->rays_casted * cfg->render_size * 4, c->stream);
		if (rc) return rc;
		d_warp = c->d_warp;
	}
	TraverseParams P;
	if ((rc = fill_traverse(c, rm, cfg, ray_begin, ray_end, d_warp, P))) return rc;
	if ((cfg->flags & (RLERC_FLAG_CLIPREGION | RLERC_FLAG_HEIGHT_COLOR)) != 0 && !(c->lanes == 0 || c->lanes == 65 || c->lanes == 68 || c->lanes == 69))
	{
		set_error("frame config flags (CLIPREGION / HEIGHT_COLOR) are implemented by the production traversal kernels only (lanes_per_ray = 0, 65, 68)");
		return RLERC_ERR_ARG;
	}
	if (slice_n > 1) { P.slice_block = slice_block; P.slice_n = slice_n; P.slice_rank = slice_rank; P.ray_begin = 0; }
	if (ids)
	{
		if (!d_ids) { set_error("rlerc_render_ids: null ids buffer"); return RLERC_ERR_ARG; }
		P.ids = d_ids;
		P.counters = c->d_counters;
		CK(cudaMemsetAsync(c->d_counters, 0, 32 * sizeof(unsigned long long), c->stream));
	}
	P.dda_mode = c->dda_mode;
	if (prof_out) P.ids = prof_out;
#if RLERC_VARIANTS
	if (c->lanes == 64 && c->producer)
	{
		// one ring per context: the pipelined path (several frames in flight on separate streams) cannot share it
		if (c->cur_states) { set_error("the DDA producer variant (lanes 64) cannot be used with rlerc_frame_submit"); return RLERC_ERR_STATE; }
		const int cap = cfg->rays_casted;
		if ((rc = ensure((void**)&c->d_ring, &c->ring_bytes, traverse_ring_bytes(cap), c->stream))) return rc;
		if ((rc = ensure((void**)&c->d_ring_ctl, &c->ring_ctl_bytes, (size_t)(2 * cap + 4) * sizeof(int), c->stream))) return rc;
		CK(cudaMemsetAsync(c->d_ring_ctl, 0, (size_t)(2 * cap + 4) * sizeof(int), c->stream));
		P.dda_ring = c->d_ring; P.dda_head = c->d_ring_ctl; P.dda_tail = c->d_ring_ctl + cap; P.dda_err = c->d_ring_ctl + 2 * cap;
	}
#endif
	const int launch_rays = (P.slice_n > 1) ? owned_count(P.ray_end, P.slice_block, P.slice_n, P.slice_rank) : P.ray_end - P.ray_begin;
	const int lanes = pick_lanes(c, launch_rays, ids);
	const bool production = lanes == 0 || lanes == 65 || lanes == 68 || lanes == 69;
	if (production)
	{
		// DDA states of this launch's ray planes: sized for the most ray planes a launch of this configuration can have,
		// so that the buffer is allocated once (a reallocation synchronises the device)
		// (most ray planes: rays_casted; most crossings: pitch 0, camera below y = 512 -> mapswitch = mip_distance, no
		// initial LOD switch)
		TraverseParams Pmax = P;
		Pmax.ray_begin = 0; Pmax.ray_end = cfg->rays_casted;
		Pmax.slice_n = 1;
		Pmax.mapswitch0 = cfg->mip_distance; Pmax.viewpos[1] = 0.0f;
		if (fill_lod_sched(Pmax) != RLERC_OK || Pmax.lod.k_total < P.lod.k_total) Pmax.lod = P.lod;
		float2** sp = c->cur_states ? c->cur_states : &c->d_states;
		size_t* sb = c->cur_states_bytes ? c->cur_states_bytes : &c->states_bytes;
		const size_t state_bytes = (traverse_dda_state_bytes(Pmax) + 255) & ~(size_t)255;
		const bool fresh = *sb < state_bytes + traverse_dda_progress_bytes(Pmax) || !*sp;
		if ((rc = ensure((void**)sp, sb, state_bytes + traverse_dda_progress_bytes(Pmax), c->stream, false))) return rc;
		P.dda_states = *sp;
		P.dda_progress = (unsigned long long*)((char*)*sp + state_bytes);
		// a new buffer's progress words are garbage that could carry any epoch
		if (fresh) CK(cudaMemsetAsync(P.dda_progress, 0, traverse_dda_progress_bytes(Pmax), c->stream));
		P.dda_epoch = ++c->dda_epoch;
		if (P.dda_epoch == 0) P.dda_epoch = ++c->dda_epoch;
	}
	if (c->timing) CK(cudaEventRecord(c->ev[0], c->stream));
	if (production) launch_dda_states(P, c->stream);
	launch_traverse(P, lanes, ids, c->stream);
	if (c->timing) { CK(cudaEventRecord(c->ev[1], c->stream)); c->ev_valid[0] = true; }
	c->last_kernel = (lanes == 0 || lanes == 65) ? (ids ? "k_traverse_f<ids>" : "k_traverse_f") : lanes == 68 ? (ids ? "k_traverse_f<ids>" : "k_traverse_p") : lanes == 69 ? (ids ? "k_traverse_f<ids>" : "k_traverse_q")
	               : lanes == 64 ? "k_traverse_w" : lanes == 66 ? "k_traverse_c<3>" : lanes == 67 ? "k_traverse_c<7>" : "k_traverse<G>";
	CK(cudaGetLastError());
	return RLERC_OK;
}
extern "C" {

int rlerc_render(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, int ray_begin, int ray_end, uint32_t* d_warp)
{
	return render_impl(c, rm, cfg, ray_begin, ray_end, d_warp, nullptr, false);
}

int rlerc_render_ids(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, int ray_begin, int ray_end, uint32_t* d_warp, uint32_t* d_ids)
{
	return render_impl(c, rm, cfg, ray_begin, ray_end, d_warp, d_ids, true);
}

int rlerc_render_counters(rlerc_ctx* c, uint64_t out[10])
{
	if (!c || !out) return RLERC_ERR_ARG;
	int rc = set_dev(c);
	if (rc) return rc;
	CK(cudaStreamSynchronize(c->stream));
	unsigned long long h[10];
	CK(cudaMemcpy(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
	for (int i = 0; i < 10; i++) out[i] = h[i];
	return RLERC_OK;
}

/* undocumented (tools/ray_profile.py): k_traverse_f with clock64() phase timers; d_out = unsigned long long[rays_casted][24] */
int rlerc_debug_profile_rays(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, void* d_out)
{
	if (!c || !d_out) return RLERC_ERR_ARG;
	const int saved_mode = c->dda_mode, saved_lanes = c->lanes;
	c->dda_mode = 99; c->lanes = 0;
	// RLERC_PROF_Q=1: the four-role kernel's own timers (a library built with -DRLERC_Q_PROF=1, tools/quad_profile.py)
	if (getenv("RLERC_PROF_Q")) { c->dda_mode = saved_mode; c->lanes = 69; }
	// RLERC_PROF_SLICE=k: only every k-th ray plane (the uncontended chain: every warp has an SM sub-partition to itself)
	const char* sl = getenv("RLERC_PROF_SLICE");
	const int k = sl ? atoi(sl) : 1;
	const int rc = render_impl(c, rm, cfg, 0, -1, nullptr, nullptr, false, 1, k > 1 ? k : 1, 0, (uint32_t*)d_out);
	c->dda_mode = saved_mode; c->lanes = saved_lanes;
	return rc;
}

/* undocumented (tests/test_lod_sched.py): the LOD / z schedule fill_traverse computes for a frame.
 * out = { nphase, k_total, ph_k[33], ph_z[32], ph_nsw[32] } = 99 ints */
int rlerc_debug_lod_sched(int mapswitch0, int z_far, float mountain, int* out)
{
	if (!out) return RLERC_ERR_ARG;
	TraverseParams P;
	memset(&P, 0, sizeof(P));
	P.mapswitch0 = mapswitch0; P.z_far = z_far; P.viewpos[1] = mountain;
	int rc = fill_lod_sched(P);
	if (rc) return rc;
	memcpy(out, &P.lod, sizeof(P.lod));
	return RLERC_OK;
}

/* undocumented (tools/l2_peak.py): read bandwidth in GB/s of `iters` passes over a zero-filled buffer of `mbytes` MB
 * (<= 64 MB: L2-resident after the warm-up pass; >= 1024 MB: HBM) */
int rlerc_debug_read_gbs(rlerc_ctx* c, int mbytes, int iters, double* gbs)
{
	if (!c || !gbs || mbytes < 1 || iters < 1) return RLERC_ERR_ARG;
	int rc = set_dev(c);
	if (rc) return rc;
	const size_t bytes = (size_t)mbytes << 20;
	void* buf = nullptr;
	CK(cudaMalloc(&buf, bytes));
	CK(cudaMemsetAsync(buf, 0, bytes, c->stream));
	launch_read_u4(buf, bytes, 2, (uint32_t*)c->d_counters, c->stream);       // warm-up: brings the buffer into L2
	CK(cudaEventRecord(c->ev[0], c->stream));
	launch_read_u4(buf, bytes, iters, (uint32_t*)c->d_counters, c->stream);
	CK(cudaEventRecord(c->ev[1], c->stream));
	CK(cudaEventSynchronize(c->ev[1]));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
	cudaFree(buf);
	*gbs = (double)bytes * iters / (ms * 1e-3) / 1e9;
	return RLERC_OK;
}

/* undocumented: raw copy of all 32 debug counter slots (tools/ only) */
int rlerc_debug_counters(rlerc_ctx* c, uint64_t out[32])
{
	if (!c || !out) return RLERC_ERR_ARG;
	int rc = set_dev(c);
	if (rc) return rc;
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaMemcpy(out, c->d_counters, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	return RLERC_OK;
}

} // extern "C"
int rlerc::unwarp_impl(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, const uint32_t* d_warp,
                       uint8_t* d_rgba, int row_begin, int row_end, int ray_begin, int ray_end,
                       int slice_block, int slice_n, int slice_rank)
{
	if (!c || !rm) { set_error("rlerc_unwarp: null argument"); return RLERC_ERR_ARG; }
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if ((rc = set_dev(c))) return rc;
	if (!d_warp)
	{
		if (!c->d_warp) { set_error("rlerc_unwarp: nothing rendered yet"); return RLERC_ERR_STATE; }
		d_warp = c->d_warp;
	}
	if (!d_rgba)
	{
		rc = ensure((void**)&c->d_rgba, &c->rgba_bytes, (size_t)cfg->width * cfg->height * 4, c->stream);
		if (rc) return rc;
		d_rgba = c->d_rgba;
	}
	UnwarpParams U;
	fill_unwarp(c, rm, cfg, d_warp, d_rgba, U);
	if (row_end < 0 || row_end > cfg->height) row_end = cfg->height;
	if (row_begin < 0) row_begin = 0;
	U.row_begin = row_begin; U.row_end = row_end;
	U.ray_begin = ray_begin; U.ray_end = ray_end;
	if (slice_n > 1) { U.slice_block = slice_block; U.slice_n = slice_n; U.slice_rank = slice_rank; }
	if (c->timing) CK(cudaEventRecord(c->ev[2], c->stream));
	launch_unwarp(U, c->stream);
	if (c->timing) { CK(cudaEventRecord(c->ev[3], c->stream)); c->ev_valid[1] = true; }
	CK(cudaGetLastError());
	return RLERC_OK;
}
extern "C" {

int rlerc_unwarp(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, const uint32_t* d_warp, uint8_t* d_rgba, int row_begin, int row_end)
{
	return unwarp_impl(c, rm, cfg, d_warp, d_rgba, row_begin, row_end, 0, -1);
}

/* undocumented (tests): the texel every window pixel samples, uint32[height][width] = ray plane << 16 | position along the ray */
int rlerc_debug_unwarp_texels(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, uint32_t* d_out)
{
	if (!c || !rm || !d_out) return RLERC_ERR_ARG;
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if ((rc = set_dev(c))) return rc;
	UnwarpParams U;
	fill_unwarp(c, rm, cfg, nullptr, (uint8_t*)d_out, U);
	launch_unwarp(U, c->stream, true);
	CK(cudaGetLastError());
	return RLERC_OK;
}

int rlerc_unwarp_slice(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, const uint32_t* d_warp, uint8_t* d_rgba, int ray_begin, int ray_end)
{
	if (ray_begin < 0 || ray_end < ray_begin) { set_error("rlerc_unwarp_slice: bad ray range"); return RLERC_ERR_ARG; }
	return unwarp_impl(c, rm, cfg, d_warp, d_rgba, 0, -1, ray_begin, ray_end);
}

static int check_slice(int block, int nranks, int rank)
{
	if (block < 1 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad interleaved slice (block %d, nranks %d, rank %d)", block, nranks, rank); return RLERC_ERR_ARG; }
	return RLERC_OK;
}

int rlerc_render_interleaved(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, int block, int nranks, int rank, uint32_t* d_warp)
{
	int rc = check_slice(block, nranks, rank);
	if (rc) return rc;
	return render_impl(c, rm, cfg, 0, -1, d_warp, nullptr, false, block, nranks, rank);
}

int rlerc_unwarp_interleaved(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, int block, int nranks, int rank, const uint32_t* d_warp, uint8_t* d_rgba)
{
	int rc = check_slice(block, nranks, rank);
	if (rc) return rc;
	return unwarp_impl(c, rm, cfg, d_warp, d_rgba, 0, -1, 0, -1, block, nranks, rank);
}

int rlerc_frame_device(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg, int block, int nranks, int rank, uint8_t* d_rgba)
{
	int rc = check_slice(block, nranks, rank);
	if (rc) return rc;
	if ((rc = render_impl(c, rm, cfg, 0, -1, nullptr, nullptr, false, block, nranks, rank))) return rc;
	return unwarp_impl(c, rm, cfg, nullptr, d_rgba, 0, -1, 0, -1, block, nranks, rank);
}

int rlerc_set_stream(rlerc_ctx* c, void* cuda_stream)
{
	if (!c) return RLERC_ERR_ARG;
	int rc = set_dev(c);
	if (rc) return rc;
	CK(cudaStreamSynchronize(c->stream));
	if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
	c->stream = (cudaStream_t)cuda_stream;
	return RLERC_OK;
}

// soft.frag:16: for (float a = 0; a < 3.1415*2.0; a += 3.1415*1.9/6.0) -> 7 taps, evaluated in float like GLSL
static void soft_taps(float* tx, float* ty)
{
	float a = 0.0f;
	const float step = (3.1415f * 1.9f) / 6.0f, lim = 3.1415f * 2.0f;
	for (int k = 0; k < 7 && a < lim; k++, a += step) { tx[k] = std::sin(a) * 0.005f; ty[k] = std::cos(a) * 0.005f; }
}

int rlerc_soft(rlerc_ctx* c, const rlerc_frame_config* cfg, const uint8_t* d_rgba_in, uint8_t* d_rgba_out)
{
	if (!c || !d_rgba_in || !d_rgba_out || d_rgba_in == d_rgba_out) { set_error("rlerc_soft: bad argument (in-place is not supported)"); return RLERC_ERR_ARG; }
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if ((rc = set_dev(c))) return rc;
	SoftParams S;
	memset(&S, 0, sizeof(S));
	S.in = d_rgba_in; S.out = d_rgba_out; S.W = cfg->width; S.H = cfg->height;
	S.fbo = 2048;                                  // static FBO fbo1(2048,2048), R/src/main.cpp:540
	while (S.fbo < cfg->width || S.fbo < cfg->height) S.fbo *= 2;
	soft_taps(S.tap_x, S.tap_y);
	launch_soft(S, c->stream);
	CK(cudaGetLastError());
	return RLERC_OK;
}

int rlerc_render_frame(rlerc_ctx* c, const float pos[3], const float rot[3], const rlerc_frame_config* cfg, uint8_t* host_rgba, rlerc_raymap* out_raymap)
{
	if (!c || !pos || !rot || !host_rgba) { set_error("rlerc_render_frame: null argument"); return RLERC_ERR_ARG; }
	int rc = check_cfg(cfg);
	if (rc) return rc;
	rlerc_raymap rm;
	memset(&rm, 0, sizeof(rm));
	if ((rc = rlerc_frame_setup(pos, rot, cfg, &rm))) return rc;
	if ((rc = render_impl(c, &rm, cfg, 0, -1, nullptr, nullptr, false))) return rc;
	if ((rc = unwarp_impl(c, &rm, cfg, nullptr, nullptr, 0, -1, 0, -1))) return rc;
	CK(cudaMemcpyAsync(host_rgba, c->d_rgba, (size_t)cfg->width * cfg->height * 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	if (out_raymap) *out_raymap = rm;
	return RLERC_OK;
}

int rlerc_frame_submit(rlerc_ctx* c, const float pos[3], const float rot[3], const rlerc_frame_config* cfg, uint8_t* host_rgba)
{
	if (!c || !pos || !rot || !host_rgba) { set_error("rlerc_frame_submit: null argument"); return RLERC_ERR_ARG; }
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if ((rc = set_dev(c))) return rc;
	const int ticket = c->next_ticket;
	FrameSlot& s = c->slot[ticket % rlerc_ctx::kSlots];
	if (s.busy) { CK(cudaEventSynchronize(s.done)); s.busy = false; }
	rlerc_raymap rm;
	memset(&rm, 0, sizeof(rm));
	if ((rc = rlerc_frame_setup(pos, rot, cfg, &rm))) return rc;
	// a caller-provided stream (rlerc_set_stream) is respected; otherwise every slot has its own
	cudaStream_t const main_stream = c->stream;
	if (c->own_stream)
	{
		if (!s.stream) CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
		c->stream = s.stream;
	}
	cudaStream_t const fs = c->stream;
	c->cur_states = &s.d_states; c->cur_states_bytes = &s.states_bytes;
	const bool was_pipelined = c->pipelined;
	c->pipelined = true;                    // frames overlap: throughput-bound, the paired kernel would only add instructions
	rc = ensure((void**)&s.d_warp, &s.warp_bytes, (size_t)cfg->rays_casted * cfg->render_size * 4, fs);
	if (!rc) rc = ensure((void**)&s.d_rgba, &s.rgba_bytes, (size_t)cfg->width * cfg->height * 4, fs);
	if (!rc) rc = render_impl(c, &rm, cfg, 0, -1, s.d_warp, nullptr, false);
	if (!rc) rc = unwarp_impl(c, &rm, cfg, s.d_warp, s.d_rgba, 0, -1, 0, -1);
	c->stream = main_stream;
	c->cur_states = nullptr; c->cur_states_bytes = nullptr;
	c->pipelined = was_pipelined;
	if (rc) return rc;
	// hand the finished frame to the copy stream so the next frame's traversal overlaps the D2H
	CK(cudaEventRecord(s.done, fs));
	CK(cudaStreamWaitEvent(c->copy_stream, s.done, 0));
	CK(cudaMemcpyAsync(host_rgba, s.d_rgba, (size_t)cfg->width * cfg->height * 4, cudaMemcpyDeviceToHost, c->copy_stream));
	CK(cudaEventRecord(s.done, c->copy_stream));
	s.busy = true;
	s.host_dst = host_rgba;
	c->next_ticket++;
	return ticket;
}

int rlerc_frame_wait(rlerc_ctx* c, int ticket)
{
	if (!c || ticket < 0 || ticket >= c->next_ticket) { set_error("rlerc_frame_wait: bad ticket"); return RLERC_ERR_ARG; }
	// a later submit that recycled this ticket's slot has already waited for it
	if (ticket < c->next_ticket - rlerc_ctx::kSlots) return RLERC_OK;
	FrameSlot& s = c->slot[ticket % rlerc_ctx::kSlots];
	int rc = set_dev(c);
	if (rc) return rc;
	if (s.busy) { CK(cudaEventSynchronize(s.done)); s.busy = false; }
	return RLERC_OK;
}

int rlerc_sync(rlerc_ctx* c)
{
	if (!c) return RLERC_ERR_ARG;
	int rc = set_dev(c);
	if (rc) return rc;
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaStreamSynchronize(c->copy_stream));
	return RLERC_OK;
}

void* rlerc_stream(rlerc_ctx* c) { return c ? (void*)c->stream : nullptr; }

int rlerc_memcpy_d2h(rlerc_ctx* c, void* host, const void* dev, size_t bytes)
{
	if (!c || !host || !dev) return RLERC_ERR_ARG;
	int rc = set_dev(c);
	if (rc) return rc;
	CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return RLERC_OK;
}

int rlerc_memcpy_h2d(rlerc_ctx* c, void* dev, const void* host, size_t bytes)
{
	if (!c || !host || !dev) return RLERC_ERR_ARG;
	int rc = set_dev(c);
	if (rc) return rc;
	CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return RLERC_OK;
}

int rlerc_host_alloc(void** p, size_t bytes)
{
	if (!p) return RLERC_ERR_ARG;
	CK(cudaMallocHost(p, bytes));
	return RLERC_OK;
}

void rlerc_host_free(void* p) { if (p) cudaFreeHost(p); }

int rlerc_last_kernel_ms(rlerc_ctx* c, float out[2])
{
	if (!c || !out) return RLERC_ERR_ARG;
	int rc = set_dev(c);
	if (rc) return rc;
	out[0] = out[1] = -1.0f;
	if (c->ev_valid[0]) { CK(cudaEventSynchronize(c->ev[1])); CK(cudaEventElapsedTime(&out[0], c->ev[0], c->ev[1])); }
	if (c->ev_valid[1]) { CK(cudaEventSynchronize(c->ev[3])); CK(cudaEventElapsedTime(&out[1], c->ev[2], c->ev[3])); }
	return RLERC_OK;
}

// ---- legacy surface (R/src/core.h:144-147, R/src/Cuda_Main.cu:124-148,183-271) ----------

int cpu_to_gpu_delta = 0;

static rlerc_ctx* g_legacy = nullptr;
static bool g_legacy_owned = false;
static rlerc_frame_config g_legacy_cfg;
static std::map<int, void*>* g_pbo = nullptr;

void* gpu_malloc(int size)
{
	void* p = nullptr;
	if (g_legacy) cudaSetDevice(g_legacy->device);
	if (cudaMalloc(&p, (size_t)size) != cudaSuccess) { set_error("gpu_malloc(%d) failed", size); return nullptr; }
	return p;
}
void gpu_memcpy(void* dst, void* src, int count) { cudaMemcpy(dst, src, (size_t)count, cudaMemcpyHostToDevice); }
void cpu_memcpy(void* dst, void* src, int count) { cudaMemcpy(dst, src, (size_t)count, cudaMemcpyDeviceToHost); }

int rlerc_legacy_init(int device, const rlerc_scene* scene, const rlerc_frame_config* cfg)
{
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if (g_legacy && g_legacy_owned) rlerc_destroy(g_legacy);
	g_legacy = nullptr;
	if ((rc = rlerc_create(device, &g_legacy))) return rc;
	g_legacy_owned = true;
	if ((rc = rlerc_scene_upload(g_legacy, scene))) return rc;
	g_legacy_cfg = *cfg;
	if (!g_pbo) g_pbo = new std::map<int, void*>();
	return RLERC_OK;
}

int rlerc_legacy_adopt(rlerc_ctx* c, const rlerc_frame_config* cfg)
{
	int rc = check_cfg(cfg);
	if (rc) return rc;
	if (!c) { set_error("rlerc_legacy_adopt: null context"); return RLERC_ERR_ARG; }
	if (g_legacy && g_legacy_owned && g_legacy != c) rlerc_destroy(g_legacy);
	g_legacy = c;
	g_legacy_owned = false;
	g_legacy_cfg = *cfg;
	if (!g_pbo) g_pbo = new std::map<int, void*>();
	return RLERC_OK;
}

int rlerc_pbo_bind(int pbo, void* device_ptr)
{
	if (!g_pbo) g_pbo = new std::map<int, void*>();
	(*g_pbo)[pbo] = device_ptr;
	return RLERC_OK;
}
void pboRegister(int pbo) { if (!g_pbo) g_pbo = new std::map<int, void*>(); if (!g_pbo->count(pbo)) (*g_pbo)[pbo] = nullptr; }
void pboUnregister(int pbo) { if (g_pbo) g_pbo->erase(pbo); }

// Same contract as the reference: renders map_line_count ray planes of `raymap` into the
// buffer behind `pbo_out` and returns once the kernel has finished (Cuda_Main.cu:241).
void cuda_main_render2(int pbo_out, int width, int height, rlerc_raymap* raymap)
{
	if (pbo_out == 0) return;                                            // Cuda_Main.cu:187
	if (!g_legacy || !g_pbo || !raymap)
	{
		// the reference would dereference garbage here; say what is missing instead of rendering nothing silently
		set_error("cuda_main_render2: no context (call RLE4::all_to_gpu / rlerc_legacy_init / rlerc_legacy_adopt first)");
		return;
	}
	auto it = g_pbo->find(pbo_out);
	if (it == g_pbo->end() || !it->second) { set_error("cuda_main_render2: pbo %d has no device buffer", pbo_out); return; }
	rlerc_frame_config cfg = g_legacy_cfg;
	cfg.render_size = width;
	(void)height;                                                        // res_x == res_y == RENDER_SIZE
	if (rlerc_render(g_legacy, raymap, &cfg, 0, -1, (uint32_t*)it->second) != RLERC_OK) return;
	cudaStreamSynchronize(g_legacy->stream);
}

} // extern "C"
