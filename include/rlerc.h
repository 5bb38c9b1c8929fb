/* rlerc — B200-native RLE voxel raycaster frame loop: C ABI.
 *
 * Drop-in boundary for the render path of sp4cerat/RLE-based-Voxel-Raycasting
 * (SURVEY.md §8b).  Every entry point names the reference interface it replaces
 * ("R/" = RLE-Raycaster/ in the reference checkout).  Plain pointers and sizes
 * only; no C++/torch types.  All functions return RLERC_OK (0) or a negative
 * rlerc_status and never hang (the reference spins in while(1) on errors,
 * R/src/Cuda_Main.cu:146,193).  A context is bound to one CUDA device and is
 * thread-compatible (one caller thread per context).
 */
#ifndef RLERC_H
#define RLERC_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLERC_MAX_MAPS 16 /* R/src/RayMap.h:43  Map4 map4_gpu[16] */

typedef enum rlerc_status {
	RLERC_OK = 0,
	RLERC_ERR_ARG = -1,      /* null / out-of-range argument */
	RLERC_ERR_IO = -2,       /* file missing or truncated */
	RLERC_ERR_FORMAT = -3,   /* malformed .rle4 / non power-of-two grid / >32-bit offsets */
	RLERC_ERR_CUDA = -4,     /* CUDA runtime error (rlerc_last_error() has the text) */
	RLERC_ERR_STATE = -5,    /* call order (no scene uploaded, no frame rendered, ...) */
	RLERC_ERR_NOMEM = -6
} rlerc_status;

/* ---- POD mirrors of the reference structs (field order kept, LP64) ---------------- */

typedef struct rlerc_vec3f { float x, y, z; } rlerc_vec3f;          /* R/src/VecMath.h:15 / CUDA float3 */

/* R/src/Rle4.h:7-21 `struct Map4` (32 bytes on LP64).
 * map   : uint32[sx*sz*2], per column {slab_offset (ushort index), n_runs | first_run<<16}
 *         (layout built by RLE4::load, R/src/Rle4.cpp:297-303)
 * slabs : uint16[slabs_size], per column [n_runs][n_vox][run * n_runs][attr16 * n_vox],
 *         columns x-fastest then z.  run = solid(6b)<<10 | skip(10b). */
typedef struct rlerc_map4 {
	int32_t   sx, sy, sz, slabs_size;
	uint32_t* map;
	uint16_t* slabs;
} rlerc_map4;

/* R/src/RayMap.h:16-54 `struct RayMap_GPU` (896 bytes on LP64). */
typedef struct rlerc_raymap {
	rlerc_vec3f vanishing_point_2d;
	int32_t     map_line_count;
	int32_t     map_line_limit;
	rlerc_vec3f rotation;
	rlerc_vec3f position;
	float       border;
	float       clip_min, clip_max;
	rlerc_map4  map4_gpu[RLERC_MAX_MAPS];
	int32_t     nummaps;
	int32_t     maxres;
	int32_t     res[4];
	rlerc_vec3f p4;
	rlerc_vec3f p_2d[8];
	rlerc_vec3f p_no[8];
	float       to3d[4][4];
	float       p_ofs_min[4];
	float       p_ofs_max[4];
} rlerc_raymap;

/* The reference's compile-time configuration (R/src/core.h:3-10) as run-time fields. */
typedef struct rlerc_frame_config {
	int32_t width, height;      /* window: SCREEN_SIZE_X, SCREEN_SIZE_Y                  */
	int32_t render_size;        /* RENDER_SIZE: res_x = res_y of the warped ray buffer   */
	int32_t rays_casted;        /* RAYS_CASTED: rows of the warped ray buffer            */
	int32_t rays_casted_res;    /* RAYS_CASTED_RES: angular resolution (maxres*4)        */
	int32_t z_far;              /* RAYS_DISTANCE                                         */
	int32_t mip_distance;       /* MIP_DISTANCE                                          */
	float   border;             /* RayMap::set_border, R/src/main.cpp:774                */
	int32_t flags;              /* RLERC_FLAG_*: compile-time options of R/src/core.h as run-time switches (0 = shipped config) */
} rlerc_frame_config;

/* R/src/core.h:18 CLIPREGION: the scene is finite, columns outside the level-0 grid are skipped
 * (R/src/Cuda_Render.h:432-437) instead of wrapping (infinite tiling). */
#define RLERC_FLAG_CLIPREGION   1
/* R/src/core.h:22 HEIGHT_COLOR: the low attribute byte is scaled by the camera height (R/src/Cuda_Render.h:674-676,716-722). */
#define RLERC_FLAG_HEIGHT_COLOR 2
/* Both are implemented by the production traversal kernels (lanes_per_ray = 0, 65, 68, 69) only. */
/* R/src/core.h:12 ANTIALIAS: pass 1 is R/bin/shader/colorize_buddha_soft_2xAA.frag instead of colorize_buddha_soft.frag
 * (R/src/main.cpp:510-522): same texel geometry, its own shading (5:5:5 normal, point light, sky gradient).  rlerc_unwarp only. */
#define RLERC_FLAG_SHADER_2XAA  4

/* Defaults exactly as R/src/core.h for a W x H window: render_size=W, rays=4W,
 * z_far=80000, mip_distance=W, border=(1-H/W)/2 (=0.125 for 1024x768, main.cpp:774-776). */
void rlerc_frame_config_default(int width, int height, rlerc_frame_config* out);

typedef struct rlerc_ctx rlerc_ctx;       /* one CUDA device + streams + device scene replica */
typedef struct rlerc_scene rlerc_scene;   /* host-side RLE4 (R/src/Rle4.h:25-52)              */

const char* rlerc_last_error(void);
const char* rlerc_version(void);
/* Host worker threads of the scene tools (compressor, tiler, synth); n <= 0 only queries. Launchers such
 * as torchrun export OMP_NUM_THREADS=1, which would make scene construction single-threaded. */
int  rlerc_set_host_threads(int n);

/* ---- scene: replaces RLE4::load/save/clear/compress_all (R/src/Rle4.cpp:16-384) ---- */

/* RLE4::load (Rle4.cpp:244-384): reads the file and rebuilds the per-column pointer map. */
int  rlerc_scene_load(const char* path, rlerc_scene** out);
/* RLE4::save (Rle4.cpp:220-242). Fails with RLERC_ERR_FORMAT if a level exceeds int32 slabs. */
int  rlerc_scene_save(const rlerc_scene* s, const char* path);
/* Deep-copies `nummaps` levels laid out like RLE4::map[] after load(). */
int  rlerc_scene_from_maps(const rlerc_map4* maps, int nummaps, rlerc_scene** out);
/* RLE4::clear (Rle4.cpp:210-218). */
void rlerc_scene_free(rlerc_scene* s);
int  rlerc_scene_nummaps(const rlerc_scene* s);
/* Borrowed view of level m (host pointers, valid until rlerc_scene_free). */
int  rlerc_scene_level(const rlerc_scene* s, int m, rlerc_map4* out, uint64_t* slabs_size64);
/* RLE4::compress_all (Rle4.cpp:16-50) + Tree::get_mipmap (tree.h:23-85) on a bit volume
 * (x fastest, then y, then z; bit x&7 of byte (x+y*sx+z*sx*sy)>>3; col1/col2 = the two material bits in the same
 * layout, both may be NULL):
 * byte-identical output, multi-threaded. */
int  rlerc_scene_compress(const uint8_t* voxel, const uint8_t* col1, const uint8_t* col2,
                          int sx, int sy, int sz, rlerc_scene** out);
/* Physical nx x nz tiling of every level in x and z (BASELINE config 3). */
int  rlerc_scene_tile(const rlerc_scene* s, int nx, int nz, rlerc_scene** out);
/* Procedural bit volumes for the synthetic benchmark scenes.
 * kind 0: terrain + boulders + caves ("synth_imrodh"); kind 1: worst-case short-run band. */
int  rlerc_synth_volume(int kind, int sx, int sy, int sz, uint32_t seed,
                        uint8_t* voxel, uint8_t* col1, uint8_t* col2);
/* Heightfield + cave floors + worst-case short-run band (one column in band_every, 32 runs of one voxel) written
 * straight into the .rle4 layout, all mip levels: BASELINE config 4 at its full size (16384 x 1024 x 16384) without
 * a 32 GiB bit volume.  Sizes: powers of two, sy <= 1024. */
int  rlerc_synth_rle(int sx, int sy, int sz, uint32_t seed, int band_every, rlerc_scene** out);

/* ---- context / device scene: replaces gpu_malloc + RLE4::all_to_gpu ----------------- */

int  rlerc_create(int device, rlerc_ctx** out);
void rlerc_destroy(rlerc_ctx* c);
/* RLE4::all_to_gpu / copy_to_gpu (Rle4.cpp:432-448): full replica of every level in HBM. */
int  rlerc_scene_upload(rlerc_ctx* c, const rlerc_scene* s);
/* Borrow the replica another context of the SAME device holds (no copy; `src` has to outlive every use of `dst`).
 * Several contexts = several frames in flight on one GPU (each context has its own stream and frame buffers). */
int  rlerc_scene_share(rlerc_ctx* dst, const rlerc_ctx* src);
/* Device-side Map4 table as main.cpp:277-278 copies it into the ray map (all levels). */
int  rlerc_scene_device_maps(rlerc_ctx* c, rlerc_map4* out16, int* nummaps);
/* Traversal kernel variant, all bit-identical: 0 = automatic (default): the production kernel k_traverse_f (one warp
 * per ray plane, column filter in front of the occlusion machinery), or k_traverse_q (four warps per ray plane: filter /
 * project / resolve / shade) when the launch has at most 5 ray planes per SM, is rendered alone (no frames in flight)
 * and is therefore bound by the serial chain of its longest ray planes (multi-GPU slices one frame at a time, small
 * windows); 65 = k_traverse_f always; 69 = k_traverse_q always; 68 = k_traverse_p (two warps per ray plane) always
 * (all run the DDA pre-pass k_dda_states first); 1,2,4,8,16,32 = k_traverse<lanes> (lane <-> run; 1 is the reference's
 * thread-per-ray scheme).
 * Only in a library built with `make VARIANTS=1` (rlerc_has_variants() == 1; measured and rejected in round 1):
 * 64 = k_traverse_w (three-stage pipeline over all columns); 66, 67 = k_traverse_c (DDA in a dedicated warp). */
int  rlerc_set_lanes_per_ray(rlerc_ctx* c, int lanes);
int  rlerc_has_variants(void);
/* Name of the traversal kernel the last rlerc_render* call launched (static string). */
const char* rlerc_last_kernel(rlerc_ctx* c);
/* k_traverse_w only: run the DDA in dedicated producer blocks or inside every warp (default). */
int  rlerc_set_dda_producer(rlerc_ctx* c, int on);
/* k_traverse_w only: 0 = serial DDA recurrence in every warp (default), 2 = merge-path DDA, 3 = closed-form
 * lane-parallel DDA (csrc/dda_closed.cuh). Same results. */
int  rlerc_set_dda_mode(rlerc_ctx* c, int mode);

/* ---- frame setup: replaces RayMap::set_border/set_ray_limit/get_ray_map ------------ */

/* RayMap::get_ray_map (R/src/RayMap.h:98-402), host only, byte-exact. Does not touch
 * out->map4_gpu / out->nummaps. */
int  rlerc_frame_setup(const float pos[3], const float rot[3], const rlerc_frame_config* cfg,
                       rlerc_raymap* out);

/* ---- render: replaces cuda_main_render2 (R/src/Cuda_Main.cu:183-271) --------------- */

/* Traversal kernel over ray planes [ray_begin, ray_end) (whole frame: 0, -1) into the
 * warped ray buffer uint32[rays_casted][render_size] in DEVICE memory (d_warp == NULL:
 * the context's own buffer).  The ray map's map4_gpu/nummaps are ignored: the uploaded
 * replica is used.  Asynchronous on the context stream. */
int  rlerc_render(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                  int ray_begin, int ray_end, uint32_t* d_warp);
/* Same, additionally writing per-pixel hit identity for parity checks:
 * d_ids uint32[rays_casted][render_size][2] = {column index vx+vz*sx at the hit level,
 * mip_level<<16 | voxel index in the column's attribute array}. Debug build of the kernel. */
int  rlerc_render_ids(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                      int ray_begin, int ray_end, uint32_t* d_warp, uint32_t* d_ids);
/* Work counters of the last rlerc_render_ids call (DETAIL_BENCH, R/src/Cuda_Render.h:14-22
 * plus the byte-model terms of SURVEY.md §8d):
 * [0] elems_total [1] elems_processed [2] voxels_processed [3] elems_rendered [4] pixels
 * [5] map entries fetched C [6] run-loop iterations E [7] fetched columns with >=1 run C1
 * [8] cleared pixels K [9] DDA steps */
int  rlerc_render_counters(rlerc_ctx* c, uint64_t out[10]);

/* ---- unwarp + shade: replaces GLSL pass 1 (R/bin/shader/colorize_buddha_soft.frag,
 *      uniforms R/src/main.cpp:578-603) ------------------------------------------------ */

/* d_warp (NULL: context buffer) -> RGBA8 uint8[height][width][4] in DEVICE memory, row 0 =
 * TOP of the window (GL row height-1). rows [row_begin,row_end) only (whole frame: 0,-1). */
int  rlerc_unwarp(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                  const uint32_t* d_warp, uint8_t* d_rgba, int row_begin, int row_end);
/* Multi-GPU compositing helper: like rlerc_unwarp, but only pixels whose ray plane lies in
 * [ray_begin, ray_end) are written; all other pixels are set to 0, so that a sum-reduce of
 * the per-GPU images equals the single-GPU image (disjoint support). */
int  rlerc_unwarp_slice(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                        const uint32_t* d_warp, uint8_t* d_rgba, int ray_begin, int ray_end);

/* Depth-aware smoothing: replaces GLSL pass 2 (R/bin/shader/soft.frag:1-75, R/src/main.cpp:628-653).
 * d_rgba_in = rlerc_unwarp output (alpha = quantised 0.001/z), d_rgba_out = RGBA8 [height][width][4]; not in place.
 * The reference's 2048^2 FBO is enlarged to the next power of two that holds the window. */
int  rlerc_soft(rlerc_ctx* c, const rlerc_frame_config* cfg, const uint8_t* d_rgba_in, uint8_t* d_rgba_out);

/* Interleaved multi-GPU slices: ray plane r belongs to `rank` iff
 * (r / block) % nranks == rank.  render: traverses only the owned ray planes; unwarp: writes
 * only pixels whose ray plane is owned, 0 elsewhere (sum over ranks == single-GPU image). */
int  rlerc_render_interleaved(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                              int block, int nranks, int rank, uint32_t* d_warp);
int  rlerc_unwarp_interleaved(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                              int block, int nranks, int rank, const uint32_t* d_warp, uint8_t* d_rgba);

/* rlerc_render_interleaved + rlerc_unwarp_interleaved on the context's own warped buffer in one call (nranks == 1: the
 * whole frame); RGBA8 into d_rgba (DEVICE memory, NULL: the context's buffer).  Asynchronous on the context stream. */
int  rlerc_frame_device(rlerc_ctx* c, const rlerc_raymap* rm, const rlerc_frame_config* cfg,
                        int block, int nranks, int rank, uint8_t* d_rgba);

/* ---- whole frame with HOST buffers (what render_to_pbo + display_pbo pass 1 do) ----- */

/* get_ray_map -> traversal -> unwarp -> D2H into host_rgba (width*height*4 bytes; pinned
 * memory recommended).  Synchronous. out_raymap may be NULL. */
int  rlerc_render_frame(rlerc_ctx* c, const float pos[3], const float rot[3],
                        const rlerc_frame_config* cfg, uint8_t* host_rgba, rlerc_raymap* out_raymap);
/* Pipelined variant: up to `depth` frames in flight (device double buffering, D2H of frame n
 * overlaps traversal of frame n+1). submit returns a ticket; wait blocks until that frame's
 * pixels are in host_rgba. */
int  rlerc_frame_submit(rlerc_ctx* c, const float pos[3], const float rot[3],
                        const rlerc_frame_config* cfg, uint8_t* host_rgba);
int  rlerc_frame_wait(rlerc_ctx* c, int ticket);

/* ---- multi-GPU frames (the reference is single-GPU) -------------------- */
/* N GPUs, a full replica of the scene in each; every frame's ray planes are dealt to the GPUs in interleaved blocks of
 * `slice_block` (R/src/Cuda_Render.h:109: ray planes are independent); compositing happens inside the unwarp kernel over
 * NVLink peer memory: GPU r produces rows [r*H/N, (r+1)*H/N) of the window, loading every texel from the warped buffer
 * of the GPU that traversed its ray plane.  The multi-GPU frame equals the single-GPU frame bit for bit. */

/* (a) all GPUs in ONE process — what a C++ host (the reference's host language) uses in place of rlerc_create. */
typedef struct rlerc_multi rlerc_multi;
int  rlerc_create_multi(const int* devices, int n, rlerc_multi** out);                 /* n <= 8 distinct devices */
void rlerc_multi_destroy(rlerc_multi* m);
int  rlerc_multi_count(const rlerc_multi* m);
rlerc_ctx* rlerc_multi_ctx(rlerc_multi* m, int i);                                     /* member i's context (borrowed) */
/* frames in flight (default 8) and ray planes per interleaved block (default 32); takes effect at the next frame */
int  rlerc_multi_set_depth(rlerc_multi* m, int depth, int slice_block);
int  rlerc_multi_scene_upload(rlerc_multi* m, const rlerc_scene* s);                   /* RLE4::all_to_gpu on every GPU */
/* rlerc_render_frame / rlerc_frame_submit / rlerc_frame_wait on all GPUs: get_ray_map -> traversal slices -> unwarp bands
 * -> every GPU copies ITS rows into host_rgba over its own PCIe link (host_rgba: width*height*4 bytes, pinned recommended). */
int  rlerc_multi_render_frame(rlerc_multi* m, const float pos[3], const float rot[3],
                              const rlerc_frame_config* cfg, uint8_t* host_rgba);
int  rlerc_multi_frame_submit(rlerc_multi* m, const float pos[3], const float rot[3],
                              const rlerc_frame_config* cfg, uint8_t* host_rgba);
int  rlerc_multi_frame_wait(rlerc_multi* m, int ticket);

/* (b) one member per context, members in one process or ONE PROCESS PER GPU (torchrun): create, exchange the
 * RLERC_GROUP_BLOB_BYTES descriptions (CUDA IPC handles) by any means, connect, then submit the same frames in the same
 * order on every member. */
typedef struct rlerc_group rlerc_group;
#define RLERC_GROUP_BLOB_BYTES 384
int  rlerc_group_create(rlerc_ctx* c, int rank, int nranks, int depth, int slice_block,
                        const rlerc_frame_config* cfg, rlerc_group** out);
void rlerc_group_destroy(rlerc_group* g);
int  rlerc_group_export(rlerc_group* g, void* blob);
int  rlerc_group_connect(rlerc_group* g, const void* blobs /* nranks * RLERC_GROUP_BLOB_BYTES, member r at r * BYTES */);
/* Enqueue this member's part of the next frame (asynchronous; returns its ticket, the same number on every member).
 * dst_rank >= 0: the finished frame is assembled in member dst_rank's image (pushed there over NVLink);
 * dst_rank < 0: every member keeps its band of rows.  host_rgba (may be NULL) = start of the WHOLE frame in host memory:
 * the member that holds the frame (or every member its band) copies it there. */
int  rlerc_group_submit(rlerc_group* g, const rlerc_raymap* rm, int dst_rank, uint8_t* host_rgba);
/* dst_rank = RLERC_GROUP_DST_ROTATE: frame t is assembled on member t mod nranks (spreads the NVLink ingest over all GPUs) */
enum { RLERC_GROUP_DST_ROTATE = -2 };
/* View batches (BASELINE config 5: one camera per GPU, finished frames gathered over NVLink): every member renders the
 * WHOLE frame of its own camera and copies it into slot [rank] of member dst_rank's view array (peer mapping).
 * rlerc_group_enable_views (before the export / connect exchange) allocates that array on a member that is to receive. */
int  rlerc_group_enable_views(rlerc_group* g);
int  rlerc_group_submit_view(rlerc_group* g, const rlerc_raymap* rm, int dst_rank, uint8_t* host_rgba);
/* the ticket's nranks views on this member: view r at d_views + r * view_stride, [height][width][4] each */
int  rlerc_group_views(rlerc_group* g, int ticket, uint8_t** d_views, size_t* view_stride);
int  rlerc_group_wait(rlerc_group* g, int ticket);
int  rlerc_group_sync(rlerc_group* g);
/* Device image of the ticket's slot [height][width][4] and the rows this member produces. */
int  rlerc_group_image(rlerc_group* g, int ticket, uint8_t** d_rgba, int* row_begin, int* row_end);
void* rlerc_group_stream(rlerc_group* g, int ticket);                                  /* cudaStream_t of the ticket's slot */
/* with rlerc_set_timing on the member's context: ms from the first kernel of the ticket's frame to its last barrier */
int  rlerc_group_last_ms(rlerc_group* g, int ticket, float* ms);

/* ---- LOD streaming: mip-level residency by distance (the reference keeps everything resident,
 *      R/src/Rle4.cpp:432-438) ------------------------------------------------------------ */
/* Like rlerc_scene_upload, but only address space is reserved: no level data is resident until rlerc_stream_prepare asks
 * for it.  `s` is borrowed and has to outlive the context's use of it (chunks are uploaded from it on demand).  Kernels
 * and results are the same as with a full replica; a frame rendered without the prepare call for its camera fails with
 * RLERC_ERR_CUDA (a fault on unmapped memory), it never draws a wrong picture. */
int  rlerc_scene_upload_streamed(rlerc_ctx* c, const rlerc_scene* s);
typedef struct rlerc_stream_stats {
	uint64_t resident_bytes, total_bytes;    /* after the call: HBM in use for level data / size of a full replica */
	uint64_t uploaded_bytes, evicted_bytes;  /* this call */
	int32_t  chunks_mapped, chunks_unmapped;
} rlerc_stream_stats;
/* Make resident what a frame at this camera can touch: of every mip level the z-rows within the distance the traversal
 * covers while it reads that level (Cuda_Render.h:343-367), plus margin_voxels of look-ahead (level-0 voxels); rows
 * beyond twice the margin are evicted (after waiting for everything in flight on the device, if anything is evicted).
 * Synchronous.  stats may be NULL. */
int  rlerc_stream_prepare(rlerc_ctx* c, const float pos[3], const float rot[3], const rlerc_frame_config* cfg,
                          int margin_voxels, rlerc_stream_stats* stats);

/* ---- plumbing ------------------------------------------------------------------------ */
int   rlerc_sync(rlerc_ctx* c);
void* rlerc_stream(rlerc_ctx* c);                    /* cudaStream_t the kernels run on */
/* Run the kernels on a caller-owned cudaStream_t (e.g. the stream NCCL collectives are enqueued on). */
int   rlerc_set_stream(rlerc_ctx* c, void* cuda_stream);
int   rlerc_warp_buffer(rlerc_ctx* c, const rlerc_frame_config* cfg, uint32_t** d_warp);
int   rlerc_memcpy_d2h(rlerc_ctx* c, void* host, const void* dev, size_t bytes);
int   rlerc_memcpy_h2d(rlerc_ctx* c, void* dev, const void* host, size_t bytes);
int   rlerc_host_alloc(void** p, size_t bytes);      /* pinned */
void  rlerc_host_free(void* p);
/* Last kernel times in ms from CUDA events on the context stream: [0] traversal [1] unwarp */
int   rlerc_last_kernel_ms(rlerc_ctx* c, float out[2]);
/* Enables per-launch CUDA-event timing (adds two event records per kernel). */
int   rlerc_set_timing(rlerc_ctx* c, int on);

/* ---- legacy surface of the reference, same names and signatures ---------------------- */
/* R/src/core.h:144-147 */
void* gpu_malloc(int size);
void  gpu_memcpy(void* dst, void* src, int count);   /* host -> device */
void  cpu_memcpy(void* dst, void* src, int count);   /* device -> host */
extern int cpu_to_gpu_delta;                           /* always 0: no mirrored arena */
/* R/src/Cuda_Main.cu:124-126.  There is no GL here: a "pbo" is a small integer handle bound to
 * a device buffer of rays_casted*render_size*4 bytes by rlerc_pbo_bind(); pboRegister/
 * pboUnregister keep their signatures. raymap is the reference's RayMap_GPU (== rlerc_raymap).
 * Uses the process-default context created by rlerc_legacy_init(). */
int   rlerc_legacy_init(int device, const rlerc_scene* scene, const rlerc_frame_config* cfg);
/* ... or make an existing context (scene already uploaded) the one the legacy entry points use; it stays the caller's. */
int   rlerc_legacy_adopt(rlerc_ctx* c, const rlerc_frame_config* cfg);
int   rlerc_pbo_bind(int pbo, void* device_ptr);
void  pboRegister(int pbo);
void  pboUnregister(int pbo);
void  cuda_main_render2(int pbo_out, int width, int height, rlerc_raymap* raymap);

#ifdef __cplusplus
}
#endif
#endif /* RLERC_H */
