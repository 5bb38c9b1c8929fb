/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Glue that compiles the reference's own host-side scene + frame-setup code
 * (SURVEY.md §8c "TU-B") in place from /root/reference (read-only, never copied):
 *     RLE-Raycaster/src/Rle4.cpp   RLE4::compress_all/compress/save/load   (:16-384)
 *     RLE-Raycaster/src/Tree.cpp + tree.h   Tree::init/sphere/cube/get_mipmap
 *     RLE-Raycaster/src/VecMath.cpp
 *     RLE-Raycaster/src/RayMap.h   RayMap::get_ray_map                     (:98-402)
 * Output: oracle/_ref/libref_host.so. Loaded only by tests/, smoke() and the
 * bench's cpu_baseline / --impl reference arm.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <vector>

/* core.h:64-74 would macro-define these; tree.h:736 uses uint(...) as a cast, which
 * only parses when uint is a type name */
typedef unsigned int uint;
typedef unsigned short ushort;
typedef unsigned char uchar;
#define uint uint
#define ushort ushort
#define uchar uchar
#include "Core.h"

static int g_ref_screen_size_x = 1024;
#undef SCREEN_SIZE_X
#define SCREEN_SIZE_X g_ref_screen_size_x

/* C helpers the reference's scene code links against (R/src/core.h:144-147, Rle4.h:4-5).
 * The oracle never touches a GPU: they are inert. */
extern "C" {
int   cpu_to_gpu_delta = 0;
void  cpu_memcpy(void*, void*, int) {}
void  gpu_memcpy(void*, void*, int) {}
void* gpu_malloc(int) { return 0; }
void  create_cuda_1d_texture(char*, int) {}
void  create_cuda_2d_texture(unsigned int*, int, int) {}
}

#include "Rle4.cpp"
#include "Tree.cpp"
#include "VecMath.cpp"
#include "RayMap.h"

RayMap ray_map;

struct RefScene { RLE4 rle4; };

extern "C" {

int ref_sizeof_raymap() { return (int)sizeof(RayMap_GPU); }
int ref_sizeof_map4() { return (int)sizeof(Map4); }

/* ---- scene: load/save/compress (Rle4.cpp) ---- */
void* ref_scene_load(const char* path)
{
	RefScene* s = new RefScene();
	s->rle4.init();
	if (!s->rle4.load((char*)path)) { delete s; return 0; }
	return s;
}
void ref_scene_save(void* h, const char* path) { ((RefScene*)h)->rle4.save((char*)path); }
void ref_scene_free(void* h) { RefScene* s = (RefScene*)h; s->rle4.clear(); delete s; }
int  ref_scene_nummaps(void* h) { return ((RefScene*)h)->rle4.nummaps; }
/* Map4 as laid out by the reference on this ABI (LP64: 32 bytes) */
const void* ref_scene_map4(void* h, int m) { return &((RefScene*)h)->rle4.map[m]; }

/* ---- Tree: bit-volume authoring (tree.h) ---- */
void* ref_tree_new(int sx, int sy, int sz, int usecolor)
{
	Tree* t = new Tree();
	t->init(sx, sy, sz, usecolor != 0);
	t->set_color(1);
	return t;
}
void  ref_tree_free(void* t) { ((Tree*)t)->exit(); delete (Tree*)t; }
void  ref_tree_set_color(void* t, int c) { ((Tree*)t)->set_color((char)c); }
void  ref_tree_sphere(void* t, float x, float y, float z, float r, int mode) { ((Tree*)t)->sphere(vec3f(x, y, z), r, mode); }
void  ref_tree_cube(void* t, float x0, float y0, float z0, float x1, float y1, float z1) { ((Tree*)t)->cube(vec3f(x0, y0, z0), vec3f(x1, y1, z1)); }
void* ref_tree_voxel(void* t) { return ((Tree*)t)->voxel; }
void* ref_tree_col1(void* t) { return ((Tree*)t)->voxel_col1; }
void* ref_tree_col2(void* t) { return ((Tree*)t)->voxel_col2; }

/* RLE4::compress_all (Rle4.cpp:16-50): mip pyramid + per-level compress.
 * NB: compress() emits a 1-uint-per-column map; load() rebuilds the 2-uint map.
 * Callers wanting the render-ready layout should save + load. */
void* ref_compress_all(void* t)
{
	RefScene* s = new RefScene();
	s->rle4.init();
	s->rle4.compress_all(*(Tree*)t);
	return s;
}

/* ---- frame setup (RayMap.h:98-402) ---- */
void ref_get_ray_map(const float pos[3], const float rot[3], float border, int rays_casted_res, void* out_raymap_gpu)
{
	g_ref_screen_size_x = rays_casted_res / 4; /* RAYS_CASTED_RES = SCREEN_SIZE_X*4 (core.h:7) */
	RayMap rm;
	memset((RayMap_GPU*)&rm, 0, sizeof(RayMap_GPU));
	rm.map_line_limit = 2500;
	rm.set_border(border);
	rm.set_ray_limit(rays_casted_res);
	rm.get_ray_map(vec3f(pos[0], pos[1], pos[2]), vec3f(rot[0], rot[1], rot[2]));
	memcpy(out_raymap_gpu, (RayMap_GPU*)&rm, sizeof(RayMap_GPU));
}

} /* extern "C" */
