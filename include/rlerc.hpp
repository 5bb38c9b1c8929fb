// rlerc.hpp — the reference's host C++ surface for the render path, on top of the C ABI (rlerc.h).
//
// Header-only.  Same type names, member names and call order as the reference host code, so that the frame loop of
// R/src/main.cpp (R/ = RLE-Raycaster/ in the reference checkout) keeps compiling when the reference's Rle4.h /
// RayMap.h are swapped for this header:
//
//     RLE4 rle4;  rle4.load("Imrodh.rle4");  rle4.all_to_gpu();              // main.cpp:214-270
//     memcpy(ray_map.map4_gpu, rle4.mapgpu, rle4.nummaps * sizeof(Map4));     // main.cpp:277 (all levels, not 10)
//     ray_map.nummaps = rle4.nummaps;
//     ray_map.set_border(0.125f); ray_map.set_ray_limit(RAYS_CASTED_RES);     // main.cpp:774-776
//     ray_map.get_ray_map(pos, rot);                                          // main.cpp:777
//     rlerc_pbo_bind(pbo, device_buffer);                                     // no GL here: a "pbo" is a handle bound to device memory
//     cuda_main_render2(pbo, RENDER_SIZE, RENDER_SIZE, &ray_map);             // main.cpp:466 (renders with the context all_to_gpu filled)
//
//   reference                                     here
//   struct Map4          R/src/Rle4.h:7-21        rlerc::Map4   (= rlerc_map4, field order kept, LP64)
//   struct RLE4          R/src/Rle4.h:25-52       rlerc::RLE4   (load / save / clear / all_to_gpu / map[] / mapgpu[] / nummaps)
//   struct RayMap_GPU    R/src/RayMap.h:16-54     rlerc::RayMap_GPU (= rlerc_raymap)
//   class  RayMap        R/src/RayMap.h:57-418    rlerc::RayMap (set_border / set_ray_limit / get_ray_map)
//   vec3f                R/src/VecMath.h:15       rlerc::vec3f
//   compile-time core.h  R/src/core.h:3-10        rlerc::Config (run-time; Config::window(W, H) = the core.h formulas)
//
// Differences a porter has to know: errors throw rlerc::Error (the reference pops a MessageBox or spins in
// while(1)); there is no 100 MB mirrored arena (cpu_to_gpu_delta == 0); RLE4::compress_all takes a bit volume
// instead of the CSG `Tree`, which is out of scope.  `using namespace rlerc;` gives the bare names.
#ifndef RLERC_HPP
#define RLERC_HPP

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include "rlerc.h"

namespace rlerc {

struct Error : std::runtime_error {
	int status;
	Error(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};
inline void check(int rc, const char* call)
{
	if (rc != RLERC_OK) throw Error(rc, std::string(call) + " failed (" + std::to_string(rc) + "): " + rlerc_last_error());
}

typedef rlerc_map4 Map4;                 // R/src/Rle4.h:7-21
typedef rlerc_raymap RayMap_GPU;         // R/src/RayMap.h:16-54
typedef rlerc_frame_config Config;       // R/src/core.h:3-10 as run-time fields

struct vec3f {                           // R/src/VecMath.h:15
	float x, y, z;
	vec3f() : x(0), y(0), z(0) {}
	vec3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

inline Config window(int width, int height)          // SCREEN_SIZE_X/Y -> RENDER_SIZE, RAYS_CASTED, ... (core.h:3-10)
{
	Config c;
	rlerc_frame_config_default(width, height, &c);
	return c;
}

// The process-wide device context the legacy entry points use (the reference has exactly one device and one scene).
class Device {
public:
	static Device& get(int device = 0)
	{
		static Device d(device);
		return d;
	}
	rlerc_ctx* ctx() const { return ctx_; }
	~Device() { if (ctx_) rlerc_destroy(ctx_); }
private:
	explicit Device(int device) : ctx_(nullptr) { check(rlerc_create(device, &ctx_), "rlerc_create"); }
	Device(const Device&);
	Device& operator=(const Device&);
	rlerc_ctx* ctx_;
};

struct RLE4 {                            // R/src/Rle4.h:25-52
	Map4 map[RLERC_MAX_MAPS], mapgpu[RLERC_MAX_MAPS];
	int nummaps;

	RLE4() : nummaps(0), scene_(nullptr) { init(); }
	~RLE4() { clear(); }

	void init()                          // Rle4.cpp:200-208
	{
		std::memset(map, 0, sizeof(map));
		std::memset(mapgpu, 0, sizeof(mapgpu));
		nummaps = 0;
	}
	void clear()                         // Rle4.cpp:210-218
	{
		if (scene_) rlerc_scene_free(scene_);
		scene_ = nullptr;
		init();
	}
	bool load(const char* filename)      // Rle4.cpp:244-384; false when the file cannot be read (reference: MessageBox + false)
	{
		clear();
		const int rc = rlerc_scene_load(filename, &scene_);
		if (rc == RLERC_ERR_IO) { scene_ = nullptr; return false; }
		check(rc, "rlerc_scene_load");
		adopt();
		return true;
	}
	void save(const char* filename) const   // Rle4.cpp:220-242
	{
		need_scene();
		check(rlerc_scene_save(scene_, filename), "rlerc_scene_save");
	}
	// RLE4::compress_all (Rle4.cpp:16-50) on a bit volume (x fastest, then y, then z) instead of the CSG Tree.
	void compress_all(const uint8_t* voxel, const uint8_t* col1, const uint8_t* col2, int sx, int sy, int sz)
	{
		clear();
		check(rlerc_scene_compress(voxel, col1, col2, sx, sy, sz, &scene_), "rlerc_scene_compress");
		adopt();
	}
	// Takes ownership of a scene made by the C ABI (rlerc_synth_rle, rlerc_scene_tile, ...).
	void adopt(rlerc_scene* s)
	{
		clear();
		scene_ = s;
		adopt();
	}
	// Rle4.cpp:432-438: every level, full replica in HBM.  The context that holds it becomes the one the reference's own
	// entry point cuda_main_render2 renders with (cfg: the constants of R/src/core.h it was compiled with; default: as shipped).
	void all_to_gpu(int device = 0, const Config* cfg = nullptr)
	{
		need_scene();
		rlerc_ctx* c = Device::get(device).ctx();
		check(rlerc_scene_upload(c, scene_), "rlerc_scene_upload");
		int n = 0;
		check(rlerc_scene_device_maps(c, mapgpu, &n), "rlerc_scene_device_maps");
		const Config shipped = window(1024, 768);
		check(rlerc_legacy_adopt(c, cfg ? cfg : &shipped), "rlerc_legacy_adopt");
	}
	rlerc_scene* handle() const { return scene_; }

private:
	RLE4(const RLE4&);
	RLE4& operator=(const RLE4&);
	void need_scene() const { if (!scene_) throw Error(RLERC_ERR_STATE, "RLE4: no scene loaded"); }
	void adopt()
	{
		nummaps = rlerc_scene_nummaps(scene_);
		for (int m = 0; m < nummaps; m++)
		{
			uint64_t n64 = 0;
			check(rlerc_scene_level(scene_, m, &map[m], &n64), "rlerc_scene_level");
		}
	}
	rlerc_scene* scene_;
};

class RayMap : public RayMap_GPU {       // R/src/RayMap.h:57-418
public:
	RayMap()                             // RayMap.h:64
	{
		std::memset(static_cast<RayMap_GPU*>(this), 0, sizeof(RayMap_GPU));
		cfg_ = window(1024, 768);
		map_line_limit = 2500;
		set_border(0);
	}
	explicit RayMap(const Config& cfg)
	{
		std::memset(static_cast<RayMap_GPU*>(this), 0, sizeof(RayMap_GPU));
		cfg_ = cfg;
		map_line_limit = cfg.rays_casted_res;
		set_border(cfg.border);
	}
	void set_border(float a) { border = a; clip_min = border; clip_max = 1 - clip_min; cfg_.border = a; }   // RayMap.h:66
	void set_ray_limit(int a) { map_line_limit = a; cfg_.rays_casted_res = a; }                               // RayMap.h:68
	void get_ray_map(vec3f pos, vec3f rot)                                                                     // RayMap.h:98-402
	{
		// map4_gpu / nummaps (copied in by the caller, main.cpp:277-278) survive: rlerc_frame_setup does not touch them
		const float p[3] = { pos.x, pos.y, pos.z }, r[3] = { rot.x, rot.y, rot.z };
		check(rlerc_frame_setup(p, r, &cfg_, this), "rlerc_frame_setup");
	}
	const Config& config() const { return cfg_; }
	Config& config() { return cfg_; }

private:
	Config cfg_;
};

// One frame into host memory: compute_ray_map + render_to_pbo + display_pbo pass 1 (main.cpp:783-874, 466, 578-626).
inline void render_frame(vec3f pos, vec3f rot, const Config& cfg, uint8_t* host_rgba, RayMap_GPU* out = nullptr, int device = 0)
{
	const float p[3] = { pos.x, pos.y, pos.z }, r[3] = { rot.x, rot.y, rot.z };
	check(rlerc_render_frame(Device::get(device).ctx(), p, r, &cfg, host_rgba, out), "rlerc_render_frame");
}

} // namespace rlerc
#endif /* RLERC_HPP */
