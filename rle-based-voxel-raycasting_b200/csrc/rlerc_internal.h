// Internal declarations shared by the host C++ and the CUDA translation units.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "rlerc.h"

namespace rlerc {

// One mip level of a scene in host memory (reference: struct Map4, R/src/Rle4.h:7-21).
struct Level {
	int sx = 0, sy = 0, sz = 0;
	std::vector<uint32_t> map;    // 2 words per column, see rlerc_map4
	std::vector<uint16_t> slabs;
	uint64_t gather_pad = 0;      // ushorts the device copy needs behind the stream (validate_columns; 0 for scenes built here)
};

void set_error(const char* fmt, ...);

// scene.cpp
int build_pointer_map(Level& lv);   // RLE4::load's scan (R/src/Rle4.cpp:284-314)
int validate_columns(Level& lv);    // external data: every column inside the stream; sizes the padding the attribute gathers may need

// raymap.cpp
void get_ray_map(const float pos[3], const float rot[3], float border, int rays_casted_res, rlerc_raymap* out);

} // namespace rlerc

struct rlerc_scene {
	std::vector<rlerc::Level> levels;
};
