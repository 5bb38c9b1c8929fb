// Host-side check of include/rlerc.hpp (no CUDA call): the reference-named classes over the C ABI.
// usage: hpp_host_check OUTDIR  -> writes OUTDIR/scene.rle4, OUTDIR/raymap.bin (896 bytes), prints a summary line
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "rlerc.hpp"
using namespace rlerc;

int main(int argc, char** argv)
{
	if (argc < 2) return 2;
	const std::string dir = argv[1];
	try
	{
		static_assert(sizeof(Map4) == 32, "Map4 layout (R/src/Rle4.h:7-21, LP64)");
		static_assert(sizeof(RayMap_GPU) == 896, "RayMap_GPU layout (R/src/RayMap.h:16-54, LP64)");
		const int N = 64;
		std::vector<uint8_t> voxel((size_t)N * N * N / 8), c1(voxel.size()), c2(voxel.size());
		check(rlerc_synth_volume(0, N, N, N, 1, voxel.data(), c1.data(), c2.data()), "rlerc_synth_volume");
		RLE4 a;
		a.compress_all(voxel.data(), c1.data(), c2.data(), N, N, N);           // RLE4::compress_all
		a.save((dir + "/scene.rle4").c_str());                                  // RLE4::save
		RLE4 b;
		if (b.load((dir + "/missing.rle4").c_str())) return 3;                  // RLE4::load on a missing file: false, no throw
		if (!b.load((dir + "/scene.rle4").c_str())) return 4;                   // RLE4::load
		if (a.nummaps != b.nummaps) return 5;
		for (int m = 0; m < a.nummaps; m++)
		{
			const Map4 &x = a.map[m], &y = b.map[m];
			if (x.sx != y.sx || x.sy != y.sy || x.sz != y.sz || x.slabs_size != y.slabs_size) return 6;
			if (std::memcmp(x.slabs, y.slabs, (size_t)x.slabs_size * 2)) return 7;
			if (std::memcmp(x.map, y.map, (size_t)x.sx * x.sz * 8)) return 8;   // pointer map rebuilt by load == built by compress
		}
		RayMap ray_map(window(1024, 768));
		ray_map.nummaps = b.nummaps;                                            // must survive get_ray_map (main.cpp:277-278)
		ray_map.set_border(0.125f);                                             // main.cpp:774
		ray_map.set_ray_limit(4096);                                            // main.cpp:776
		ray_map.get_ray_map(vec3f(10000.0f, -818.0f, 10000.0f), vec3f(0.40f, (float)(0.30 + 1.57079632679489661923), 0.0f));
		if (ray_map.nummaps != b.nummaps) return 9;
		FILE* f = std::fopen((dir + "/raymap.bin").c_str(), "wb");
		if (!f) return 10;
		std::fwrite(static_cast<const RayMap_GPU*>(&ray_map), 1, sizeof(RayMap_GPU), f);
		std::fclose(f);
		b.clear();                                                              // RLE4::clear
		if (b.nummaps != 0 || b.map[0].slabs != nullptr) return 11;
		try { b.save((dir + "/x.rle4").c_str()); return 12; } catch (const Error& e) { if (e.status != RLERC_ERR_STATE) return 13; }
		std::printf("ok levels %d map_line_count %d\n", a.nummaps, ray_map.map_line_count);
	}
	catch (const Error& e)
	{
		std::fprintf(stderr, "hpp_host_check: %s\n", e.what());
		return 1;
	}
	return 0;
}
