// headless — the reference's frame loop without GL, written against the host C++ mirror (include/rlerc.hpp).
//
// What R/src/main.cpp does per frame (display(): compute_ray_map() :766-778 -> render_to_pbo() :441-470 ->
// display_pbo() pass 1 :578-626), minus the window: load (or synthesise) a scene, upload it, set up the ray map for
// a camera, run the traversal into the warped ray buffer, unwarp + shade, and dump
//     PREFIX.ppm        the final frame (P6, row 0 = top of the window)                     BASELINE config 1
//     PREFIX.warp.raw   the warped ray buffer, uint32[map_line_count][render_size]
//     PREFIX.txt        map_line_count, sizes, kernel times
// for golden comparison against the reference's raycast logic compiled for the host (tests/test_headless.py).
// There is no CPU fallback: without a CUDA device this exits with status 3 and the error text.
//
// usage: headless (--scene FILE.rle4 | --synth N) [--size W H] [--pos X Y Z] [--rot X Y Z] [--out PREFIX]
//                 [--frames K]   (K > 1: additionally times K frames of the scripted fly-through, pipelined)
//                 [--gpus N]     (N > 1: the same frame and fly-through on N GPUs through rlerc_create_multi — ray-plane slices per
//                                 GPU, bands of rows unwarped per GPU with the texels pulled over NVLink, every GPU copies its
//                                 band to the host; the multi-GPU frame has to equal the single-GPU frame byte for byte)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include "rlerc.hpp"

using namespace rlerc;

static void write_file(const std::string& path, const void* data, size_t bytes, const char* header = nullptr)
{
	FILE* f = std::fopen(path.c_str(), "wb");
	if (!f) throw Error(RLERC_ERR_IO, "cannot write " + path);
	if (header) std::fputs(header, f);
	if (bytes && std::fwrite(data, 1, bytes, f) != bytes) { std::fclose(f); throw Error(RLERC_ERR_IO, "short write " + path); }
	std::fclose(f);
}

int main(int argc, char** argv)
{
	std::string scene_path, out = "frame";
	int synth = 0, W = 1024, H = 768, frames = 1, gpus = 1;  // core.h:3-4 SCREEN_SIZE_X/Y
	vec3f pos(10000.0f, -818.0f, 10000.0f);                   // main.cpp:316-320,344-347
	vec3f rot(0.40f, (float)(0.30 + 1.57079632679489661923), 0.0f);
	bool pos_given = false;
	for (int i = 1; i < argc; i++)
	{
		const std::string a = argv[i];
		if (a == "--scene" && i + 1 < argc) scene_path = argv[++i];
		else if (a == "--synth" && i + 1 < argc) synth = std::atoi(argv[++i]);
		else if (a == "--size" && i + 2 < argc) { W = std::atoi(argv[++i]); H = std::atoi(argv[++i]); }
		else if (a == "--pos" && i + 3 < argc) { pos.x = (float)std::atof(argv[++i]); pos.y = (float)std::atof(argv[++i]); pos.z = (float)std::atof(argv[++i]); pos_given = true; }
		else if (a == "--rot" && i + 3 < argc) { rot.x = (float)std::atof(argv[++i]); rot.y = (float)std::atof(argv[++i]); rot.z = (float)std::atof(argv[++i]); }
		else if (a == "--out" && i + 1 < argc) out = argv[++i];
		else if (a == "--frames" && i + 1 < argc) frames = std::atoi(argv[++i]);
		else if (a == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
		else
		{
			std::fprintf(stderr, "usage: %s (--scene FILE.rle4 | --synth N) [--size W H] [--pos X Y Z] [--rot X Y Z] [--out PREFIX] [--frames K] [--gpus N]\n", argv[0]);
			return 2;
		}
	}
	if (scene_path.empty() && synth <= 0) { std::fprintf(stderr, "headless: give --scene FILE.rle4 or --synth N\n"); return 2; }
	try
	{
		// ---- init (main.cpp:214-278) -------------------------------------------------------------------------
		RLE4 rle4;
		if (!scene_path.empty())
		{
			if (!rle4.load(scene_path.c_str())) { std::fprintf(stderr, "headless: cannot read %s\n", scene_path.c_str()); return 1; }
		}
		else
		{
			const size_t n = (size_t)synth * synth * synth;
			std::vector<uint8_t> voxel(n / 8), col1(n / 8), col2(n / 8);   // bit volumes: solid, material bit 0, material bit 1
			check(rlerc_synth_volume(0, synth, synth, synth, 1, voxel.data(), col1.data(), col2.data()), "rlerc_synth_volume");
			rle4.compress_all(voxel.data(), col1.data(), col2.data(), synth, synth, synth);
			if (!pos_given) pos.y = -0.15f * (float)synth;         // above the synthetic terrain
		}
		std::printf("scene: %d levels, level 0 %d x %d x %d\n", rle4.nummaps, rle4.map[0].sx, rle4.map[0].sy, rle4.map[0].sz);
		rle4.all_to_gpu();                                        // main.cpp:270

		const Config cfg = window(W, H);
		RayMap ray_map(cfg);
		std::memcpy(ray_map.map4_gpu, rle4.mapgpu, rle4.nummaps * sizeof(Map4));   // main.cpp:277-278
		ray_map.nummaps = rle4.nummaps;

		// ---- one frame, staged like display() ----------------------------------------------------------------
		ray_map.set_border(cfg.border);                           // compute_ray_map(), main.cpp:766-778
		ray_map.set_ray_limit(cfg.rays_casted_res);
		ray_map.get_ray_map(pos, rot);

		rlerc_ctx* ctx = Device::get().ctx();
		check(rlerc_set_timing(ctx, 1), "rlerc_set_timing");
		uint32_t* d_warp = nullptr;
		check(rlerc_warp_buffer(ctx, &cfg, &d_warp), "rlerc_warp_buffer");
		check(rlerc_render(ctx, &ray_map, &cfg, 0, -1, nullptr), "rlerc_render");        // render_to_pbo(), main.cpp:466
		uint8_t* rgba = nullptr;
		check(rlerc_host_alloc((void**)&rgba, (size_t)W * H * 4), "rlerc_host_alloc");
		{
			// display_pbo() pass 1: unwarp + shade on the device, then read the frame back
			void* d_rgba = gpu_malloc(W * H * 4);
			if (!d_rgba) throw Error(RLERC_ERR_NOMEM, "gpu_malloc");
			check(rlerc_unwarp(ctx, &ray_map, &cfg, nullptr, (uint8_t*)d_rgba, 0, -1), "rlerc_unwarp");
			check(rlerc_sync(ctx), "rlerc_sync");
			check(rlerc_memcpy_d2h(ctx, rgba, d_rgba, (size_t)W * H * 4), "rlerc_memcpy_d2h");
		}
		float ms[2] = { 0, 0 };
		check(rlerc_last_kernel_ms(ctx, ms), "rlerc_last_kernel_ms");

		const int lines = ray_map.map_line_count < cfg.rays_casted ? ray_map.map_line_count : cfg.rays_casted;
		std::vector<uint32_t> warp((size_t)lines * cfg.render_size);
		check(rlerc_memcpy_d2h(ctx, warp.data(), d_warp, warp.size() * 4), "rlerc_memcpy_d2h");
		write_file(out + ".warp.raw", warp.data(), warp.size() * 4);

		std::vector<uint8_t> rgb((size_t)W * H * 3);
		for (size_t p = 0; p < (size_t)W * H; p++) { rgb[3 * p] = rgba[4 * p]; rgb[3 * p + 1] = rgba[4 * p + 1]; rgb[3 * p + 2] = rgba[4 * p + 2]; }
		char hdr[64];
		std::snprintf(hdr, sizeof(hdr), "P6\n%d %d\n255\n", W, H);
		write_file(out + ".ppm", rgb.data(), rgb.size(), hdr);

		char txt[512];
		std::snprintf(txt, sizeof(txt), "map_line_count %d\nrender_size %d\nrays_casted %d\nwidth %d\nheight %d\ntraverse_ms %.4f\nunwarp_ms %.4f\n",
		              ray_map.map_line_count, cfg.render_size, cfg.rays_casted, W, H, ms[0], ms[1]);
		write_file(out + ".txt", txt, std::strlen(txt));
		std::printf("%s.ppm %dx%d, %d ray planes, traversal %.3f ms, unwarp %.3f ms\n", out.c_str(), W, H, ray_map.map_line_count, ms[0], ms[1]);

		// ---- K frames of the scripted fly-through, four frames in flight ----
		if (frames > 1)
		{
			check(rlerc_set_timing(ctx, 0), "rlerc_set_timing");
			const int DEPTH = 4;
			uint8_t* pin[DEPTH];
			for (int k = 0; k < DEPTH; k++) check(rlerc_host_alloc((void**)&pin[k], (size_t)W * H * 4), "rlerc_host_alloc");
			const float sy = (float)rle4.map[0].sy;
			std::vector<int> tickets;
			const auto t0 = std::chrono::steady_clock::now();
			for (int i = 0; i < frames; i++)
			{
				const float a = 6.28318530717958647692f * (float)i / (float)frames;
				const float p[3] = { 10000.0f + 4000.0f * std::sin(a), scene_path.empty() ? -(0.15f + 0.08f * std::sin(2 * a)) * sy : -818.0f + 300.0f * std::sin(2 * a),
				                     10000.0f + 4000.0f * std::cos(a) };
				const float r[3] = { 0.35f + 0.3f * std::sin(3 * a), a + 1.57079632679489661923f, 0.0f };
				if ((int)tickets.size() >= DEPTH) { check(rlerc_frame_wait(ctx, tickets.front()), "rlerc_frame_wait"); tickets.erase(tickets.begin()); }
				const int t = rlerc_frame_submit(ctx, p, r, &cfg, pin[i % DEPTH]);
				if (t < 0) check(t, "rlerc_frame_submit");
				tickets.push_back(t);
			}
			for (size_t k = 0; k < tickets.size(); k++) check(rlerc_frame_wait(ctx, tickets[k]), "rlerc_frame_wait");
			check(rlerc_sync(ctx), "rlerc_sync");
			const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			std::printf("fly-through: %d frames in %.3f s = %.1f frames/s, %.1f Mrays/s (host buffers, %d frames in flight)\n",
			            frames, s, frames / s, (double)W * H * frames / s / 1e6, DEPTH);
			for (int k = 0; k < DEPTH; k++) rlerc_host_free(pin[k]);
		}
		// ---- the same on N GPUs behind one object (rlerc_create_multi) ------------------------------------------
		if (gpus > 1)
		{
			std::vector<int> devices(gpus);
			for (int g = 0; g < gpus; g++) devices[g] = g;
			rlerc_multi* multi = nullptr;
			check(rlerc_create_multi(devices.data(), gpus, &multi), "rlerc_create_multi");
			check(rlerc_multi_set_depth(multi, 2 * gpus < 8 ? 8 : 2 * gpus, 32), "rlerc_multi_set_depth");
			check(rlerc_multi_scene_upload(multi, rle4.handle()), "rlerc_multi_scene_upload");      // a full replica per GPU
			uint8_t* mrgba = nullptr;
			check(rlerc_host_alloc((void**)&mrgba, (size_t)W * H * 4), "rlerc_host_alloc");
			const float p0[3] = { pos.x, pos.y, pos.z }, r0[3] = { rot.x, rot.y, rot.z };
			check(rlerc_multi_render_frame(multi, p0, r0, &cfg, mrgba), "rlerc_multi_render_frame");
			const bool same = std::memcmp(mrgba, rgba, (size_t)W * H * 4) == 0;
			std::printf("%d GPUs: frame %s the single-GPU frame\n", gpus, same ? "is byte-identical to" : "DIFFERS from");
			if (frames > 1)
			{
				const int DEPTH = 2 * gpus < 8 ? 8 : 2 * gpus;
				std::vector<uint8_t*> pin(DEPTH);
				for (int k = 0; k < DEPTH; k++) check(rlerc_host_alloc((void**)&pin[k], (size_t)W * H * 4), "rlerc_host_alloc");
				const float sy = (float)rle4.map[0].sy;
				std::vector<int> tickets;
				const auto t0 = std::chrono::steady_clock::now();
				for (int i = 0; i < frames; i++)
				{
					const float a = 6.28318530717958647692f * (float)i / (float)frames;
					const float p[3] = { 10000.0f + 4000.0f * std::sin(a), scene_path.empty() ? -(0.15f + 0.08f * std::sin(2 * a)) * sy : -818.0f + 300.0f * std::sin(2 * a),
					                     10000.0f + 4000.0f * std::cos(a) };
					const float r[3] = { 0.35f + 0.3f * std::sin(3 * a), a + 1.57079632679489661923f, 0.0f };
					if ((int)tickets.size() >= DEPTH) { check(rlerc_multi_frame_wait(multi, tickets.front()), "rlerc_multi_frame_wait"); tickets.erase(tickets.begin()); }
					const int t = rlerc_multi_frame_submit(multi, p, r, &cfg, pin[i % DEPTH]);
					if (t < 0) check(t, "rlerc_multi_frame_submit");
					tickets.push_back(t);
				}
				for (size_t k = 0; k < tickets.size(); k++) check(rlerc_multi_frame_wait(multi, tickets[k]), "rlerc_multi_frame_wait");
				const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
				std::printf("fly-through on %d GPUs: %d frames in %.3f s = %.1f frames/s, %.1f Mrays/s (host buffers, %d frames in flight)\n",
				            gpus, frames, s, frames / s, (double)W * H * frames / s / 1e6, DEPTH);
				for (int k = 0; k < DEPTH; k++) rlerc_host_free(pin[k]);
			}
			rlerc_host_free(mrgba);
			rlerc_multi_destroy(multi);
			if (!same) { rlerc_host_free(rgba); return 4; }
		}
		rlerc_host_free(rgba);
	}
	catch (const Error& e)
	{
		std::fprintf(stderr, "headless: %s\n", e.what());
		return e.status == RLERC_ERR_CUDA ? 3 : 1;
	}
	return 0;
}
