// k_traverse_c — k_traverse_f (column filter + occlusion machinery, one warp per ray plane) with the DDA taken
// out of the ray-plane warps.  Replaces cudaRender + Render::render_line (R/src/Cuda_Main.cu:150-181,
// R/src/Cuda_Render.h:96-737); same warped ray buffer bit for bit.
//
// Why: the DDA of a ray plane is a serial, warp-UNIFORM float recurrence (Cuda_Render.h:286-300,398-414); executed
// inside the ray-plane warp all 32 lanes compute the same thing, and once the filter had removed the dead columns
// it was 48 % of all instructions the kernel issued (ncu, profiles/).  Here a block is WPB ray-plane warps plus ONE
// DDA warp whose lanes 0..WPB-1 run the recurrences of the block's ray planes side by side — lane <-> ray plane:
// the instruction stream that used to serve one ray plane now serves WPB.  It writes CHUNKS of 128 crossing
// records into a two-slot ring per ray plane in shared memory and runs ahead of the consumers; head[r] / tail[r]
// (chunks published / released) are the only synchronisation, a lane produces whenever its ray plane has a free
// slot, a ray-plane warp waits only if its next chunk is not there yet.  No block-wide barrier after set-up, so
// ray planes of one block do not wait for each other.
//
// The filter / consume stages are those of k_traverse_f (see there); because no DDA runs between a pointer-map
// gather and its use any more, the filter keeps TWO batches in flight (alternating register sets).
#include <stdint.h>
#include <limits.h>
#include "kernels.cuh"
#include "device_common.cuh"
#include "traverse_common.cuh"

namespace rlerc {

#define RLERC_CH 128                        // crossings per chunk
#define RLERC_CQCAP 64                      // live-column queue capacity (ring, power of two)
#define RLERC_C_QUEUE (8 * RLERC_CQCAP)     // words: 8 fields x QCAP, field-major
#define RLERC_C_STATE 24                    // words of DDA state per ray plane in shared memory
#define RLERC_C_SPIN (1 << 21)              // polling limit: a protocol failure must not hang the GPU

// DDA state of one ray plane as the duty lane keeps it (shared memory between chunks)
struct DdaC {
	float d0, x0, y0;        // x-track: dds_dist0, isect0           (Cuda_Render.h:286-300)
	float nd1, x1, y1;       // z-track: -dds_dist1, isect1 (negated distance: the record marks the z-track by its sign)
	float gd0, gx0, gy0;     // grad_dist0, grad0
	float ngd1, gx1, gy1;    // -grad_dist1, grad1
	int mip, zi, dzi, mapswitch;   // z and dz are integer valued
	float csd, cpx, cpy;     // record of the last crossing made
	int done;                // z_far reached (Cuda_Render.h:366-367) or ray plane closed
};
static_assert(sizeof(DdaC) <= RLERC_C_STATE * 4, "DdaC must fit its shared-memory slot");

// One chunk for the ray planes of the DDA warp's lanes: up to RLERC_CH crossings each; out[0] = record before the
// chunk, out[1 + s] = record of crossing s {signed distance, pos.x, pos.y, mip}.  The inner loop carries no LOD /
// z_far test: it runs for the smallest budget of the participating lanes (crossings until z > mapswitch or
// z + dz > z_far; dz is a power of two), warp-uniform.  Lanes with active = false only take part in the votes.
// Returns the crossings this lane made.
__device__ __forceinline__ int ddac_chunk(DdaC& Q, float4* out, int last_map, int zfar_i, bool active)
{
	const unsigned FULL = 0xffffffffu;
	if (active) out[0] = make_float4(Q.csd, Q.cpx, Q.cpy, 0.0f);
	bool run = active && !Q.done;
	int s = 0;
	while (__any_sync(FULL, run))
	{
		int budget = RLERC_CH;
		if (run)
		{
			while (Q.zi > Q.mapswitch)                               // Cuda_Render.h:343-365
			{
				if (Q.mip < last_map) Q.mip++;
				Q.gx0 *= 2; Q.gy0 *= 2; Q.gx1 *= 2; Q.gy1 *= 2; Q.gd0 *= 2; Q.ngd1 *= 2;
				Q.mapswitch *= 2; Q.dzi *= 2;
			}
			const int sh = 31 - __clz(Q.dzi);
			const int lod_free = ((Q.mapswitch - Q.zi) >> sh) + 1;    // crossings before z > mapswitch
			const int far_free = (zfar_i - Q.zi) >> sh;               // crossings with z + dz <= z_far (Cuda_Render.h:366-367)
			if (far_free <= 0) { Q.done = 1; run = false; }
			else
			{
				budget = RLERC_CH - s;
				budget = budget < lod_free ? budget : lod_free;
				budget = budget < far_free ? budget : far_free;
			}
		}
		const int n = __reduce_min_sync(FULL, budget);               // >= 1 for every running lane
		if (run)
		{
			const float mipf = __int_as_float(Q.mip);
			float4* o = out + 1 + s;
			#pragma unroll 4
			for (int j = 0; j < n; j++)
			{
				const bool t1 = -Q.nd1 < Q.d0;                       // Cuda_Render.h:398-414
				o[j] = make_float4(t1 ? Q.nd1 : Q.d0, t1 ? Q.x1 : Q.x0, t1 ? Q.y1 : Q.y0, mipf);
				if (t1) { Q.nd1 += Q.ngd1; Q.x1 += Q.gx1; Q.y1 += Q.gy1; }
				else    { Q.d0 += Q.gd0; Q.x0 += Q.gx0; Q.y0 += Q.gy0; }
			}
			s += n;
			Q.zi += n << (31 - __clz(Q.dzi));
			if (s >= RLERC_CH) run = false;
		}
	}
	if (active && s > 0)
	{
		const float4 last = out[s];
		Q.csd = last.x; Q.cpx = last.y; Q.cpy = last.z;
	}
	return s;
}

// registers of one filter batch whose pointer-map gather is in flight
struct FilterSet {
	Geo g;
	unsigned e0, e1;
	bool have;
	int n;                   // crossings in the batch (0: nothing in flight)
};

__device__ __forceinline__ int ld_vol(const volatile int* p) { return *p; }

template <bool IDS, int WPB>
__global__ void __launch_bounds__((WPB + 1) * 32, 16 / (WPB + 1))
k_traverse_c(const __grid_constant__ TraverseParams P, int rays)
{
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int G = 32;
	const int gl = threadIdx.x & 31;
	const int wid = threadIdx.x >> 5;
	const unsigned FULL = 0xffffffffu;
	const unsigned lt_mask = (1u << gl) - 1u;

	// shared per block: crossing-record ring [2][WPB][CH + 1] | DDA states [WPB] | chunk sizes [2][WPB] | stop, head, tail [WPB]
	// shared per warp:  live-column queue | DrawJob | RW x 32 projected runs (int2) | RW x 32 deferred short spans | occlusion bits
	float4* ring = reinterpret_cast<float4*>(smem);
	uint32_t* after_ring = smem + 2 * WPB * (RLERC_CH + 1) * 4;
	DdaC* ddast = reinterpret_cast<DdaC*>(after_ring);
	int* ringn = reinterpret_cast<int*>(after_ring + WPB * RLERC_C_STATE);
	volatile int* stop = reinterpret_cast<volatile int*>(after_ring + WPB * RLERC_C_STATE + 2 * WPB);
	volatile int* head = stop + WPB;                                 // chunks published by the DDA warp
	volatile int* tail = stop + 2 * WPB;                             // chunks released by the ray-plane warp
	uint32_t* warps0 = after_ring + WPB * RLERC_C_STATE + 8 * WPB;
	const bool dda_warp = wid == WPB;
	const int per_warp = (RLERC_C_QUEUE + 16 + RLERC_PS_WORDS + P.mask_words + 3) & ~3;
	uint32_t* wbase = warps0 + (size_t)(dda_warp ? 0 : wid) * per_warp;
	uint32_t* queue = wbase;                                         // [8][QCAP]
	DrawJob* job = reinterpret_cast<DrawJob*>(wbase + RLERC_C_QUEUE);
	int2* proj = reinterpret_cast<int2*>(wbase + RLERC_C_QUEUE + 16);
	uint32_t* shade = wbase + RLERC_C_QUEUE + 16 + RLERC_RW * 64;
	uint32_t* ymask = wbase + RLERC_C_QUEUE + 16 + RLERC_PS_WORDS;

	const int ray_i = (int)blockIdx.x * WPB + wid;                  // launch-local ray index
	const int x = owned_ray(P, ray_i);
	const int res_y = P.res_y;
	const float res_y2 = (float)(res_y / 2);             // Cuda_Render.h:108 (integer division)
	const int zfar_i = P.z_far;
	const int last_map = P.nummaps - 1;

	// A warp without a ray plane (grid tail, or one of the early returns of Cuda_Render.h:226,250) still takes its
	// DDA duty turns and its barriers.
	bool finished = dda_warp || (ray_i >= rays) || (x >= P.ray_end);
	uint32_t* row = P.warp + (size_t)(finished ? 0 : x) * res_y;
	RayInit ri;
	ri.ray_x = ri.ray_z = ri.rx2mr = 0; ri.ycmin = ri.ycmax = 0; ri.vertical = false; ri.skip = true;
	if (!finished)
	{
		ray_init(P, x, ri);
		clear_outside<G>(row, res_y, ri, gl);
		if (ri.skip) finished = true;
	}
	const float ray_x = ri.ray_x, ray_z = ri.ray_z, rx2mr = ri.rx2mr;
	const bool vertical = ri.vertical;
	const float sin_x = P.sin_x, cos_x = P.cos_x;
	HorizonState Hs;
	Hs.ycmin = ri.ycmin; Hs.ycmax = ri.ycmax; Hs.hiw = 0;
	const int ymin0 = Hs.ycmin, ymax0 = Hs.ycmax;
	const bool has_row = !finished;

	if (!dda_warp) for (int w = gl; w < P.mask_words; w += G) ymask[w] = 0;

	const float vpx = P.viewpos[0], mountain = P.viewpos[1], vpz = P.viewpos[2];
	int fixx = 0, fixz = 0;
	{
		// DDA set-up of my ray plane (Cuda_Render.h:270-305), parked in shared memory for whoever is on duty
		DdaC Q;
		memset(&Q, 0, sizeof(Q));
		Q.done = 1;
		if (!finished)
		{
			Dda dd;
			dda_init(P, ray_x, ray_z, dd);
			fixx = dd.fixx; fixz = dd.fixz;
			Q.d0 = dd.d0; Q.x0 = dd.i0x; Q.y0 = dd.i0y; Q.nd1 = -dd.d1; Q.x1 = dd.i1x; Q.y1 = dd.i1y;
			Q.gd0 = dd.gd0; Q.gx0 = dd.g0x; Q.gy0 = dd.g0y; Q.ngd1 = -dd.gd1; Q.gx1 = dd.g1x; Q.gy1 = dd.g1y;
			Q.mip = 0; Q.zi = 0; Q.dzi = 1;                          // z and dz (Cuda_Render.h:181,325)
			Q.mapswitch = P.mapswitch0;
			Q.csd = 0.0f; Q.cpx = 0.0f; Q.cpy = 0.0f;                // no crossing yet: distance 0, x-track
			Q.done = 0;
			// The y_map_switch half of the LOD loop condition (Cuda_Render.h:343) can only be true on the first
			// crossing (it halves until <= 512 and never grows), where z = 0 < mapswitch.
			for (float yms = mountain; yms > 512.0f; yms = yms * 0.5f)
			{
				if (Q.mip < last_map) Q.mip++;
				Q.gx0 *= 2; Q.gy0 *= 2; Q.gx1 *= 2; Q.gy1 *= 2; Q.gd0 *= 2; Q.ngd1 *= 2;
				Q.mapswitch *= 2; Q.dzi *= 2;
			}
		}
		if (gl == 0 && !dda_warp) { ddast[wid] = Q; stop[wid] = finished ? 1 : 0; head[wid] = 0; tail[wid] = 0; }
	}
	const float pz_add = sin_x;                                  // pos3d_z_add (Cuda_Render.h:313)
	const float py_add = (vertical ? cos_x : 0.0f) * rx2mr;      // pos3d_y_add (Cuda_Render.h:314-315)

	Counters Cn;
	memset(&Cn, 0, sizeof(Cn));

	RayCtx R;
	R.row = row; R.ymask = ymask; R.ids = (IDS && has_row) ? P.ids + (size_t)x * res_y * 2 : nullptr;
	R.res_y2 = res_y2; R.pz_add = pz_add; R.py_add = py_add; R.mountain = mountain; R.gl = gl;
	R.stat = nullptr; R.hc_on = 0; R.hc = 0;

	FilterSet fa, fb;                  // two filter batches in flight
	fa.g.pz = fa.g.py = fa.g.czz = fa.g.cyy = 0; fa.g.cmip = 0; fa.g.cidx = 0;
	fa.e0 = fa.e1 = 0; fa.have = false; fa.n = 0;
	fb = fa;
	int fsteps = 0;                    // filter steps made: even -> set a, odd -> set b
	int qhead = 0, qcount = 0;         // live-column queue (uniform)
	Stage s0;                          // consume: the batch whose run words are in flight
	Geo g0 = fa.g;
	s0.nvalid = 0; s0.have = false; s0.e0 = s0.e1 = 0;
	#pragma unroll
	for (int k = 0; k < 4; k++) s0.rw[k] = 0;

	// F2. first-run test of a batch whose entries have arrived; live columns -> queue
	auto filter_test = [&](FilterSet& f)
	{
		if (f.n <= 0) return;
		const int ycmin = Hs.ycmin;
		bool live = false;
		if (gl < f.n && f.have)
		{
			const int slen = (int)(f.e1 & 0xffffu);
			const unsigned first = f.e1 >> 16;
			const int solid = (int)(first >> 10), skip = (int)(first & 1023u);
			if (IDS) live = true;                                // the byte model counts no-op columns too
			else if (slen == 0) live = false;                    // empty column: the run loop does not execute
			else if (solid == 0) live = true;                    // pure skip run: undecided, let the machinery look
			else
			{
				const float ft = (float)(skip << f.g.cmip);        // Cuda_Render.h:529-543 for run 0
				float zz1 = f.g.pz, yy1 = f.g.py;
				if (mountain + ft >= 0) { zz1 += f.g.czz; yy1 += f.g.cyy; }
				const float z1 = zz1 + pz_add * ft;
				if (z1 <= 0) live = true;                          // `continue`: a later run may be the first visible one
				else
				{
					const float y1 = yy1 + py_add * ft;
					live = f2i(res_y2 + y1 / z1) > ycmin;          // else: break, now and under every later horizon
				}
			}
		}
		const unsigned lb = __ballot_sync(FULL, live);
		if (live)
		{
			uint32_t* q = queue + ((qhead + qcount + __popc(lb & lt_mask)) & (RLERC_CQCAP - 1));
			q[0 * RLERC_CQCAP] = __float_as_uint(f.g.pz); q[1 * RLERC_CQCAP] = __float_as_uint(f.g.py);
			q[2 * RLERC_CQCAP] = __float_as_uint(f.g.czz); q[3 * RLERC_CQCAP] = __float_as_uint(f.g.cyy);
			q[4 * RLERC_CQCAP] = (uint32_t)f.g.cmip; q[5 * RLERC_CQCAP] = (uint32_t)f.g.cidx;
			q[6 * RLERC_CQCAP] = f.e0; q[7 * RLERC_CQCAP] = f.e1;
		}
		qcount += __popc(lb);
		f.n = 0;
		__syncwarp();
	};

	// F3. geometry of `nvalid` new crossings (records rec[0..nvalid]), conservative top clip, pointer-map gather
	auto filter_issue = [&](FilterSet& f, const float4* rec, int nvalid)
	{
		f.n = nvalid;
		f.have = false;
		if (gl < nvalid)
		{
			const float4 ra = rec[gl], rb = rec[gl + 1];           // state before / after crossing gl
			const float db = fabsf(ra.x), dn = fabsf(rb.x);
			const int ib = __float_as_int(ra.x) < 0 ? 1 : 0;        // index_before: sign bit of the record
			f.g.cmip = __float_as_int(rb.w);
			const int fix_x = (1 - ib) * fixx, fix_z = ib * fixz;    // Cuda_Render.h:418-419
			const float ddelta = dn - db;
			const float vsx = ray_x * db, vsz = ray_z * db;
			const int voxel_x = f2i(vpx + ra.y) + fix_x;             // Cuda_Render.h:429-430
			const int voxel_z = f2i(vpz + ra.z) + fix_z;
			const int gx = P.level[f.g.cmip].sx, gz = P.level[f.g.cmip].sz;
			const int vx = (voxel_x >> f.g.cmip) & (gx - 1);         // Cuda_Render.h:441-442
			const int vz = (voxel_z >> f.g.cmip) & (gz - 1);
			f.g.cidx = vx + vz * gx;
			const float corx = ray_x * ddelta, corz = ray_z * ddelta;
			f.g.pz = cos_x * vsz + sin_x * mountain;                 // Cuda_Render.h:459-464
			f.g.py = vertical ? (cos_x * mountain - sin_x * vsz) : vsx;
			f.g.py *= rx2mr;
			f.g.czz = cos_x * corz;                                  // Cuda_Render.h:483-486
			f.g.cyy = vertical ? (-sin_x * corz) : corx;
			f.g.cyy *= rx2mr;
			// The horizon only rises.  For pz > 0 a column culled now stays culled; for pz <= 0 (or NaN)
			// the test can flip, so keep those.
			f.have = !(f.g.pz * res_y2 + f.g.py <= f.g.pz * (float)Hs.ycmin) || !(f.g.pz > 0);   // Cuda_Render.h:467
			if (f.have)
			{
				const uint2 ent = __ldg(P.level[f.g.cmip].map + f.g.cidx);       // Cuda_Render.h:474-478
				f.e0 = ent.x; f.e1 = ent.y;
			}
		}
	};

	// One consume round: C1 for the next batch of live columns (take `32`, or what is left when draining), then
	// C2 + B for the batch whose run words were requested one round ago.  Returns true when the ray plane is closed.
	auto consume_round = [&](bool drain) -> bool
	{
		Stage s1;
		Geo g1;
		{
			const int n1 = qcount >= 32 ? 32 : (drain ? qcount : 0);
			s1.nvalid = n1; s1.have = gl < n1;
			s1.e0 = s1.e1 = 0;
			#pragma unroll
			for (int k = 0; k < 4; k++) s1.rw[k] = 0;
			g1.pz = g1.py = g1.czz = g1.cyy = 0; g1.cmip = 0; g1.cidx = 0;
			if (s1.have)
			{
				const uint32_t* q = queue + ((qhead + gl) & (RLERC_CQCAP - 1));
				g1.pz = __uint_as_float(q[0 * RLERC_CQCAP]); g1.py = __uint_as_float(q[1 * RLERC_CQCAP]);
				g1.czz = __uint_as_float(q[2 * RLERC_CQCAP]); g1.cyy = __uint_as_float(q[3 * RLERC_CQCAP]);
				g1.cmip = (int)q[4 * RLERC_CQCAP]; g1.cidx = (int)q[5 * RLERC_CQCAP];
				s1.e0 = q[6 * RLERC_CQCAP]; s1.e1 = q[7 * RLERC_CQCAP];
				const int sl = (int)(s1.e1 & 0xffffu);
				// element i0 of the slab stream is run 0; runs 0..7 are fetched as aligned 32-bit words
				const unsigned i0 = 2u + s1.e0;
				const uint32_t* w32 = reinterpret_cast<const uint32_t*>(P.level[g1.cmip].slabs);
				const uint32_t* p = w32 + ((i0 + (i0 & 1u)) >> 1);
				const int odd = (int)(i0 & 1u);                            // odd: words hold runs (1,2) (3,4) (5,6) (7,8)
				s1.rw[0] = (sl > 1) ? __ldg(p) : 0u;
				s1.rw[1] = (sl > 2 + odd) ? __ldg(p + 1) : 0u;
				s1.rw[2] = (sl > 4 + odd) ? __ldg(p + 2) : 0u;
				s1.rw[3] = (sl > 6 + odd) ? __ldg(p + 3) : 0u;
			}
			qhead = (qhead + n1) & (RLERC_CQCAP - 1);
			qcount -= n1;
		}
		__syncwarp();
		bool closed = false;
		if (s0.nvalid > 0)
		{
			// ---- C2. project the runs of batch s0 (their words were requested one round ago) ----
			const int ycmin = Hs.ycmin;
			int slen = 0, nr = 0;
			bool longcol = false;
			unsigned flags = 0;               // bit r: run r can be seen (z1 > 0); bit 8+r: its bottom too (z2 > 0)
			if (s0.have)
			{
				{	// run words as loaded in C1 -> runs 0..7, two per register (run 0 rides in the map entry)
					const unsigned first = s0.e1 >> 16;
					const unsigned a = s0.rw[0], b = s0.rw[1], c = s0.rw[2], d = s0.rw[3];
					if (!((2u + s0.e0) & 1u)) s0.rw[0] = first | (a & 0xffff0000u);
					else
					{
						s0.rw[0] = first | (a << 16);
						s0.rw[1] = __funnelshift_r(a, b, 16);
						s0.rw[2] = __funnelshift_r(b, c, 16);
						s0.rw[3] = __funnelshift_r(c, d, 16);
					}
				}
				slen = (int)(s0.e1 & 0xffffu);
				nr = slen < RLERC_RW ? slen : RLERC_RW;
				longcol = slen > RLERC_RW;
				int blen = 0;
				for (int r = 0; r < nr; r++)
				{
					const unsigned rw = run_word(s0.rw, r);
					const int skip = (int)(rw & 1023u), solid = (int)(rw >> 10);
					const int top = (blen + skip) << g0.cmip;                // sti_general_sti_skip
					const int bot = top + (solid << g0.cmip);                // sti_general
					blen += skip + solid;
					if (solid == 0) continue;
					const float ft = (float)top, fb2 = (float)bot;          // Cuda_Render.h:529-560
					float zz1 = g0.pz, yy1 = g0.py;
					if (mountain + ft >= 0) { zz1 += g0.czz; yy1 += g0.cyy; }
					const float z1 = zz1 + pz_add * ft;
					if (z1 <= 0) continue;
					flags |= 1u << r;
					const float y1 = yy1 + py_add * ft;
					const int sy2 = f2i(res_y2 + y1 / z1);
					int sy1 = 0;
					if (sy2 > ycmin)
					{
						float zz2 = g0.pz, yy2 = g0.py;
						if (mountain + fb2 < 0) { zz2 += g0.czz; yy2 += g0.cyy; }
						const float z2 = zz2 + pz_add * fb2;
						if (!(z2 <= 0))
						{
							flags |= 1u << (8 + r);
							const float y2 = yy2 + py_add * fb2;
							sy1 = f2i(res_y2 + y2 / z2 - 1);
						}
					}
					proj[r * 32 + gl] = make_int2(sy1, sy2);
					if (sy2 <= ycmin)
					{
						// breaks now, hence under every later (higher) horizon: later runs are dead
						nr = r + 1; longcol = false;
						break;
					}
				}
			}
			// ---- B / B0 / B1 / S. consume batch s0 (traverse_common.cuh) ----
			closed = consume_batch<IDS>(P, R, Hs, Cn, s0, g0, slen, nr, longcol, flags, proj, shade, job);
			if (Hs.ycmin >= Hs.ycmax) closed = true;                 // Cuda_Render.h:370
		}
		s0 = s1; g0 = g1;
		return closed;
	};

	__syncthreads();                                             // DDA states and flags are in place (the only block barrier)

	if (dda_warp)
	{
		// ---- the DDA warp: lane r <-> ray plane r of the block ----
		const bool mine = gl < WPB;
		DdaC Q;
		memset(&Q, 0, sizeof(Q));
		Q.done = 1;
		if (mine) Q = ddast[gl];
		bool retired = !mine || Q.done != 0;
		int produced = 0, spins = 0;
		while (true)
		{
			bool can = false;
			if (!retired)
			{
				if (ld_vol(stop + gl)) retired = true;
				else can = (produced - ld_vol(tail + gl)) < 2;
			}
			if (!__any_sync(FULL, can))
			{
				if (__all_sync(FULL, retired)) break;
				if (++spins > RLERC_C_SPIN) break;
				__nanosleep(64);
				continue;
			}
			spins = 0;
			const int slot = produced & 1;
			const int n = ddac_chunk(Q, ring + (size_t)(slot * WPB + (mine ? gl : 0)) * (RLERC_CH + 1), last_map, zfar_i, can);
			if (can)
			{
				ringn[slot * WPB + gl] = n;
				__threadfence_block();
				produced++;
				head[gl] = produced;
				if (Q.done) retired = true;      // the chunk that ended at z_far is out: nothing more to produce
			}
			__syncwarp();
		}
		return;
	}

	// ---- my ray plane: filter + consume chunk after chunk ----
	for (int c = 0; !finished; c++)
	{
		{
			int spins = 0;
			while (ld_vol(head + wid) <= c)
			{
				if (++spins > RLERC_C_SPIN) { finished = true; break; }      // protocol failure: give up, do not hang
				__nanosleep(32);
			}
			__threadfence_block();
			if (finished) break;
		}
		const int slot = c & 1;
		const float4* recs = ring + (size_t)(slot * WPB + wid) * (RLERC_CH + 1);
		const int nc = ringn[slot * WPB + wid];
		for (int b0 = 0; b0 < nc && !finished; b0 += 32)
		{
			const int nvalid = (nc - b0) < 32 ? (nc - b0) : 32;
			if (IDS && gl == 0) Cn.c_steps += nvalid;
			if (!(fsteps & 1)) { filter_test(fa); filter_issue(fa, recs + b0, nvalid); }
			else               { filter_test(fb); filter_issue(fb, recs + b0, nvalid); }
			fsteps++;
			while (qcount >= 32 && !finished) finished = consume_round(false);
		}
		__syncwarp();
		if (gl == 0) tail[wid] = c + 1;                          // the records of this chunk are in registers / queue now
		if (!finished && nc < RLERC_CH)
		{
			// z_far reached inside this chunk (Cuda_Render.h:367): test what is in flight (older set first), drain
			if (!(fsteps & 1)) { filter_test(fa); while (qcount >= 32 && !finished) finished = consume_round(false); filter_test(fb); }
			else               { filter_test(fb); while (qcount >= 32 && !finished) finished = consume_round(false); filter_test(fa); }
			while (!finished && (qcount > 0 || s0.nvalid > 0)) finished = consume_round(true);
			finished = true;
		}
	}
	if (gl == 0) stop[wid] = 1;

	if (has_row)
	{
		// sky sentinel on every pixel of the clip range that no run covered (Cuda_Render.h:255-264)
		__syncwarp();
		for (int y = ymin0 + gl; y <= ymax0; y += G)
			if (!((ymask[y >> 5] >> (y & 31)) & 1u)) row[y] = RLERC_SKY;
		if (IDS) flush_counters(P, Cn, gl, ymax0 - ymin0 + 1);
	}
}

template <bool IDS, int WPB>
static void launch_c(const TraverseParams& p, cudaStream_t st)
{
	const int rays = (p.slice_n > 1) ? owned_count(p.ray_end, p.slice_block, p.slice_n, p.slice_rank) : p.ray_end - p.ray_begin;
	if (rays <= 0) return;
	const int blocks = (rays + WPB - 1) / WPB;                        // + one DDA warp per block
	const size_t per_warp = (RLERC_C_QUEUE + 16 + RLERC_PS_WORDS + p.mask_words + 3) & ~3;
	const size_t words = (size_t)2 * WPB * (RLERC_CH + 1) * 4 + (size_t)WPB * RLERC_C_STATE + 8 * WPB + (size_t)WPB * per_warp;
	const size_t smem = words * sizeof(uint32_t);
	// dynamic shared memory above 48 KB is an opt-in per kernel AND per device
	static size_t configured_on[64] = { 0 };
	int dev = 0;
	cudaGetDevice(&dev);
	size_t& configured = configured_on[dev & 63];
	if (smem > configured)
	{
		cudaFuncSetAttribute(k_traverse_c<IDS, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = smem;
	}
	k_traverse_c<IDS, WPB><<<blocks, (WPB + 1) * 32, smem, st>>>(p, rays);
}

void launch_traverse_chunk(const TraverseParams& p, bool ids, int wpb, cudaStream_t st)
{
	if (wpb == 7) { if (ids) launch_c<true, 7>(p, st); else launch_c<false, 7>(p, st); }
	else          { if (ids) launch_c<true, 3>(p, st); else launch_c<false, 3>(p, st); }
}

} // namespace rlerc
