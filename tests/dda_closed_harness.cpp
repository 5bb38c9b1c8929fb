/* TEST INFRASTRUCTURE.  Pins rle-based-voxel-raycasting_b200/csrc/dda_closed.cuh (the closed form of the DDA's
 * float recurrences that k_traverse_w evaluates lane-parallel) against the serial recurrence of the reference
 * (R/src/Cuda_Render.h:286-300,343-367,398-414) on the CPU.  Built and run by tests/test_dda_closed.py.
 *
 *   var  N seed   : single-variable recurrences v <- fl(v + g), random magnitudes, LOD doublings, random advances
 *   fast N seed   : the single-regime fast path of k_traverse_f (rank of every firing by count estimate + exact fix-up,
 *                   validity check after the fact, serial fallback), compared crossing by crossing with the serial loop.
 *   dda  N seed   : whole two-track DDA, batch by batch, exactly as the kernel organises it (6 owner "lanes", plan
 *                   table, merge-path search per lane, validity prefix, serial fallback), compared crossing by
 *                   crossing with the serial loop.
 * Prints "ok <checked> <closed-form batches> <fallback batches>" or the first mismatch; exit code 0 / 1.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dda_closed.cuh"

using namespace rlerc;

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double urand() { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
static float fadd(float a, float b) { volatile float r = a + b; return r; }

static int test_var(long N)
{
	long checked = 0;
	for (long t = 0; t < N; t++)
	{
		const int kind = (int)(rnd() % 8);
		float g = (float)ldexp(1.0 + urand(), (int)(rnd() % 60) - 30);
		if (kind == 6) g = (float)ldexp(1.0, (int)(rnd() % 40) - 20);                 /* power of two: ties everywhere */
		if (kind == 7) { int b = dda_f2b(g); b &= ~((1 << (rnd() % 23)) - 1); g = dda_b2f(b); }   /* few mantissa bits */
		if (rnd() & 1) g = -g;
		float v;
		if (kind == 0) v = g * (float)urand();                                          /* ray start */
		else if (kind == 1) v = -g * (float)urand();                                    /* opposite sign start */
		else if (kind == 2) v = 0.0f;
		else if (kind == 3) v = g * (float)ldexp(1.0 + urand(), (int)(rnd() % 40) - 8); /* anywhere */
		else v = g * (float)((double)(rnd() % 100000) + urand());                       /* k-th crossing */
		DdaVar S;
		dda_var_init(S, v, g);
		float sv = v, sg = g;
		for (int round = 0; round < 200; round++)
		{
			if (rnd() % 37 == 0) { sg = sg * 2.0f; dda_var_double(S); }
			DdaPlan p; int F2, L2;
			dda_var_plan(S, 33, p, F2, L2);
			if (p.V < 1) { printf("V < 1: v %a g %a\n", sv, sg); return 1; }
			const int reach = p.V < 40 ? p.V : 40;
			float w = sv;
			for (int i = 0; i <= reach; i++)
			{
				const int got = dda_plan_eval(p, i);
				if (got != dda_f2b(w))
				{
					printf("mismatch: t %ld round %d i %d  v0 %a g %a  want %a (%08x) got %a (%08x)  plan L1 %d D1 %d V %d\n",
					       t, round, i, sv, sg, w, dda_f2b(w), dda_b2f(got), got, p.L1, p.D1, p.V);
					return 1;
				}
				w = fadd(w, sg);
				checked++;
			}
			const int maxn = p.V < 32 ? p.V : 32;
			const int n = (int)(rnd() % (maxn + 1));
			dda_var_advance(S, p, F2, L2, n);
			for (int i = 0; i < n; i++) sv = fadd(sv, sg);
			if (S.b != dda_f2b(sv)) { printf("advance mismatch t %ld round %d n %d\n", t, round, n); return 1; }
		}
	}
	printf("ok %ld 0 0\n", checked);
	return 0;
}

/* ---- the whole DDA ------------------------------------------------------------------------------------------ */
struct Serial {
	float g0x, g0y, g1x, g1y, i0x, i0y, i1x, i1y, gd0, gd1, d0, d1;
	float posx, posy, dist_now; int index, mip, zi, dzi, mapswitch;
};
struct Rec { float sd, px, py; int mip; };   /* signed distance (negative: z-track fired), position, mip level */

static void lod_switch(Serial& S, int last_map)
{
	if (S.mip < last_map) S.mip++;
	S.g0x *= 2; S.g0y *= 2; S.g1x *= 2; S.g1y *= 2; S.gd0 *= 2; S.gd1 *= 2;
	S.mapswitch *= 2; S.dzi *= 2;
}

/* one crossing of the serial loop (Cuda_Render.h:343-367,398-414); false at z_far */
static bool serial_step(Serial& S, int last_map, int zfar, Rec& r)
{
	while (S.zi > S.mapswitch) lod_switch(S, last_map);
	if (S.zi + S.dzi > zfar) return false;
	S.zi += S.dzi;
	const bool t1 = S.d1 < S.d0;
	S.dist_now = t1 ? S.d1 : S.d0;
	S.posx = t1 ? S.i1x : S.i0x;
	S.posy = t1 ? S.i1y : S.i0y;
	S.index = t1 ? 1 : 0;
	r.sd = t1 ? -S.d1 : S.d0; r.px = S.posx; r.py = S.posy; r.mip = S.mip;
	if (t1) { S.d1 = fadd(S.d1, S.gd1); S.i1x = fadd(S.i1x, S.g1x); S.i1y = fadd(S.i1y, S.g1y); }
	else    { S.d0 = fadd(S.d0, S.gd0); S.i0x = fadd(S.i0x, S.g0x); S.i0y = fadd(S.i0y, S.g0y); }
	return true;
}

static bool same(float a, float b) { return dda_f2b(a) == dda_f2b(b) || (a != a && b != b); }

static int test_dda(long N)
{
	long checked = 0, closed = 0, fallback = 0;
	for (long t = 0; t < N; t++)
	{
		/* a ray plane as dda_init makes it (Cuda_Render.h:270-305) */
		const double ang = urand() * 6.283185307179586;
		float drx = (float)cos(ang), dry = (float)sin(ang);
		if (rnd() % 16 == 0) dry = (float)ldexp(urand(), -(int)(rnd() % 30));           /* nearly axis aligned */
		if (rnd() % 16 == 0) drx = (float)ldexp(urand(), -(int)(rnd() % 30));
		if (rnd() % 64 == 0) dry = 0.0f;                                                /* exactly axis aligned */
		float vpx = (float)(urand() * 20000.0 - (rnd() % 4 == 0 ? 10000.0 : 0.0)), vpz = (float)(urand() * 20000.0);
		if (rnd() % 32 == 0) vpx = floorf(vpx);
		float fx = vpx - (float)(int)vpx, fy = vpz - (float)(int)vpz;
		float sgx = -1, sgy = -1;
		if (drx >= 0) { sgx = 1; fx = 1 - fx; }
		if (dry >= 0) { sgy = 1; fy = 1 - fy; }
		Serial S;
		S.g0y = dry / fabsf(drx); S.g0x = sgx;
		S.g1x = drx / fabsf(dry); S.g1y = sgy;
		S.i0x = S.g0x * fx; S.i0y = S.g0y * fx;
		S.i1x = S.g1x * fy; S.i1y = S.g1y * fy;
		S.gd0 = sqrtf(S.g0x * S.g0x + S.g0y * S.g0y);
		S.gd1 = sqrtf(S.g1x * S.g1x + S.g1y * S.g1y);
		S.d0 = sqrtf(S.i0x * S.i0x + S.i0y * S.i0y);
		S.d1 = sqrtf(S.i1x * S.i1x + S.i1y * S.i1y);
		S.posx = S.posy = S.dist_now = 0; S.index = 0; S.mip = 0; S.zi = 0; S.dzi = 1;
		S.mapswitch = 100 + (int)(rnd() % 2000);
		const int last_map = 9, zfar = 80000;
		for (int k = (int)(rnd() % 3); k > 0; k--) lod_switch(S, last_map);            /* y_map_switch > 512 at the start */
		Serial Q = S;                                                                   /* the serial truth */

		/* kernel-side state: six owner lanes + uniform integers */
		const bool merge_ok = !(S.d0 != S.d0) && !(S.d1 != S.d1) && !(S.gd0 != S.gd0) && !(S.gd1 != S.gd1);
		DdaVar var[6];
		float carry_sd = 0, carry_px = 0, carry_py = 0;       /* state before the batch (record of the last crossing) */
		int mip = S.mip, zi = S.zi, dzi = S.dzi, mapswitch = S.mapswitch;
		float* const vals[6] = { &S.d0, &S.i0x, &S.i0y, &S.d1, &S.i1x, &S.i1y };
		float* const grads[6] = { &S.gd0, &S.g0x, &S.g0y, &S.gd1, &S.g1x, &S.g1y };
		for (int k = 0; k < 6; k++) dda_var_init(var[k], *vals[k], *grads[k]);
		bool done = false;
		int guard = 0;
		while (!done && guard++ < 4000)
		{
			Rec out[32]; int n = 0;
			bool use_serial = !merge_ok;
			/* LOD switches due before this batch */
			while (zi > mapswitch)
			{
				if (mip < last_map) mip++;
				for (int k = 0; k < 6; k++) dda_var_double(var[k]);
				mapswitch *= 2; dzi *= 2;
			}
			const int lod_free = (mapswitch - zi) / dzi + 1;
			const int far_free = (zfar - zi) / dzi;
			if (far_free <= 0) break;
			int want = 32; if (lod_free < want) want = lod_free; if (far_free < want) want = far_free;
			DdaPlan plan[6]; int F2[6], L2[6];
			int ia[32], fired[32];
			if (!use_serial)
			{
				for (int k = 0; k < 6; k++) dda_var_plan(var[k], 33, plan[k], F2[k], L2[k]);
				int V0 = plan[0].V, V1 = plan[3].V;
				for (int k = 1; k < 3; k++) { if (plan[k].V < V0) V0 = plan[k].V; if (plan[3 + k].V < V1) V1 = plan[3 + k].V; }
				int T0[32], T1[32];
				for (int l = 0; l < 32; l++)
				{
					T0[l] = l <= V0 ? dda_plan_eval(plan[0], l) : RLERC_DDA_INF_BITS;
					T1[l] = l <= V1 ? dda_plan_eval(plan[3], l) : RLERC_DDA_INF_BITS;
				}
				int first_bad = 32;
				for (int s = 31; s >= 0; s--)
				{
					const int i = dda_merge_search(T0, T1, s), j = s - i;
					ia[s] = i;
					if (i > V0 || j > V1) { first_bad = s; continue; }
					fired[s] = T1[j] < T0[i] ? 1 : 0;
				}
				if (first_bad < want) use_serial = true;
				else
				{
					n = want;
					for (int s = 0; s < n; s++)
					{
						const int i = ia[s], j = s - i, tr = fired[s];
						const int idx = tr ? j : i;
						const float d = dda_b2f(dda_plan_eval(plan[tr * 3], idx));
						out[s].sd = tr ? -d : d;
						out[s].px = dda_b2f(dda_plan_eval(plan[tr * 3 + 1], idx));
						out[s].py = dda_b2f(dda_plan_eval(plan[tr * 3 + 2], idx));
						out[s].mip = mip;
					}
					const int n0 = ia[n - 1] + (fired[n - 1] ? 0 : 1), n1 = n - n0;
					for (int k = 0; k < 3; k++) { dda_var_advance(var[k], plan[k], F2[k], L2[k], n0); dda_var_advance(var[3 + k], plan[3 + k], F2[3 + k], L2[3 + k], n1); }
					zi += n * dzi;
					carry_sd = out[n - 1].sd; carry_px = out[n - 1].px; carry_py = out[n - 1].py;
					closed++;
				}
			}
			if (use_serial)
			{
				/* rebuild the uniform float state from the owner lanes, run the serial batch, hand it back */
				Serial T;
				memset(&T, 0, sizeof(T));
				T.d0 = dda_b2f(var[0].b); T.i0x = dda_b2f(var[1].b); T.i0y = dda_b2f(var[2].b);
				T.d1 = dda_b2f(var[3].b); T.i1x = dda_b2f(var[4].b); T.i1y = dda_b2f(var[5].b);
				T.gd0 = dda_b2f(var[0].gb); T.g0x = dda_b2f(var[1].gb); T.g0y = dda_b2f(var[2].gb);
				T.gd1 = dda_b2f(var[3].gb); T.g1x = dda_b2f(var[4].gb); T.g1y = dda_b2f(var[5].gb);
				T.mip = mip; T.zi = zi; T.dzi = dzi; T.mapswitch = mapswitch;
				n = 0;
				for (; n < want; n++) if (!serial_step(T, last_map, zfar, out[n])) break;
				float* const tv[6] = { &T.d0, &T.i0x, &T.i0y, &T.d1, &T.i1x, &T.i1y };
				float* const tg[6] = { &T.gd0, &T.g0x, &T.g0y, &T.gd1, &T.g1x, &T.g1y };
				for (int k = 0; k < 6; k++) dda_var_init(var[k], *tv[k], *tg[k]);
				mip = T.mip; zi = T.zi; dzi = T.dzi; mapswitch = T.mapswitch;
				if (n) { carry_sd = out[n - 1].sd; carry_px = out[n - 1].px; carry_py = out[n - 1].py; }
				fallback++;
			}
			/* compare with the serial truth */
			for (int s = 0; s < n; s++)
			{
				Rec r;
				if (!serial_step(Q, last_map, zfar, r)) { printf("serial ended early t %ld\n", t); return 1; }
				if (!same(r.sd, out[s].sd) || !same(r.px, out[s].px) || !same(r.py, out[s].py) || r.mip != out[s].mip)
				{
					printf("dda mismatch t %ld batch %d s %d: want (%a %a %a %d) got (%a %a %a %d) drx %a dry %a serial %d\n",
					       t, guard, s, r.sd, r.px, r.py, r.mip, out[s].sd, out[s].px, out[s].py, out[s].mip, drx, dry, (int)use_serial);
					return 1;
				}
				checked++;
			}
			if (zi != Q.zi) { printf("zi mismatch t %ld\n", t); return 1; }
			(void)carry_sd; (void)carry_px; (void)carry_py;
		}
		/* both must be at z_far now */
		Rec r;
		if (serial_step(Q, last_map, zfar, r)) { printf("closed form ended early t %ld (zi %d)\n", t, zi); return 1; }
	}
	printf("ok %ld %ld %ld\n", checked, closed, fallback);
	return 0;
}


/* ---- the single-regime fast path of k_traverse_f, batch by batch -------------------------------------------- */
static int test_fast(long N)
{
	long checked = 0, fastb = 0, fallback = 0;
	for (long t = 0; t < N; t++)
	{
		const double ang = urand() * 6.283185307179586;
		float drx = (float)cos(ang), dry = (float)sin(ang);
		if (rnd() % 16 == 0) dry = (float)ldexp(urand(), -(int)(rnd() % 30));
		if (rnd() % 16 == 0) drx = (float)ldexp(urand(), -(int)(rnd() % 30));
		if (rnd() % 64 == 0) dry = 0.0f;
		float vpx = (float)(urand() * 20000.0 - (rnd() % 4 == 0 ? 10000.0 : 0.0)), vpz = (float)(urand() * 20000.0);
		if (rnd() % 32 == 0) vpx = floorf(vpx);
		float fx = vpx - (float)(int)vpx, fy = vpz - (float)(int)vpz;
		float sgx = -1, sgy = -1;
		if (drx >= 0) { sgx = 1; fx = 1 - fx; }
		if (dry >= 0) { sgy = 1; fy = 1 - fy; }
		Serial S;
		S.g0y = dry / fabsf(drx); S.g0x = sgx;
		S.g1x = drx / fabsf(dry); S.g1y = sgy;
		S.i0x = S.g0x * fx; S.i0y = S.g0y * fx;
		S.i1x = S.g1x * fy; S.i1y = S.g1y * fy;
		S.gd0 = sqrtf(S.g0x * S.g0x + S.g0y * S.g0y);
		S.gd1 = sqrtf(S.g1x * S.g1x + S.g1y * S.g1y);
		S.d0 = sqrtf(S.i0x * S.i0x + S.i0y * S.i0y);
		S.d1 = sqrtf(S.i1x * S.i1x + S.i1y * S.i1y);
		S.posx = S.posy = S.dist_now = 0; S.index = 0; S.mip = 0; S.zi = 0; S.dzi = 1;
		S.mapswitch = 100 + (int)(rnd() % 2000);
		const int last_map = 9, zfar = 80000;
		for (int k = (int)(rnd() % 3); k > 0; k--) lod_switch(S, last_map);
		Serial Q = S;                                                                   /* the serial truth */
		const bool sorted = !(S.d0 != S.d0) && !(S.d1 != S.d1) && !(S.gd0 != S.gd0) && !(S.gd1 != S.gd1);
		/* kernel-side state: float heads + gradients (uniform), D / L per variable (owner lanes) */
		float* const hv[6] = { &S.d0, &S.i0x, &S.i0y, &S.d1, &S.i1x, &S.i1y };
		float* const gv[6] = { &S.gd0, &S.g0x, &S.g0y, &S.gd1, &S.g1x, &S.g1y };
		int D[6], L[6];
		for (int k = 0; k < 6; k++) dda_fast_regime(dda_f2b(*hv[k]), dda_f2b(*gv[k]), D[k], L[k]);
		int guard = 0;
		while (guard++ < 4000)
		{
			Rec out[32]; int n = 0;
			bool lod = false;
			while (S.zi > S.mapswitch) { lod_switch(S, last_map); lod = true; }
			if (lod) for (int k = 0; k < 6; k++) dda_fast_regime(dda_f2b(*hv[k]), dda_f2b(*gv[k]), D[k], L[k]);
			const int lod_free = (S.mapswitch - S.zi) / S.dzi + 1;
			const int far_free = (zfar - S.zi) / S.dzi;
			if (far_free <= 0) break;
			int want = 32; if (lod_free < want) want = lod_free; if (far_free < want) want = far_free;
			int L0 = L[0], L1 = L[3];
			for (int k = 1; k < 3; k++) { if (L[k] < L0) L0 = L[k]; if (L[3 + k] < L1) L1 = L[3 + k]; }
			bool done_fast = false;
			if (sorted && L0 >= 1 && L1 >= 1)
			{
				bool ok = true;
				int slot_tr[33], slot_ix[33], filled[33];
				memset(filled, 0, sizeof(filled));
				const int b0 = dda_f2b(S.d0), b1 = dda_f2b(S.d1);
				const float inv0 = 1.0f / S.gd0, inv1 = 1.0f / S.gd1;
				int n0 = 0;
				for (int j = 0; j < 32; j++)
				{
					const int A0 = b0 + j * D[0], A1 = b1 + j * D[3];
					const int r0 = j + dda_count_below(b1, D[3], S.d1, inv1, A0, false, ok);
					const int r1 = j + dda_count_below(b0, D[0], S.d0, inv0, A1, true, ok);
					if (r0 < want) { n0++; filled[r0]++; slot_tr[r0] = 0; slot_ix[r0] = j; }
					if (r1 < want) { filled[r1]++; slot_tr[r1] = 1; slot_ix[r1] = j; }
				}
				const int n1 = want - n0;
				if (ok && n0 <= L0 && n1 <= L1)
				{
					for (int s = 0; s < want; s++) if (filled[s] != 1) { printf("slot %d filled %d times (t %ld batch %d)\n", s, filled[s], t, guard); return 1; }
					for (int s = 0; s < want; s++)
					{
						const int tr = slot_tr[s], ix = slot_ix[s];
						const float d = dda_b2f(dda_f2b(*hv[tr * 3]) + ix * D[tr * 3]);
						out[s].sd = tr ? -d : d;
						out[s].px = dda_b2f(dda_f2b(*hv[tr * 3 + 1]) + ix * D[tr * 3 + 1]);
						out[s].py = dda_b2f(dda_f2b(*hv[tr * 3 + 2]) + ix * D[tr * 3 + 2]);
						out[s].mip = S.mip;
					}
					for (int k = 0; k < 3; k++)
					{
						*hv[k] = dda_b2f(dda_f2b(*hv[k]) + n0 * D[k]); if (L[k] < RLERC_DDA_LBIG) L[k] -= n0;
						*hv[3 + k] = dda_b2f(dda_f2b(*hv[3 + k]) + n1 * D[3 + k]); if (L[3 + k] < RLERC_DDA_LBIG) L[3 + k] -= n1;
					}
					S.zi += want * S.dzi;
					n = want;
					done_fast = true;
					fastb++;
				}
			}
			if (!done_fast)
			{
				for (; n < want; n++) if (!serial_step(S, last_map, zfar, out[n])) break;
				for (int k = 0; k < 6; k++) dda_fast_regime(dda_f2b(*hv[k]), dda_f2b(*gv[k]), D[k], L[k]);
				fallback++;
			}
			for (int s = 0; s < n; s++)
			{
				Rec r;
				if (!serial_step(Q, last_map, zfar, r)) { printf("serial ended early t %ld\n", t); return 1; }
				if (!same(r.sd, out[s].sd) || !same(r.px, out[s].px) || !same(r.py, out[s].py) || r.mip != out[s].mip)
				{
					printf("fast mismatch t %ld batch %d s %d: want (%a %a %a %d) got (%a %a %a %d) fast %d\n",
					       t, guard, s, r.sd, r.px, r.py, r.mip, out[s].sd, out[s].px, out[s].py, out[s].mip, (int)done_fast);
					return 1;
				}
				checked++;
			}
			if (S.zi != Q.zi) { printf("zi mismatch t %ld\n", t); return 1; }
			for (int k = 0; k < 6; k++)
			{
				float* const qv[6] = { &Q.d0, &Q.i0x, &Q.i0y, &Q.d1, &Q.i1x, &Q.i1y };
				if (!same(*hv[k], *qv[k])) { printf("head %d mismatch t %ld batch %d fast %d\n", k, t, guard, (int)done_fast); return 1; }
			}
		}
		Rec r;
		if (serial_step(Q, last_map, zfar, r)) { printf("fast path ended early t %ld (zi %d)\n", t, S.zi); return 1; }
	}
	printf("ok %ld %ld %ld\n", checked, fastb, fallback);
	return 0;
}

int main(int argc, char** argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s var|dda N [seed]\n", argv[0]); return 2; }
	const long N = atol(argv[2]);
	if (argc > 3) rng_state ^= (uint64_t)atoll(argv[3]) * 0x9E3779B97F4A7C15ull;
	for (int i = 0; i < 8; i++) rnd();
	if (!strcmp(argv[1], "var")) return test_var(N);
	if (!strcmp(argv[1], "dda")) return test_dda(N);
	if (!strcmp(argv[1], "fast")) return test_fast(N);
	return 2;
}
