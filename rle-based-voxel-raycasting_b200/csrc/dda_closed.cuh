// Closed form of the DDA's float recurrences (Cuda_Render.h:286-300,398-414), exact to the bit.
//
// Every one of the six DDA variables (dds_dist0/1, isect0/1 .x/.y) is a recurrence v <- fl(v + g) with a
// constant increment g of the same sign as v (g doubles at a LOD switch).  While v stays inside one binade
// [2^e, 2^(e+1)) its spacing u = ulp(v) is constant, so the rounded sum moves the BIT PATTERN of v by a
// constant integer:
//     G = floor(|g| / u),  rem = |g| - G*u
//     rem <  u/2 : +G          rem > u/2 : +G+1
//     rem == u/2 : round-half-even: the result is even, so after the first step the increment is G
//                  rounded up to even; the first step adds G|1 from an odd mantissa, that same even
//                  increment from an even one.
// (consecutive floats of one sign have consecutive bit patterns, also across the binade boundary, so the
// integer formula may END exactly on 2^(e+1).)  A step is covered by the formula iff its exact sum is
// below 2^(e+1); the step that leaves the binade is taken with one real float add, after which a second
// regime of the same kind starts.  Two regimes cover >= 33 steps of a variable unless it doubles twice
// within 33 steps, which only happens in the first crossings of a ray plane; the kernel falls back to the
// serial recurrence for those batches (and for NaN / opposite-sign states, where nothing here is assumed).
//
// Host + device: tests/dda_closed_harness.cpp pins this file against the serial recurrence on the CPU.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RLERC_HD __host__ __device__ __forceinline__
#else
#define RLERC_HD inline
#endif

namespace rlerc {

#define RLERC_DDA_INF_BITS 0x7f800000
#define RLERC_DDA_LBIG (1 << 28)

RLERC_HD int dda_f2b(float f)
{
#if defined(__CUDA_ARCH__)
	return __float_as_int(f);
#else
	int b; memcpy(&b, &f, 4); return b;
#endif
}
RLERC_HD float dda_b2f(int b)
{
#if defined(__CUDA_ARCH__)
	return __int_as_float(b);
#else
	float f; memcpy(&f, &b, 4); return f;
#endif
}
// v + g in round-to-nearest-even single precision, never contracted or reassociated
RLERC_HD float dda_fadd(float a, float b)
{
#if defined(__CUDA_ARCH__)
	return __fadd_rn(a, b);
#else
	volatile float r = a + b; return r;
#endif
}

// One variable of the DDA as its owner lane keeps it: bit pattern of the current head value, of the
// increment, and the linear regime that starts at the head: value after j >= 1 steps = b + F + (j-1)*D for
// j <= L.
struct DdaVar { int b, gb, F, D, L; };

// (F, D, L) of the regime starting at bit pattern b with increment bits gb
RLERC_HD void dda_regime(int b, int gb, int& F, int& D, int& L)
{
	const unsigned bm = (unsigned)b & 0x7fffffffu, gm = (unsigned)gb & 0x7fffffffu;
	const int eb = (int)(bm >> 23), eg = (int)(gm >> 23);
	F = 0; D = 0; L = 0;
	if (eb == 255 || eg == 255) return;                              // inf / NaN: exact adds only
	if (bm != 0 && ((b ^ gb) < 0)) return;                           // opposite signs: |v| shrinks, no regime
	if (gm == 0) { L = RLERC_DDA_LBIG; return; }                     // v + 0 = v (v = -0 + 0 is excluded: bm != 0 or same sign)
	const int a = (int)(eb ? ((bm & 0x7fffffu) | 0x800000u) : bm);   // mantissa in units of u
	const int mg = (int)(eg ? ((gm & 0x7fffffu) | 0x800000u) : gm);
	const int s = (eb ? eb : 1) - (eg ? eg : 1);                     // u = 2^s units of g's last place
	if (s < 0) return;                                               // |g| has the larger exponent: the sum leaves the binade
	int G;
	if (s >= 26) { G = 0; F = D = 0; }                               // |g| < u/4
	else if (s == 0) { G = mg; F = D = G; }
	else
	{
		G = mg >> s;
		const int rem = mg & ((1 << s) - 1), half = 1 << (s - 1);
		if (rem == half) { D = (G + 1) & ~1; F = (a & 1) ? (G | 1) : D; }
		else { D = G + (rem > half ? 1 : 0); F = D; }
	}
	// step 1 is covered iff a + G < 2^24; step j >= 2 iff a + F + (j-2)*D + G < 2^24
	const int lim = 1 << 24;
	if (a + G >= lim) return;
	const int N = lim - 1 - G - a - F;
	if (N < 0) { L = 1; return; }
	if (D == 0) { L = RLERC_DDA_LBIG; return; }
	L = 2 + N / D;
	if (L > RLERC_DDA_LBIG) L = RLERC_DDA_LBIG;
}

RLERC_HD void dda_var_init(DdaVar& v, float value, float g)
{
	v.b = dda_f2b(value); v.gb = dda_f2b(g);
	dda_regime(v.b, v.gb, v.F, v.D, v.L);
}

// LOD switch: the increment doubles (Cuda_Render.h:357-364)
RLERC_HD void dda_var_double(DdaVar& v)
{
	v.gb = dda_f2b(dda_b2f(v.gb) * 2.0f);
	dda_regime(v.b, v.gb, v.F, v.D, v.L);
}

// The next steps of one variable, as 8 words the whole warp can evaluate:
//   value(i) = i <= L1 ? (i == 0 ? b0 : c1 + i*D1) : (i == L1+1 ? bx : c2 + i*D2)     for 0 <= i <= V
struct DdaPlan { int b0, c1, D1, L1, bx, c2, D2, V; };

RLERC_HD int dda_plan_eval(const DdaPlan& p, int i)
{
	if (i <= p.L1) return i == 0 ? p.b0 : p.c1 + i * p.D1;
	return i == p.L1 + 1 ? p.bx : p.c2 + i * p.D2;
}

// Plan for at least `want` steps if two regimes reach that far.  F2/L2 of the second regime are returned for
// dda_var_advance.
RLERC_HD void dda_var_plan(const DdaVar& v, int want, DdaPlan& p, int& F2, int& L2)
{
	p.b0 = v.b; p.c1 = v.b + v.F - v.D; p.D1 = v.D;
	F2 = 0; L2 = 0;
	if (v.L >= want) { p.L1 = v.L; p.bx = 0; p.c2 = 0; p.D2 = 0; p.V = v.L; return; }
	p.L1 = v.L;
	const int x = v.L + 1;
	const int last = v.L == 0 ? v.b : p.c1 + v.L * v.D;
	p.bx = dda_f2b(dda_fadd(dda_b2f(last), dda_b2f(v.gb)));
	int D2;
	dda_regime(p.bx, v.gb, F2, D2, L2);
	p.D2 = D2;
	p.c2 = p.bx + F2 - (x + 1) * D2;
	const long long V = (long long)x + L2;
	p.V = V > RLERC_DDA_LBIG ? RLERC_DDA_LBIG : (int)V;
}

// The variable made n <= p.V steps of its plan.
RLERC_HD void dda_var_advance(DdaVar& v, const DdaPlan& p, int F2, int L2, int n)
{
	if (n <= 0) return;
	v.b = dda_plan_eval(p, n);
	if (n <= p.L1) { v.F = v.D; v.L = (v.L >= RLERC_DDA_LBIG) ? v.L : v.L - n; return; }
	const int k = n - (p.L1 + 1);            // steps made inside the second regime
	v.D = p.D2;
	if (k == 0) { v.F = F2; v.L = L2; }
	else { v.F = p.D2; v.L = (L2 >= RLERC_DDA_LBIG) ? L2 : L2 - k; }
}

// Merge-path search: T0 / T1 = the next 32 firing distances of the two tracks (bit patterns of non-negative
// floats, +inf beyond what is known).  Returns how many of the first s crossings the x-track (track 0) makes;
// the serial loop fires the z-track iff d1 < d0 (Cuda_Render.h:398), i.e. ties go to track 0.
RLERC_HD int dda_merge_search(const int* T0, const int* T1, int s)
{
	int lo = 0, hi = s;
	for (int it = 0; it < 5; it++)
	{
		if (lo < hi)
		{
			const int mid = (lo + hi) >> 1;
			// non-negative floats order like their bit patterns
			if (T0[mid] <= T1[s - mid - 1]) lo = mid + 1; else hi = mid;
		}
	}
	return lo;
}

// ---- single-regime fast path (k_traverse_f) ------------------------------------------------------------------
// While none of a track's three variables leaves its binade, firing i of the track has the bit patterns
// b + i*D.  Lane j takes firing j of BOTH tracks and finds its position in the merged order directly:
//     rank(x-track firing i) = i + #{k : d1[k] <  d0[i]}        (the serial loop fires the z-track iff d1 < d0,
//     rank(z-track firing j) = j + #{k : d0[k] <= d1[j]}         Cuda_Render.h:398: ties go to the x-track)
// The count is estimated from the real-valued crossing (A - head) / gradient and fixed up with exact integer
// compares of the bit patterns (non-negative floats order like their bits).  Values past the regime's exact
// range are linear extrapolations; they stay monotone and above every exact value, so a comparison against them
// still answers correctly for every firing that belongs to the batch as long as the batch consumes no more than
// L exact firings of either track (checked after the fact: n0 <= L0, n1 <= L1).

// #{k in [0, 34] : (b + k*D) < A}, or <= A with le.  ok = false when the fix-up did not converge.
RLERC_HD int dda_count_below(int b, int D, float head, float inv_g, int A, bool le, bool& ok)
{
	const float t = (dda_b2f(A) - head) * inv_g;
	int c = !(t > -1.0f) ? 0 : (t >= 33.0f ? 34 : (int)t + 1);
	const int adj = le ? 1 : 0;                                      // (x <= A) == (x < A + 1) on integers
	#pragma unroll
	for (int it = 0; it < 3; it++) if (c < 34 && (b + c * D) < A + adj) c++;
	#pragma unroll
	for (int it = 0; it < 3; it++) if (c > 0 && !((b + (c - 1) * D) < A + adj)) c--;
	if ((c < 34 && (b + c * D) < A + adj) || (c > 0 && !((b + (c - 1) * D) < A + adj))) ok = false;
	return c;
}

// regime of one variable for the fast path: D, and L = exact firings available (0: take the serial recurrence)
RLERC_HD void dda_fast_regime(int b, int gb, int& D, int& L)
{
	int F;
	dda_regime(b, gb, F, D, L);
	if (F != D || D > (1 << 22) || D < 0) L = 0;                     // odd start of a tie regime / huge increment
	if (((unsigned)b & 0x7fffffffu) > 0x70000000u) L = 0;            // b + 34*D must stay a positive integer
}

} // namespace rlerc
