// Per-frame camera / ray-map setup on the host: frustum -> vanishing point -> the four
// screen quadrants and their ray counts.  Replaces RayMap::get_ray_map
// (R/src/RayMap.h:98-402) and must reproduce its single-precision arithmetic to the
// bit, because int() snapping of the quadrant borders amplifies ulp differences into
// different ray counts.  This translation unit is therefore compiled with
// -ffp-contract=off, and every expression keeps the reference's evaluation order and
// float/double typing (cites inline).  Uses libm sinf/cosf/acosf/sqrtf exactly where the
// reference's C++ overloads resolve to them.
#include <cmath>
#include <cstring>
#include "rlerc_internal.h"

namespace rlerc {
namespace {

struct V3 { float x, y, z; };

inline V3 v3(float x, float y, float z) { V3 r = { x, y, z }; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
// vec3f::dot(a) = a.x*x + a.y*y + a.z*z (VecMath.h:60); commutative per term, same order
inline float dot(V3 self, V3 a) { return a.x * self.x + a.y * self.y + a.z * self.z; }
inline float length(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct M44 { float m[4][4]; };

inline void ident(M44& a)
{
	for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) a.m[i][j] = (i == j) ? 1.0f : 0.0f;
}
// _matrix44::rotate_x / rotate_y (R/inc/mathlib/_matrix44.h:529-560)
inline void rotate_x(M44& a, float ang)
{
	const float c = std::cos(ang), s = std::sin(ang);
	for (int i = 0; i < 4; i++)
	{
		const float m1 = a.m[i][1], m2 = a.m[i][2];
		a.m[i][1] = m1 * c + m2 * -s;
		a.m[i][2] = m1 * s + m2 * c;
	}
}
inline void rotate_y(M44& a, float ang)
{
	const float c = std::cos(ang), s = std::sin(ang);
	for (int i = 0; i < 4; i++)
	{
		const float m0 = a.m[i][0], m2 = a.m[i][2];
		a.m[i][0] = m0 * c + m2 * s;
		a.m[i][2] = m0 * -s + m2 * c;
	}
}
// _matrix44 * _vector3 (R/inc/mathlib/_matrix44.h:863-869): row-vector convention
inline V3 xform(const M44& a, V3 v)
{
	return v3(a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z + a.m[3][0],
	          a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z + a.m[3][1],
	          a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z + a.m[3][2]);
}
// _matrix44::invert_simpler (R/inc/mathlib/_matrix44.h:431-441)
inline void invert_simpler(M44& a)
{
	std::swap(a.m[0][1], a.m[1][0]);
	std::swap(a.m[0][2], a.m[2][0]);
	std::swap(a.m[1][2], a.m[2][1]);
	const float m30 = -(a.m[0][0] * a.m[3][0] + a.m[1][0] * a.m[3][1] + a.m[2][0] * a.m[3][2]);
	const float m31 = -(a.m[0][1] * a.m[3][0] + a.m[1][1] * a.m[3][1] + a.m[2][1] * a.m[3][2]);
	a.m[3][2] = -(a.m[0][2] * a.m[3][0] + a.m[1][2] * a.m[3][1] + a.m[2][2] * a.m[3][2]);
	a.m[3][1] = m31;
	a.m[3][0] = m30;
}

inline rlerc_vec3f out3(V3 a) { rlerc_vec3f r = { a.x, a.y, a.z }; return r; }
inline V3 in3(rlerc_vec3f a) { return v3(a.x, a.y, a.z); }

// One side of the screen as seen from the vanishing point: the two border points of the
// quadrant's fan, snapped outward to the ray grid, and the ray count between them.
// kind 0 = upper (RayMap.h:209-240), 1 = lower (:243-277), 2 = left (:280-313), 3 = right (:316-352)
struct Quad { V3 a, b; int res; float ofs_min, ofs_max; bool active; };

} // namespace

void get_ray_map(const float pos[3], const float rot[3], float border, int rays_casted_res, rlerc_raymap* rm)
{
	rm->rotation = out3(v3(rot[0], rot[1], rot[2]));
	rm->position = out3(v3(pos[0], pos[1], pos[2]));
	rm->border = border;
	rm->clip_min = border;
	rm->clip_max = 1 - border;            // RayMap::set_border (RayMap.h:66)
	rm->map_line_limit = rays_casted_res; // RayMap::set_ray_limit as main.cpp:777 calls it

	// frustum corners, eye, vanishing point (RayMap.h:107-141)
	V3 p[6] = { v3(1, 1, 1), v3(-1, 1, 1), v3(-1, -1, 1), v3(1, -1, 1), v3(0, 0, 0), v3(0, 0, 0) };
	M44 m;
	ident(m);
	rotate_x(m, rot[0]);
	rotate_y(m, rot[1]);
	m.m[3][0] += 3.0f; m.m[3][1] += 2.0f; m.m[3][2] += 0.0f; // translate(3,2,0)
	for (int i = 0; i < 5; i++) p[i] = xform(m, p[i]);

	const V3 down = v3(0, -1, 0);
	const V3 view = (p[0] + p[2]) / 2 - p[4];
	float ang;
	{	// vec3f::angle (VecMath.h:71-81), this = down, v = view
		const float d = view.x * down.x + view.y * down.y + view.z * down.z;
		float len = length(view) * length(down);
		if (len == 0) len = 0.00001f;
		float in = d / len;
		if (in < -1) in = -1;
		if (in > 1) in = 1;
		ang = std::acos(in);
	}
	const float alpha = float(M_PI) / 2 - ang;
	const float scale = 1 / std::sin(alpha);
	p[5] = p[4] + down * scale;

	// frustum -> unit square matrix and its inverse (RayMap.h:145-161)
	const V3 nrm = cross(p[1] - p[0], p[3] - p[0]);
	V3 d1 = p[1] - p[0], d2 = p[3] - p[0], d3 = nrm, d4 = p[0];
	d1 = d1 * (1 / (d1.x * d1.x + d1.y * d1.y + d1.z * d1.z));
	d2 = d2 * (1 / (d2.x * d2.x + d2.y * d2.y + d2.z * d2.z));
	d3 = d3 * (1 / (d3.x * d3.x + d3.y * d3.y + d3.z * d3.z));
	M44 to2d;
	const V3 rows[4] = { d1, d2, d3, d4 };
	for (int i = 0; i < 4; i++)
	{
		to2d.m[i][0] = rows[i].x; to2d.m[i][1] = rows[i].y; to2d.m[i][2] = rows[i].z;
		to2d.m[i][3] = 1.0f; // _vector4(const _vector3&) sets w = 1 (R/inc/mathlib/_vector4.h:120-127)
	}
	std::memcpy(rm->to3d, to2d.m, sizeof(to2d.m));
	invert_simpler(to2d);

	V3 p2[8];
	for (int i = 0; i < 8; i++) p2[i] = in3(rm->p_2d[i]);
	for (int i = 0; i < 6; i++) p2[i] = xform(to2d, p[i]);
	rm->vanishing_point_2d = out3(p2[5]);

	const int maxres = rays_casted_res / 4;
	rm->maxres = maxres;
	const int safety = 2;
	const float ys_min = border, ys_max = 1 - border;
	const V3 plist[4] = { v3(-1, -1, 0), v3(1, -1, 0), v3(1, 1, 0), v3(-1, 1, 0) };
	const V3 plist2[4] = { v3(0, ys_min, 0), v3(1, ys_min, 0), v3(1, ys_max, 0), v3(0, ys_max, 0) };
	int res[4] = { 0, 0, 0, 0 };
	p2[5].z = 0;
	const V3 vp = p2[5];
	V3 pn[8];
	for (int i = 0; i < 8; i++) pn[i] = in3(rm->p_no[i]);

	// helper for the recurring "re-aim this border point at screen corner c" expression:
	// vp + (c - vp) * abs(num / den)
	#define AIM(c, num, den) (vp + ((c) - vp) * std::abs((num) / (den)))

	if (vp.y > border) // upper part
	{
		const float e = std::abs(vp.y - border);
		pn[0] = vp + plist[0] * e;
		pn[1] = vp + plist[1] * e;
		const V3 in = pn[1];
		if (vp.x > 1) { if (pn[1].x > 1) pn[1].x = 1; }
		else if (dot(pn[0] - vp, plist2[2] - vp) > 0) pn[1] = AIM(plist2[2], plist2[0].y - vp.y, plist2[2].y - vp.y);
		if (vp.x < 0) { if (pn[0].x < 0) pn[0].x = 0; }
		else if (dot(in - vp, plist2[3] - vp) > 0) pn[0] = AIM(plist2[3], plist2[1].y - vp.y, plist2[3].y - vp.y);
		pn[0].y = pn[1].y = plist2[0].y;
		pn[0].x = float(int(maxres * pn[0].x) - safety) / maxres;
		pn[1].x = float(int(maxres * pn[1].x) + safety) / maxres;
		res[0] = maxres * std::abs(pn[0].x - pn[1].x);
		if (pn[0].x - pn[1].x > 0) res[0] = 0;
		if (res[0] > maxres * 3) res[0] = maxres * 3;
		rm->p_ofs_min[0] = pn[0].x;
		rm->p_ofs_max[0] = pn[1].x;
	}
	if (vp.y < 1 - border) // lower part
	{
		const float e = std::abs(vp.y - 1 + border);
		pn[2] = vp + plist[3] * e;
		pn[3] = vp + plist[2] * e;
		const V3 in = pn[2];
		if (vp.x < 0) { if (pn[2].x < 0) pn[2].x = 0; }
		else if (dot(pn[3] - vp, plist2[0] - vp) > 0) pn[2] = AIM(plist2[0], plist2[2].y - vp.y, plist2[0].y - vp.y);
		if (vp.x > 1) { if (pn[3].x > 1) pn[3].x = 1; }
		else
		{
			const V3 delta = plist2[1] - vp;
			if (dot(in - vp, delta) > 0) pn[3] = vp + delta * std::abs((plist2[3].y - vp.y) / delta.y);
		}
		pn[2].y = pn[3].y = plist2[2].y;
		pn[2].x = float(int(maxres * pn[2].x) - safety) / maxres;
		pn[3].x = float(int(maxres * pn[3].x) + safety) / maxres;
		res[1] = maxres * std::abs(pn[2].x - pn[3].x);
		if (pn[2].x - pn[3].x > 0) res[1] = 0;
		if (res[1] > maxres * 3) res[1] = maxres * 3;
		rm->p_ofs_min[1] = pn[2].x;
		rm->p_ofs_max[1] = pn[3].x;
	}
	if (vp.x > 0) // left part
	{
		const float e = std::abs(vp.x);
		pn[4] = vp + plist[0] * e;
		pn[5] = vp + plist[3] * e;
		const V3 in = pn[5];
		if (vp.y > 1 - border) { if (pn[5].y > 1 - border) pn[5].y = 1 - border; }
		else if (dot(pn[4] - vp, plist2[2] - vp) > 0) pn[5] = AIM(plist2[2], plist2[3].x - vp.x, plist2[2].x - vp.x);
		if (vp.y < border) { if (pn[4].y < border) pn[4].y = border; }
		else if (dot(in - vp, plist2[1] - vp) > 0) pn[4] = AIM(plist2[1], plist2[0].x - vp.x, plist2[1].x - vp.x);
		pn[4].x = pn[5].x = 0;
		pn[4].y = float(int(maxres * pn[4].y) - safety) / maxres;
		pn[5].y = float(int(maxres * pn[5].y) + safety) / maxres;
		res[2] = maxres * std::abs(pn[4].y - pn[5].y);
		if (pn[4].y - pn[5].y > 0) res[2] = 0;
		if (res[2] > maxres * 3) res[2] = maxres * 3;
		rm->p_ofs_min[2] = pn[4].y;
		rm->p_ofs_max[2] = pn[5].y;
	}
	if (vp.x < 1) // right part
	{
		const float e = std::abs(1 - vp.x);
		pn[6] = vp + plist[1] * e;
		pn[7] = vp + plist[2] * e;
		const V3 in = pn[7];
		if (vp.y > 1 - border) { if (pn[7].y > 1 - border) pn[7].y = 1 - border; }
		else if (dot(pn[6] - vp, plist2[3] - vp) > 0) pn[7] = AIM(plist2[3], plist2[2].x - vp.x, plist2[3].x - vp.x);
		if (vp.y < border) { if (pn[6].y < border) pn[6].y = border; }
		else if (dot(in - vp, plist2[0] - vp) > 0) pn[6] = AIM(plist2[0], plist2[1].x - vp.x, plist2[0].x - vp.x);
		pn[6].x = pn[7].x = 1;
		pn[6].y = float(int(maxres * pn[6].y) - safety) / maxres;
		pn[7].y = float(int(maxres * pn[7].y) + safety) / maxres;
		res[3] = maxres * std::abs(pn[6].y - pn[7].y);
		if (pn[6].y - pn[7].y > 0) res[3] = 0;
		if (res[3] > maxres * 3) res[3] = maxres * 3;
		rm->p_ofs_min[3] = pn[6].y;
		rm->p_ofs_max[3] = pn[7].y;
	}
	#undef AIM

	for (int i = 0; i < 8; i++) { rm->p_2d[i] = out3(p2[i]); rm->p_no[i] = out3(pn[i]); }
	for (int i = 0; i < 4; i++) rm->res[i] = res[i];
	rm->p4 = out3(p[4]);
	rm->map_line_count = res[0] + res[1] + res[2] + res[3];
}

} // namespace rlerc
