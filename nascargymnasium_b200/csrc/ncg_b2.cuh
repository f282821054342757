// ncg_b2.cuh -- the rigid-body half of one car-step: what box2d-py's b2World.Step(1/60, 6, 4) does to a
// single dynamic box among static wall boxes (call site /root/reference/src/car_physics.py:363), written
// for one GPU thread per car with fixed-size state.  float32 throughout, like Box2D.
//
// Specialisations that keep the arithmetic of the general algorithm: body B of every contact is a static
// wall (invMassB = invIB = 0, vB = wB = 0), so those terms are dropped (x + 0*y == x); every polygon is a
// 4-vertex box; the only moving broad-phase proxy is the car's, so pairs come from a uniform-grid query.
#pragma once
#include "ncg_defs.cuh"

namespace ncg {

struct V2 { float x, y; };
NCG_HD V2 mk(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
NCG_HD V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
NCG_HD V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
NCG_HD V2 operator-(V2 a) { return mk(-a.x, -a.y); }
NCG_HD V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
NCG_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
NCG_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
NCG_HD V2 cross(V2 a, float s) { return mk(s * a.y, -s * a.x); }
NCG_HD V2 cross(float s, V2 a) { return mk(-s * a.y, s * a.x); }
NCG_HD float length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
NCG_HD float normalize(V2& v) {
    float len = length(v);
    if (len < NCG_B2_EPS) return 0.0f;
    float inv = 1.0f / len;
    v.x *= inv; v.y *= inv;
    return len;
}
NCG_HD float fminb(float a, float b) { return a < b ? a : b; }
NCG_HD float fmaxb(float a, float b) { return a > b ? a : b; }
NCG_HD float clampb(float a, float lo, float hi) { return fmaxb(lo, fminb(a, hi)); }

struct Rot { float s, c; };
// sin and cos of a heading.  The library sincosf carries its large-argument path (Payne-Hanek, ~100 instructions and a
// local array) inline at every call site; the physics warp then jumps over it twice per step and lands on a cold
// instruction-cache line each time.  Headings stay small (a car turns a few radians per lap), so: three-term Cody-Waite
// reduction by pi/2 with FMAs (exact to ~1e-11 rad for |a| < 4.8e4), the Cephes single-precision minimax polynomials on
// [-pi/4, pi/4] (~1 ulp), quadrant fix-up with selects; anything larger goes to the library out of line.  sin(0) = 0
// and cos(0) = 1 exactly, which keeps the axis-aligned sensor rays of the start pose axis-aligned.
NCG_HDN void sincos_large(float a, float* s, float* c) { sincosf(a, s, c); }
NCG_HD void sincos_heading(float a, float* sp, float* cp) {
    if (!(fabsf(a) < 48000.0f)) { sincos_large(a, sp, cp); return; }
    const float q = rintf(a * 0.63661977236758138f);
    float r = fmaf(q, -1.5707963705062866f, a);             // pi/2 = 1.5707963705062866 - 4.371138828673793e-8 - 1.7151245100058819e-15 - 1e-23
    r = fmaf(q, 4.371138828673793e-8f, r);
    r = fmaf(q, 1.7151245100058819e-15f, r);
    const float z = r * r;
    const float sn = fmaf(r * z, fmaf(z, fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), r);
    const float cs = fmaf(z * z, fmaf(z, fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f), fmaf(z, -0.5f, 1.0f));
    const int k = (int)q;
    const float s1 = (k & 1) ? cs : sn, c1 = (k & 1) ? sn : cs;
    *sp = (k & 2) ? -s1 : s1;
    *cp = ((k + 1) & 2) ? -c1 : c1;
}
NCG_HD Rot rot(float a) { Rot q; sincos_heading(a, &q.s, &q.c); return q; }
struct Xf { V2 p; Rot q; };
NCG_HD V2 mul(Rot q, V2 v) { return mk(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
NCG_HD V2 mulT(Rot q, V2 v) { return mk(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
NCG_HD V2 mul(const Xf& T, V2 v) { return mk((T.q.c * v.x - T.q.s * v.y) + T.p.x, (T.q.s * v.x + T.q.c * v.y) + T.p.y); }
NCG_HD V2 mulT(const Xf& T, V2 v) {
    float px = v.x - T.p.x, py = v.y - T.p.y;
    return mk(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
NCG_HD Xf mulT(const Xf& A, const Xf& B) {
    Xf C;
    C.q.s = A.q.c * B.q.s - A.q.s * B.q.c;
    C.q.c = A.q.c * B.q.c + A.q.s * B.q.s;
    C.p = mulT(A.q, B.p - A.p);
    return C;
}

// A box polygon (b2PolygonShape::SetAsBox): vertex/normal i from the half extents.
struct Box { float hx, hy; };
NCG_HD V2 box_v(const Box& b, int i) { return mk((i == 1 || i == 2) ? b.hx : -b.hx, (i >= 2) ? b.hy : -b.hy); }
NCG_HD V2 box_n(int i) { return mk(i == 1 ? 1.0f : (i == 3 ? -1.0f : 0.0f), i == 2 ? 1.0f : (i == 0 ? -1.0f : 0.0f)); }
NCG_HD int box_support(const Box& b, V2 d) {
    int best = 0; float bv = dot(box_v(b, 0), d);
#pragma unroll
    for (int i = 1; i < 4; ++i) { float val = dot(box_v(b, i), d); if (val > bv) { best = i; bv = val; } }
    return best;
}
struct AABB { float lx, ly, ux, uy; };
NCG_HD AABB box_aabb(const Box& b, const Xf& xf) {
    V2 lo = mul(xf, box_v(b, 0)), hi = lo;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        V2 w = mul(xf, box_v(b, i));
        lo = mk(fminb(lo.x, w.x), fminb(lo.y, w.y)); hi = mk(fmaxb(hi.x, w.x), fmaxb(hi.y, w.y));
    }
    AABB a; a.lx = lo.x - NCG_B2_POLY_RADIUS; a.ly = lo.y - NCG_B2_POLY_RADIUS; a.ux = hi.x + NCG_B2_POLY_RADIUS; a.uy = hi.y + NCG_B2_POLY_RADIUS;
    return a;
}
NCG_HD bool aabb_contains(const AABB& a, const AABB& b) { return a.lx <= b.lx && a.ly <= b.ly && b.ux <= a.ux && b.uy <= a.uy; }
NCG_HD bool aabb_overlap(const AABB& a, const AABB& b) {
    if (b.lx - a.ux > 0.0f || b.ly - a.uy > 0.0f) return false;
    if (a.lx - b.ux > 0.0f || a.ly - b.uy > 0.0f) return false;
    return true;
}

struct Sweep { V2 c0, c; float a0, a, alpha0; };   // localCenter == 0 for the car and the walls
NCG_HD Xf sweep_xf(const Sweep& s, float beta) {
    Xf xf;
    xf.p = (1.0f - beta) * s.c0 + beta * s.c;
    float angle = (1.0f - beta) * s.a0 + beta * s.a;
    xf.q = rot(angle);
    xf.p = xf.p - mul(xf.q, mk(0.0f, 0.0f));
    return xf;
}
NCG_HD void sweep_advance(Sweep& s, float alpha) {
    float beta = (alpha - s.alpha0) / (1.0f - s.alpha0);
    s.c0 = s.c0 + beta * (s.c - s.c0);
    s.a0 += beta * (s.a - s.a0);
    s.alpha0 = alpha;
}
NCG_HD void sweep_normalize(Sweep& s) {
    float twoPi = 2.0f * NCG_B2_PI;
    float d = twoPi * floorf(s.a0 / twoPi);
    s.a0 -= d; s.a -= d;
}

// ------------------------------------------------------------------ b2CollidePolygons (box vs box)
enum { FACE_A = 1, FACE_B = 2 };
struct Manifold {
    V2 lp[2]; float ni[2], ti[2]; uint32_t key[2];
    V2 localNormal, localPoint; int type, pc;
};
NCG_HD uint32_t mk_key(int ia, int ib, int ta, int tb) {
    return (uint32_t)(ia & 255) | ((uint32_t)(ib & 255) << 8) | ((uint32_t)(ta & 255) << 16) | ((uint32_t)(tb & 255) << 24);
}
struct ClipV { V2 v; int ia, ib, ta, tb; };

NCG_HD float find_max_separation(int* edge, const Box& b1, const Xf& xf1, const Box& b2, const Xf& xf2) {
    Xf xf = mulT(xf2, xf1);
    int best = 0; float maxSep = -NCG_B2_MAXFLOAT;
    for (int i = 0; i < 4; ++i) {
        V2 n = mul(xf.q, box_n(i));
        V2 v1 = mul(xf, box_v(b1, i));
        float si = NCG_B2_MAXFLOAT;
#pragma unroll
        for (int j = 0; j < 4; ++j) { float sij = dot(n, box_v(b2, j) - v1); if (sij < si) si = sij; }
        if (si > maxSep) { maxSep = si; best = i; }
    }
    *edge = best; return maxSep;
}
// Box2D 2.3.0's form of the same search (b2EdgeSeparation + hill climb from the edge facing the other centroid): the one
// place the 2.3.x releases differ for this world, selectable because which one box2d-py 2.3.8 bundles cannot be checked
// offline (NcgConfig.contacts = 2; oracle/b2lite.h g_collide_variant; profiles/r02_b2_version_study.json)
NCG_HD float edge_separation230(const Box& b1, const Xf& xf1, int edge1, const Box& b2, const Xf& xf2) {
    V2 n1w = mul(xf1.q, box_n(edge1));
    V2 n1 = mulT(xf2.q, n1w);
    int index = 0; float minDot = NCG_B2_MAXFLOAT;
    for (int i = 0; i < 4; ++i) { float d = dot(box_v(b2, i), n1); if (d < minDot) { minDot = d; index = i; } }
    V2 v1 = mul(xf1, box_v(b1, edge1)), v2 = mul(xf2, box_v(b2, index));
    return dot(v2 - v1, n1w);
}
NCG_HDN float find_max_separation230(int* edgeOut, const Box& b1, const Xf& xf1, const Box& b2, const Xf& xf2) {
    V2 d = mul(xf2, mk(0.0f, 0.0f)) - mul(xf1, mk(0.0f, 0.0f));
    V2 dl = mulT(xf1.q, d);
    int edge = 0; float maxDot = -NCG_B2_MAXFLOAT;
    for (int i = 0; i < 4; ++i) { float dt = dot(box_n(i), dl); if (dt > maxDot) { maxDot = dt; edge = i; } }
    float s = edge_separation230(b1, xf1, edge, b2, xf2);
    int prevEdge = edge - 1 >= 0 ? edge - 1 : 3;
    float sPrev = edge_separation230(b1, xf1, prevEdge, b2, xf2);
    int nextEdge = edge + 1 < 4 ? edge + 1 : 0;
    float sNext = edge_separation230(b1, xf1, nextEdge, b2, xf2);
    int bestEdge, increment; float bestSep;
    if (sPrev > s && sPrev > sNext) { increment = -1; bestEdge = prevEdge; bestSep = sPrev; }
    else if (sNext > s) { increment = 1; bestEdge = nextEdge; bestSep = sNext; }
    else { *edgeOut = edge; return s; }
    for (;;) {
        edge = increment == -1 ? (bestEdge - 1 >= 0 ? bestEdge - 1 : 3) : (bestEdge + 1 < 4 ? bestEdge + 1 : 0);
        s = edge_separation230(b1, xf1, edge, b2, xf2);
        if (s > bestSep) { bestEdge = edge; bestSep = s; } else break;
    }
    *edgeOut = bestEdge; return bestSep;
}
NCG_HD int clip_segment(ClipV out[2], const ClipV in[2], V2 normal, float offset, int vertexIndexA) {
    int n = 0;
    float d0 = dot(normal, in[0].v) - offset, d1 = dot(normal, in[1].v) - offset;
    if (d0 <= 0.0f) out[n++] = in[0];
    if (d1 <= 0.0f) out[n++] = in[1];
    if (d0 * d1 < 0.0f) {
        float interp = d0 / (d0 - d1);
        out[n].v = in[0].v + interp * (in[1].v - in[0].v);
        out[n].ia = vertexIndexA; out[n].ib = in[0].ib; out[n].ta = 0; out[n].tb = 1;
        ++n;
    }
    return n;
}
// v230: false = b2CollidePolygons of Box2D 2.3.1 and later (the default), true = of 2.3.0
// *sep (optional) receives a lower bound of the distance between the two boxes at these poses: the separation along the
// face normal on which the test ended (a separating axis never overestimates the distance).
NCG_HD void collide_boxes(Manifold* m, const Box& bA, const Xf& xfA, const Box& bB, const Xf& xfB, bool v230 = false, float* sep = nullptr) {
    m->pc = 0;
    const float totalRadius = NCG_B2_POLY_RADIUS + NCG_B2_POLY_RADIUS;
    int edgeA = 0; float sepA = v230 ? find_max_separation230(&edgeA, bA, xfA, bB, xfB) : find_max_separation(&edgeA, bA, xfA, bB, xfB);
    if (sep) *sep = sepA;
    if (sepA > totalRadius) return;
    int edgeB = 0; float sepB = v230 ? find_max_separation230(&edgeB, bB, xfB, bA, xfA) : find_max_separation(&edgeB, bB, xfB, bA, xfA);
    if (sep) *sep = fmaxb(sepA, sepB);
    if (sepB > totalRadius) return;
    Box b1, b2; Xf xf1, xf2; int edge1, flip;
    const float k_tol = 0.1f * NCG_B2_LINEAR_SLOP;
    if (v230 ? (sepB > 0.98f * sepA + 0.001f) : (sepB > sepA + k_tol)) { b1 = bB; b2 = bA; xf1 = xfB; xf2 = xfA; edge1 = edgeB; m->type = FACE_B; flip = 1; }
    else { b1 = bA; b2 = bB; xf1 = xfA; xf2 = xfB; edge1 = edgeA; m->type = FACE_A; flip = 0; }
    ClipV inc[2];
    {   // b2FindIncidentEdge
        V2 normal1 = mulT(xf2.q, mul(xf1.q, box_n(edge1)));
        int index = 0; float minDot = NCG_B2_MAXFLOAT;
#pragma unroll
        for (int i = 0; i < 4; ++i) { float d = dot(normal1, box_n(i)); if (d < minDot) { minDot = d; index = i; } }
        int i1 = index, i2 = i1 + 1 < 4 ? i1 + 1 : 0;
        inc[0].v = mul(xf2, box_v(b2, i1)); inc[0].ia = edge1; inc[0].ib = i1; inc[0].ta = 1; inc[0].tb = 0;
        inc[1].v = mul(xf2, box_v(b2, i2)); inc[1].ia = edge1; inc[1].ib = i2; inc[1].ta = 1; inc[1].tb = 0;
    }
    int iv1 = edge1, iv2 = edge1 + 1 < 4 ? edge1 + 1 : 0;
    V2 v11 = box_v(b1, iv1), v12 = box_v(b1, iv2);
    V2 localTangent = v12 - v11; normalize(localTangent);
    V2 localNormal = cross(localTangent, 1.0f);
    V2 planePoint = 0.5f * (v11 + v12);
    V2 tangent = mul(xf1.q, localTangent);
    V2 normal = cross(tangent, 1.0f);
    v11 = mul(xf1, v11); v12 = mul(xf1, v12);
    float frontOffset = dot(normal, v11);
    float sideOffset1 = -dot(tangent, v11) + totalRadius;
    float sideOffset2 = dot(tangent, v12) + totalRadius;
    ClipV c1[2], c2[2];
    int np = clip_segment(c1, inc, -tangent, sideOffset1, iv1);
    if (np < 2) return;
    np = clip_segment(c2, c1, tangent, sideOffset2, iv2);
    if (np < 2) return;
    m->localNormal = localNormal; m->localPoint = planePoint;
    int pc = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float separation = dot(normal, c2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            m->lp[pc] = mulT(xf2, c2[i].v);
            m->key[pc] = flip ? mk_key(c2[i].ib, c2[i].ia, c2[i].tb, c2[i].ta) : mk_key(c2[i].ia, c2[i].ib, c2[i].ta, c2[i].tb);
            m->ni[pc] = 0.0f; m->ti[pc] = 0.0f;
            ++pc;
        }
    }
    m->pc = pc;
}

// b2WorldManifold::Initialize (radiusA == radiusB == polygon radius)
NCG_HD void world_manifold(V2* normal, V2 pts[2], const Manifold& m, const Xf& xfA, const Xf& xfB) {
    const float rA = NCG_B2_POLY_RADIUS, rB = NCG_B2_POLY_RADIUS;
    if (m.type == FACE_A) {
        V2 n = mul(xfA.q, m.localNormal);
        V2 planePoint = mul(xfA, m.localPoint);
        for (int i = 0; i < m.pc; ++i) {
            V2 clip = mul(xfB, m.lp[i]);
            V2 cA = clip + (rA - dot(clip - planePoint, n)) * n;
            V2 cB = clip - rB * n;
            pts[i] = 0.5f * (cA + cB);
        }
        *normal = n;
    } else {
        V2 n = mul(xfB.q, m.localNormal);
        V2 planePoint = mul(xfB, m.localPoint);
        for (int i = 0; i < m.pc; ++i) {
            V2 clip = mul(xfA, m.lp[i]);
            V2 cB = clip + (rB - dot(clip - planePoint, n)) * n;
            V2 cA = clip - rA * n;
            pts[i] = 0.5f * (cA + cB);
        }
        *normal = -n;
    }
}

// ------------------------------------------------------------------ b2Distance (GJK), useRadii = false
struct SimplexCache { float metric; int count; int ia[3], ib[3]; };
struct SVertex { V2 wA, wB, w; float a; int ia, ib; };
struct Simplex { SVertex v[3]; int count; };
NCG_HD float simplex_metric(const Simplex& s) {
    if (s.count == 2) return length(s.v[0].w - s.v[1].w);
    if (s.count == 3) return cross(s.v[1].w - s.v[0].w, s.v[2].w - s.v[0].w);
    return 0.0f;
}
NCG_HD void simplex_solve2(Simplex& s) {
    V2 w1 = s.v[0].w, w2 = s.v[1].w, e12 = w2 - w1;
    float d12_2 = -dot(w1, e12);
    if (d12_2 <= 0.0f) { s.v[0].a = 1.0f; s.count = 1; return; }
    float d12_1 = dot(w2, e12);
    if (d12_1 <= 0.0f) { s.v[1].a = 1.0f; s.count = 1; s.v[0] = s.v[1]; return; }
    float inv = 1.0f / (d12_1 + d12_2);
    s.v[0].a = d12_1 * inv; s.v[1].a = d12_2 * inv; s.count = 2;
}
NCG_HD void simplex_solve3(Simplex& s) {
    V2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;
    V2 e12 = w2 - w1; float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12); float d12_1 = w2e12, d12_2 = -w1e12;
    V2 e13 = w3 - w1; float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13); float d13_1 = w3e13, d13_2 = -w1e13;
    V2 e23 = w3 - w2; float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23); float d23_1 = w3e23, d23_2 = -w2e23;
    float n123 = cross(e12, e13);
    float d123_1 = n123 * cross(w2, w3), d123_2 = n123 * cross(w3, w1), d123_3 = n123 * cross(w1, w2);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) { s.v[0].a = 1.0f; s.count = 1; return; }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) { float inv = 1.0f / (d12_1 + d12_2); s.v[0].a = d12_1 * inv; s.v[1].a = d12_2 * inv; s.count = 2; return; }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) { float inv = 1.0f / (d13_1 + d13_2); s.v[0].a = d13_1 * inv; s.v[2].a = d13_2 * inv; s.count = 2; s.v[1] = s.v[2]; return; }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) { s.v[1].a = 1.0f; s.count = 1; s.v[0] = s.v[1]; return; }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) { s.v[2].a = 1.0f; s.count = 1; s.v[0] = s.v[2]; return; }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) { float inv = 1.0f / (d23_1 + d23_2); s.v[1].a = d23_1 * inv; s.v[2].a = d23_2 * inv; s.count = 2; s.v[0] = s.v[2]; return; }
    float inv = 1.0f / (d123_1 + d123_2 + d123_3);
    s.v[0].a = d123_1 * inv; s.v[1].a = d123_2 * inv; s.v[2].a = d123_3 * inv; s.count = 3;
}
NCG_HD float gjk_distance(SimplexCache* cache, const Box& bA, const Xf& xfA, const Box& bB, const Xf& xfB) {
    Simplex s;
    s.count = cache->count;
    for (int i = 0; i < s.count; ++i) {
        SVertex& v = s.v[i];
        v.ia = cache->ia[i]; v.ib = cache->ib[i];
        v.wA = mul(xfA, box_v(bA, v.ia)); v.wB = mul(xfB, box_v(bB, v.ib)); v.w = v.wB - v.wA; v.a = 0.0f;
    }
    if (s.count > 1) {
        float metric1 = cache->metric, metric2 = simplex_metric(s);
        if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < NCG_B2_EPS) s.count = 0;
    }
    if (s.count == 0) {
        SVertex& v = s.v[0];
        v.ia = 0; v.ib = 0; v.wA = mul(xfA, box_v(bA, 0)); v.wB = mul(xfB, box_v(bB, 0)); v.w = v.wB - v.wA; v.a = 1.0f; s.count = 1;
    }
    int saveA[3], saveB[3], saveCount = 0, iter = 0;
    while (iter < 20) {
        saveCount = s.count;
        for (int i = 0; i < saveCount; ++i) { saveA[i] = s.v[i].ia; saveB[i] = s.v[i].ib; }
        if (s.count == 2) simplex_solve2(s); else if (s.count == 3) simplex_solve3(s);
        if (s.count == 3) break;
        V2 d;
        if (s.count == 1) d = -s.v[0].w;
        else {
            V2 e12 = s.v[1].w - s.v[0].w;
            float sgn = cross(e12, -s.v[0].w);
            d = sgn > 0.0f ? cross(1.0f, e12) : cross(e12, 1.0f);
        }
        if (dot(d, d) < NCG_B2_EPS * NCG_B2_EPS) break;
        SVertex& vx = s.v[s.count];
        vx.ia = box_support(bA, mulT(xfA.q, -d)); vx.wA = mul(xfA, box_v(bA, vx.ia));
        vx.ib = box_support(bB, mulT(xfB.q, d)); vx.wB = mul(xfB, box_v(bB, vx.ib));
        vx.w = vx.wB - vx.wA;
        ++iter;
        bool dup = false;
        for (int i = 0; i < saveCount; ++i) if (vx.ia == saveA[i] && vx.ib == saveB[i]) { dup = true; break; }
        if (dup) break;
        ++s.count;
    }
    V2 pa, pb;
    if (s.count == 1) { pa = s.v[0].wA; pb = s.v[0].wB; }
    else if (s.count == 2) { pa = s.v[0].a * s.v[0].wA + s.v[1].a * s.v[1].wA; pb = s.v[0].a * s.v[0].wB + s.v[1].a * s.v[1].wB; }
    else { pa = s.v[0].a * s.v[0].wA + s.v[1].a * s.v[1].wA + s.v[2].a * s.v[2].wA; pb = pa; }
    float dist = length(pa - pb);
    cache->metric = simplex_metric(s); cache->count = s.count;
    for (int i = 0; i < s.count; ++i) { cache->ia[i] = s.v[i].ia; cache->ib[i] = s.v[i].ib; }
    return dist;
}

// ------------------------------------------------------------------ b2TimeOfImpact
enum { TOI_UNKNOWN, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };
struct SepFn { Box bA, bB; Sweep sA, sB; int type; V2 localPoint, axis; };
NCG_HD void sep_init(SepFn& f, const SimplexCache& cache, float t1) {
    Xf xfA = sweep_xf(f.sA, t1), xfB = sweep_xf(f.sB, t1);
    if (cache.count == 1) {
        f.type = 0;
        V2 pointA = mul(xfA, box_v(f.bA, cache.ia[0])), pointB = mul(xfB, box_v(f.bB, cache.ib[0]));
        f.axis = pointB - pointA; normalize(f.axis);
    } else if (cache.ia[0] == cache.ia[1]) {
        f.type = 2;
        V2 b1 = box_v(f.bB, cache.ib[0]), b2 = box_v(f.bB, cache.ib[1]);
        f.axis = cross(b2 - b1, 1.0f); normalize(f.axis);
        V2 normal = mul(xfB.q, f.axis);
        f.localPoint = 0.5f * (b1 + b2);
        V2 pointB = mul(xfB, f.localPoint), pointA = mul(xfA, box_v(f.bA, cache.ia[0]));
        float s = dot(pointA - pointB, normal);
        if (s < 0.0f) f.axis = -f.axis;
    } else {
        f.type = 1;
        V2 a1 = box_v(f.bA, cache.ia[0]), a2 = box_v(f.bA, cache.ia[1]);
        f.axis = cross(a2 - a1, 1.0f); normalize(f.axis);
        V2 normal = mul(xfA.q, f.axis);
        f.localPoint = 0.5f * (a1 + a2);
        V2 pointA = mul(xfA, f.localPoint), pointB = mul(xfB, box_v(f.bB, cache.ib[0]));
        float s = dot(pointB - pointA, normal);
        if (s < 0.0f) f.axis = -f.axis;
    }
}
NCG_HD float sep_find_min(const SepFn& f, int* ia, int* ib, float t) {
    Xf xfA = sweep_xf(f.sA, t), xfB = sweep_xf(f.sB, t);
    if (f.type == 0) {
        V2 axisA = mulT(xfA.q, f.axis), axisB = mulT(xfB.q, -f.axis);
        *ia = box_support(f.bA, axisA); *ib = box_support(f.bB, axisB);
        V2 pointA = mul(xfA, box_v(f.bA, *ia)), pointB = mul(xfB, box_v(f.bB, *ib));
        return dot(pointB - pointA, f.axis);
    } else if (f.type == 1) {
        V2 normal = mul(xfA.q, f.axis), pointA = mul(xfA, f.localPoint);
        V2 axisB = mulT(xfB.q, -normal);
        *ia = -1; *ib = box_support(f.bB, axisB);
        V2 pointB = mul(xfB, box_v(f.bB, *ib));
        return dot(pointB - pointA, normal);
    } else {
        V2 normal = mul(xfB.q, f.axis), pointB = mul(xfB, f.localPoint);
        V2 axisA = mulT(xfA.q, -normal);
        *ib = -1; *ia = box_support(f.bA, axisA);
        V2 pointA = mul(xfA, box_v(f.bA, *ia));
        return dot(pointA - pointB, normal);
    }
}
NCG_HD float sep_eval(const SepFn& f, int ia, int ib, float t) {
    Xf xfA = sweep_xf(f.sA, t), xfB = sweep_xf(f.sB, t);
    if (f.type == 0) {
        V2 pointA = mul(xfA, box_v(f.bA, ia)), pointB = mul(xfB, box_v(f.bB, ib));
        return dot(pointB - pointA, f.axis);
    } else if (f.type == 1) {
        V2 normal = mul(xfA.q, f.axis), pointA = mul(xfA, f.localPoint), pointB = mul(xfB, box_v(f.bB, ib));
        return dot(pointB - pointA, normal);
    } else {
        V2 normal = mul(xfB.q, f.axis), pointB = mul(xfB, f.localPoint), pointA = mul(xfA, box_v(f.bA, ia));
        return dot(pointA - pointB, normal);
    }
}
NCG_HD void time_of_impact(int* state, float* tOut, const Box& bA, Sweep sweepA, const Box& bB, Sweep sweepB, float tMax) {
    *state = TOI_UNKNOWN; *tOut = tMax;
    sweep_normalize(sweepA); sweep_normalize(sweepB);
    const float totalRadius = NCG_B2_POLY_RADIUS + NCG_B2_POLY_RADIUS;
    const float target = fmaxb(NCG_B2_LINEAR_SLOP, totalRadius - 3.0f * NCG_B2_LINEAR_SLOP);
    const float tolerance = 0.25f * NCG_B2_LINEAR_SLOP;
    float t1 = 0.0f; int iter = 0;
    SimplexCache cache; cache.count = 0; cache.metric = 0.0f;
    SepFn fcn; fcn.bA = bA; fcn.bB = bB; fcn.sA = sweepA; fcn.sB = sweepB;
    for (;;) {
        Xf xfA = sweep_xf(sweepA, t1), xfB = sweep_xf(sweepB, t1);
        float dist = gjk_distance(&cache, bA, xfA, bB, xfB);
        if (dist <= 0.0f) { *state = TOI_OVERLAPPED; *tOut = 0.0f; break; }
        if (dist < target + tolerance) { *state = TOI_TOUCHING; *tOut = t1; break; }
        sep_init(fcn, cache, t1);
        bool done = false; float t2 = tMax; int pushBackIter = 0;
        for (;;) {
            int ia, ib;
            float s2 = sep_find_min(fcn, &ia, &ib, t2);
            if (s2 > target + tolerance) { *state = TOI_SEPARATED; *tOut = tMax; done = true; break; }
            if (s2 > target - tolerance) { t1 = t2; break; }
            float s1 = sep_eval(fcn, ia, ib, t1);
            if (s1 < target - tolerance) { *state = TOI_FAILED; *tOut = t1; done = true; break; }
            if (s1 <= target + tolerance) { *state = TOI_TOUCHING; *tOut = t1; done = true; break; }
            int rootIter = 0; float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIter & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                ++rootIter;
                float s = sep_eval(fcn, ia, ib, t);
                if (fabsf(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                if (rootIter == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == 8) break;
        }
        ++iter;
        if (done) break;
        if (iter == 20) { *state = TOI_FAILED; *tOut = t1; break; }
    }
}

}  // namespace ncg
