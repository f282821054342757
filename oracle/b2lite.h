// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// b2lite: a float32 CPU restatement of the subset of Box2D 2.3.x that
// NascarGymnasium's CarEnv reaches through box2d-py==2.3.8 (requirements.txt:4):
// one dynamic box (the car, /root/reference/src/car.py:200-239) against static
// wall boxes (/root/reference/src/car_physics.py:275-339), stepped by
// b2World.Step(1/60, 6, 4) (/root/reference/src/car_physics.py:363), ray cast by
// b2World.RayCast (/root/reference/src/distance_sensor.py:113) and observed by
// a b2ContactListener (/root/reference/src/car_physics.py:693-848).
//
// PARITY UNPINNED for this file: Box2D itself is not vendored in /root/reference
// and box2d-py is not installable in the build container, and the reference has
// no tests/golden vectors.  The algorithm below restates upstream Box2D v2.3.1+
// (b2World::Step/Solve/SolveTOI, b2Island::Solve/SolveTOI, b2ContactSolver,
// b2CollidePolygons, b2Distance, b2TimeOfImpact, b2PolygonShape::RayCast,
// b2DynamicTree::MoveProxy fat-AABB hysteresis) in upstream operation order.
// Simplifications that cannot change results for this world (one dynamic body,
// static walls, zero gravity/damping, no joints, no sensors, no bullets):
//   * the dynamic AABB tree is replaced by a linear scan over wall fat AABBs
//     (pair order = ascending wall index, which is what ascending proxy ids give);
//   * island building is the car plus its touching contacts, newest contact first
//     (contacts are head-inserted into the body's contact list).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace b2 {

typedef float f32;
static const f32 kPi = 3.14159265359f;
static const f32 kEps = 1.1920928955078125e-7f;   // FLT_EPSILON
static const f32 kMaxFloat = 3.402823466e+38f;
static const f32 kLinearSlop = 0.005f;
static const f32 kAngularSlop = 2.0f / 180.0f * kPi;
static const f32 kPolygonRadius = 2.0f * kLinearSlop;
static const f32 kAabbExtension = 0.1f;
static const f32 kAabbMultiplier = 2.0f;
static const f32 kVelocityThreshold = 1.0f;
static const f32 kBaumgarte = 0.2f;
static const f32 kToiBaumgarte = 0.75f;
static const f32 kMaxLinearCorrection = 0.2f;
static const f32 kMaxTranslation = 2.0f;
static const f32 kMaxTranslationSq = kMaxTranslation * kMaxTranslation;
static const f32 kMaxRotation = 0.5f * kPi;
static const f32 kMaxRotationSq = kMaxRotation * kMaxRotation;
static const int kMaxSubSteps = 8;
static const int kMaxTOIContacts = 32;
static const f32 kTimeToSleep = 0.5f;
static const f32 kLinearSleepTol = 0.01f;
static const f32 kAngularSleepTol = 2.0f / 180.0f * kPi;
static const int kMaxPolygonVertices = 8;

struct V2 { f32 x, y; };
static inline V2 mk(f32 x, f32 y) { V2 v; v.x = x; v.y = y; return v; }
static inline V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
static inline V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
static inline V2 operator-(V2 a) { return mk(-a.x, -a.y); }
static inline V2 operator*(f32 s, V2 a) { return mk(s * a.x, s * a.y); }
static inline f32 dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline f32 cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline V2 cross(V2 a, f32 s) { return mk(s * a.y, -s * a.x); }
static inline V2 cross(f32 s, V2 a) { return mk(-s * a.y, s * a.x); }
static inline f32 length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
static inline f32 lengthSq(V2 a) { return a.x * a.x + a.y * a.y; }
static inline f32 normalize(V2& v) {
    f32 len = length(v);
    if (len < kEps) return 0.0f;
    f32 inv = 1.0f / len;
    v.x *= inv; v.y *= inv;
    return len;
}
static inline f32 fmin2(f32 a, f32 b) { return a < b ? a : b; }
static inline f32 fmax2(f32 a, f32 b) { return a > b ? a : b; }
static inline f32 clampf(f32 a, f32 lo, f32 hi) { return fmax2(lo, fmin2(a, hi)); }
static inline V2 vmin(V2 a, V2 b) { return mk(fmin2(a.x, b.x), fmin2(a.y, b.y)); }
static inline V2 vmax(V2 a, V2 b) { return mk(fmax2(a.x, b.x), fmax2(a.y, b.y)); }

struct Rot { f32 s, c; void set(f32 a) { s = sinf(a); c = cosf(a); } };
struct Xf { V2 p; Rot q; };
static inline V2 mul(Rot q, V2 v) { return mk(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
static inline V2 mulT(Rot q, V2 v) { return mk(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
static inline V2 mul(const Xf& T, V2 v) {
    return mk((T.q.c * v.x - T.q.s * v.y) + T.p.x, (T.q.s * v.x + T.q.c * v.y) + T.p.y);
}
static inline V2 mulT(const Xf& T, V2 v) {
    f32 px = v.x - T.p.x, py = v.y - T.p.y;
    return mk(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
static inline Rot mulT(Rot q, Rot r) { Rot o; o.s = q.c * r.s - q.s * r.c; o.c = q.c * r.c + q.s * r.s; return o; }
static inline Xf mulT(const Xf& A, const Xf& B) { Xf C; C.q = mulT(A.q, B.q); C.p = mulT(A.q, B.p - A.p); return C; }

struct AABB { V2 lo, hi; };
static inline bool contains(const AABB& a, const AABB& b) {
    bool r = true;
    r = r && a.lo.x <= b.lo.x; r = r && a.lo.y <= b.lo.y;
    r = r && b.hi.x <= a.hi.x; r = r && b.hi.y <= a.hi.y;
    return r;
}
static inline bool overlap(const AABB& a, const AABB& b) {
    V2 d1 = b.lo - a.hi, d2 = a.lo - b.hi;
    if (d1.x > 0.0f || d1.y > 0.0f) return false;
    if (d2.x > 0.0f || d2.y > 0.0f) return false;
    return true;
}

// Box polygon as b2PolygonShape::SetAsBox(hx, hy) builds it.
struct Poly {
    V2 v[4]; V2 n[4]; f32 radius;
    void setAsBox(f32 hx, f32 hy) {
        v[0] = mk(-hx, -hy); v[1] = mk(hx, -hy); v[2] = mk(hx, hy); v[3] = mk(-hx, hy);
        n[0] = mk(0.0f, -1.0f); n[1] = mk(1.0f, 0.0f); n[2] = mk(0.0f, 1.0f); n[3] = mk(-1.0f, 0.0f);
        radius = kPolygonRadius;
    }
    AABB computeAABB(const Xf& xf) const {
        V2 lo = mul(xf, v[0]), hi = lo;
        for (int i = 1; i < 4; ++i) { V2 w = mul(xf, v[i]); lo = vmin(lo, w); hi = vmax(hi, w); }
        V2 r = mk(radius, radius);
        AABB a; a.lo = lo - r; a.hi = hi + r; return a;
    }
    int support(V2 d) const {
        int best = 0; f32 bv = dot(v[0], d);
        for (int i = 1; i < 4; ++i) { f32 val = dot(v[i], d); if (val > bv) { best = i; bv = val; } }
        return best;
    }
    // b2PolygonShape::RayCast
    bool rayCast(V2 P1, V2 P2, f32 maxFraction, const Xf& xf, f32* fraction) const {
        V2 p1 = mulT(xf.q, P1 - xf.p);
        V2 p2 = mulT(xf.q, P2 - xf.p);
        V2 d = p2 - p1;
        f32 lower = 0.0f, upper = maxFraction;
        int index = -1;
        for (int i = 0; i < 4; ++i) {
            f32 numerator = dot(n[i], v[i] - p1);
            f32 denominator = dot(n[i], d);
            if (denominator == 0.0f) {
                if (numerator < 0.0f) return false;
            } else {
                if (denominator < 0.0f && numerator < lower * denominator) {
                    lower = numerator / denominator; index = i;
                } else if (denominator > 0.0f && numerator < upper * denominator) {
                    upper = numerator / denominator;
                }
            }
            if (upper < lower) return false;
        }
        if (index >= 0) { *fraction = lower; return true; }
        return false;
    }
    bool testPoint(const Xf& xf, V2 p) const {
        V2 pl = mulT(xf.q, p - xf.p);
        for (int i = 0; i < 4; ++i) { if (dot(n[i], pl - v[i]) > 0.0f) return false; }
        return true;
    }
};

struct Sweep {
    V2 localCenter, c0, c; f32 a0, a, alpha0;
    void getTransform(Xf* xf, f32 beta) const {
        xf->p = (1.0f - beta) * c0 + beta * c;
        f32 angle = (1.0f - beta) * a0 + beta * a;
        xf->q.set(angle);
        xf->p = xf->p - mul(xf->q, localCenter);
    }
    void advance(f32 alpha) {
        f32 beta = (alpha - alpha0) / (1.0f - alpha0);
        c0 = c0 + beta * (c - c0);
        a0 += beta * (a - a0);
        alpha0 = alpha;
    }
    void normalize() {
        f32 twoPi = 2.0f * kPi;
        f32 d = twoPi * floorf(a0 / twoPi);
        a0 -= d; a -= d;
    }
};

// ------------------------------------------------------------------ manifold
enum { FACE_A = 1, FACE_B = 2 };
struct MPoint { V2 localPoint; f32 normalImpulse, tangentImpulse; uint32_t key; };
struct Manifold { MPoint points[2]; V2 localNormal, localPoint; int type, pointCount; };
static inline uint32_t mkKey(int indexA, int indexB, int typeA, int typeB) {
    return (uint32_t)(indexA & 255) | ((uint32_t)(indexB & 255) << 8) | ((uint32_t)(typeA & 255) << 16) |
           ((uint32_t)(typeB & 255) << 24);
}
struct ClipVertex { V2 v; int indexA, indexB, typeA, typeB; };   // type: 0 vertex, 1 face

// Which upstream minor's b2CollidePolygons is restated -- the one place where 2.3.x releases differ for this world:
//   0 = v2.3.1 and later: b2FindMaxSeparation is a brute-force max over all edge normals, evaluated in poly2's frame,
//       and the reference face flips to B when separationB > separationA + 0.1 * b2_linearSlop;
//   1 = v2.3.0: b2EdgeSeparation per edge normal (support vertex of poly2, separation measured in world coordinates),
//       b2FindMaxSeparation hill-climbs from the edge whose normal points most towards poly2's centroid, and the flip
//       test is separationB > 0.98 * separationA + 0.001 (k_relativeTol / k_absoluteTol).
// Which of the two box2d-py 2.3.8 bundles cannot be checked offline; tools/b2_version_study.py measures how often the
// two disagree on this world's contacts, and tests/test_b2_variants.py keeps the goldens identical under both.
static int g_collide_variant = 0;

static inline f32 edgeSeparation230(const Poly& p1, const Xf& xf1, int edge1, const Poly& p2, const Xf& xf2) {
    V2 normal1World = mul(xf1.q, p1.n[edge1]);
    V2 normal1 = mulT(xf2.q, normal1World);
    int index = 0; f32 minDot = kMaxFloat;
    for (int i = 0; i < 4; ++i) { f32 d = dot(p2.v[i], normal1); if (d < minDot) { minDot = d; index = i; } }
    V2 v1 = mul(xf1, p1.v[edge1]);
    V2 v2 = mul(xf2, p2.v[index]);
    return dot(v2 - v1, normal1World);
}
static inline f32 findMaxSeparation230(int* edgeIndex, const Poly& p1, const Xf& xf1, const Poly& p2, const Xf& xf2) {
    // vector from the centroid of poly1 to the centroid of poly2 (SetAsBox: both centroids are the body origins)
    V2 d = mul(xf2, mk(0.0f, 0.0f)) - mul(xf1, mk(0.0f, 0.0f));
    V2 dLocal1 = mulT(xf1.q, d);
    int edge = 0; f32 maxDot = -kMaxFloat;
    for (int i = 0; i < 4; ++i) { f32 dt = dot(p1.n[i], dLocal1); if (dt > maxDot) { maxDot = dt; edge = i; } }
    f32 s = edgeSeparation230(p1, xf1, edge, p2, xf2);
    int prevEdge = edge - 1 >= 0 ? edge - 1 : 3;
    f32 sPrev = edgeSeparation230(p1, xf1, prevEdge, p2, xf2);
    int nextEdge = edge + 1 < 4 ? edge + 1 : 0;
    f32 sNext = edgeSeparation230(p1, xf1, nextEdge, p2, xf2);
    int bestEdge; f32 bestSeparation; int increment;
    if (sPrev > s && sPrev > sNext) { increment = -1; bestEdge = prevEdge; bestSeparation = sPrev; }
    else if (sNext > s) { increment = 1; bestEdge = nextEdge; bestSeparation = sNext; }
    else { *edgeIndex = edge; return s; }
    for (;;) {
        if (increment == -1) edge = bestEdge - 1 >= 0 ? bestEdge - 1 : 3;
        else edge = bestEdge + 1 < 4 ? bestEdge + 1 : 0;
        s = edgeSeparation230(p1, xf1, edge, p2, xf2);
        if (s > bestSeparation) { bestEdge = edge; bestSeparation = s; } else break;
    }
    *edgeIndex = bestEdge; return bestSeparation;
}
static inline f32 findMaxSeparation(int* edgeIndex, const Poly& p1, const Xf& xf1, const Poly& p2, const Xf& xf2) {
    if (g_collide_variant == 1) return findMaxSeparation230(edgeIndex, p1, xf1, p2, xf2);
    Xf xf = mulT(xf2, xf1);
    int bestIndex = 0; f32 maxSep = -kMaxFloat;
    for (int i = 0; i < 4; ++i) {
        V2 n = mul(xf.q, p1.n[i]);
        V2 v1 = mul(xf, p1.v[i]);
        f32 si = kMaxFloat;
        for (int j = 0; j < 4; ++j) { f32 sij = dot(n, p2.v[j] - v1); if (sij < si) si = sij; }
        if (si > maxSep) { maxSep = si; bestIndex = i; }
    }
    *edgeIndex = bestIndex; return maxSep;
}
static inline void findIncidentEdge(ClipVertex c[2], const Poly& p1, const Xf& xf1, int edge1, const Poly& p2,
                                    const Xf& xf2) {
    V2 normal1 = mulT(xf2.q, mul(xf1.q, p1.n[edge1]));
    int index = 0; f32 minDot = kMaxFloat;
    for (int i = 0; i < 4; ++i) { f32 d = dot(normal1, p2.n[i]); if (d < minDot) { minDot = d; index = i; } }
    int i1 = index, i2 = i1 + 1 < 4 ? i1 + 1 : 0;
    c[0].v = mul(xf2, p2.v[i1]); c[0].indexA = edge1; c[0].indexB = i1; c[0].typeA = 1; c[0].typeB = 0;
    c[1].v = mul(xf2, p2.v[i2]); c[1].indexA = edge1; c[1].indexB = i2; c[1].typeA = 1; c[1].typeB = 0;
}
static inline int clipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], V2 normal, f32 offset, int vertexIndexA) {
    int numOut = 0;
    f32 d0 = dot(normal, vIn[0].v) - offset;
    f32 d1 = dot(normal, vIn[1].v) - offset;
    if (d0 <= 0.0f) vOut[numOut++] = vIn[0];
    if (d1 <= 0.0f) vOut[numOut++] = vIn[1];
    if (d0 * d1 < 0.0f) {
        f32 interp = d0 / (d0 - d1);
        vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
        vOut[numOut].indexA = vertexIndexA; vOut[numOut].indexB = vIn[0].indexB;
        vOut[numOut].typeA = 0; vOut[numOut].typeB = 1;
        ++numOut;
    }
    return numOut;
}
// b2CollidePolygons (variant 0: v2.3.1+, brute-force max separation, k_tol = 0.1*linearSlop; variant 1: v2.3.0)
static inline void collidePolygons(Manifold* m, const Poly& polyA, const Xf& xfA, const Poly& polyB, const Xf& xfB) {
    m->pointCount = 0;
    f32 totalRadius = polyA.radius + polyB.radius;
    int edgeA = 0; f32 sepA = findMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
    if (sepA > totalRadius) return;
    int edgeB = 0; f32 sepB = findMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
    if (sepB > totalRadius) return;
    const Poly *poly1, *poly2; Xf xf1, xf2; int edge1; int flip;
    const f32 k_tol = 0.1f * kLinearSlop;
    const bool flipToB = g_collide_variant == 1 ? (sepB > 0.98f * sepA + 0.001f) : (sepB > sepA + k_tol);
    if (flipToB) { poly1 = &polyB; poly2 = &polyA; xf1 = xfB; xf2 = xfA; edge1 = edgeB; m->type = FACE_B; flip = 1; }
    else { poly1 = &polyA; poly2 = &polyB; xf1 = xfA; xf2 = xfB; edge1 = edgeA; m->type = FACE_A; flip = 0; }
    ClipVertex incident[2];
    findIncidentEdge(incident, *poly1, xf1, edge1, *poly2, xf2);
    int iv1 = edge1, iv2 = edge1 + 1 < 4 ? edge1 + 1 : 0;
    V2 v11 = poly1->v[iv1], v12 = poly1->v[iv2];
    V2 localTangent = v12 - v11; normalize(localTangent);
    V2 localNormal = cross(localTangent, 1.0f);
    V2 planePoint = 0.5f * (v11 + v12);
    V2 tangent = mul(xf1.q, localTangent);
    V2 normal = cross(tangent, 1.0f);
    v11 = mul(xf1, v11); v12 = mul(xf1, v12);
    f32 frontOffset = dot(normal, v11);
    f32 sideOffset1 = -dot(tangent, v11) + totalRadius;
    f32 sideOffset2 = dot(tangent, v12) + totalRadius;
    ClipVertex cp1[2], cp2[2];
    int np = clipSegmentToLine(cp1, incident, -tangent, sideOffset1, iv1);
    if (np < 2) return;
    np = clipSegmentToLine(cp2, cp1, tangent, sideOffset2, iv2);
    if (np < 2) return;
    m->localNormal = localNormal; m->localPoint = planePoint;
    int pc = 0;
    for (int i = 0; i < 2; ++i) {
        f32 separation = dot(normal, cp2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            MPoint* cp = m->points + pc;
            cp->localPoint = mulT(xf2, cp2[i].v);
            if (flip) cp->key = mkKey(cp2[i].indexB, cp2[i].indexA, cp2[i].typeB, cp2[i].typeA);
            else cp->key = mkKey(cp2[i].indexA, cp2[i].indexB, cp2[i].typeA, cp2[i].typeB);
            cp->normalImpulse = 0.0f; cp->tangentImpulse = 0.0f;
            ++pc;
        }
    }
    m->pointCount = pc;
}

struct WorldManifold {
    V2 normal, points[2];
    void initialize(const Manifold* m, const Xf& xfA, f32 rA, const Xf& xfB, f32 rB) {
        if (m->pointCount == 0) return;
        if (m->type == FACE_A) {
            normal = mul(xfA.q, m->localNormal);
            V2 planePoint = mul(xfA, m->localPoint);
            for (int i = 0; i < m->pointCount; ++i) {
                V2 clipPoint = mul(xfB, m->points[i].localPoint);
                V2 cA = clipPoint + (rA - dot(clipPoint - planePoint, normal)) * normal;
                V2 cB = clipPoint - rB * normal;
                points[i] = 0.5f * (cA + cB);
            }
        } else {
            normal = mul(xfB.q, m->localNormal);
            V2 planePoint = mul(xfB, m->localPoint);
            for (int i = 0; i < m->pointCount; ++i) {
                V2 clipPoint = mul(xfA, m->points[i].localPoint);
                V2 cB = clipPoint + (rB - dot(clipPoint - planePoint, normal)) * normal;
                V2 cA = clipPoint - rA * normal;
                points[i] = 0.5f * (cA + cB);
            }
            normal = -normal;
        }
    }
};

// ------------------------------------------------------------------ b2Distance (GJK)
struct SimplexCache { f32 metric; int count; int indexA[3], indexB[3]; };
struct SimplexVertex { V2 wA, wB, w; f32 a; int indexA, indexB; };
struct Simplex {
    SimplexVertex v[3]; int count;
    f32 getMetric() const {
        if (count == 2) return length(v[0].w - v[1].w);
        if (count == 3) return cross(v[1].w - v[0].w, v[2].w - v[0].w);
        return 0.0f;
    }
    void readCache(const SimplexCache* cache, const Poly& pA, const Xf& xfA, const Poly& pB, const Xf& xfB) {
        count = cache->count;
        for (int i = 0; i < count; ++i) {
            SimplexVertex* s = v + i;
            s->indexA = cache->indexA[i]; s->indexB = cache->indexB[i];
            s->wA = mul(xfA, pA.v[s->indexA]); s->wB = mul(xfB, pB.v[s->indexB]);
            s->w = s->wB - s->wA; s->a = 0.0f;
        }
        if (count > 1) {
            f32 metric1 = cache->metric, metric2 = getMetric();
            if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEps) count = 0;
        }
        if (count == 0) {
            SimplexVertex* s = v;
            s->indexA = 0; s->indexB = 0;
            s->wA = mul(xfA, pA.v[0]); s->wB = mul(xfB, pB.v[0]);
            s->w = s->wB - s->wA; s->a = 1.0f; count = 1;
        }
    }
    void writeCache(SimplexCache* cache) const {
        cache->metric = getMetric(); cache->count = count;
        for (int i = 0; i < count; ++i) { cache->indexA[i] = v[i].indexA; cache->indexB[i] = v[i].indexB; }
    }
    V2 searchDirection() const {
        if (count == 1) return -v[0].w;
        V2 e12 = v[1].w - v[0].w;
        f32 sgn = cross(e12, -v[0].w);
        if (sgn > 0.0f) return cross(1.0f, e12);
        return cross(e12, 1.0f);
    }
    void witnessPoints(V2* pA, V2* pB) const {
        if (count == 1) { *pA = v[0].wA; *pB = v[0].wB; }
        else if (count == 2) { *pA = v[0].a * v[0].wA + v[1].a * v[1].wA; *pB = v[0].a * v[0].wB + v[1].a * v[1].wB; }
        else { *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA; *pB = *pA; }
    }
    void solve2() {
        V2 w1 = v[0].w, w2 = v[1].w, e12 = w2 - w1;
        f32 d12_2 = -dot(w1, e12);
        if (d12_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
        f32 d12_1 = dot(w2, e12);
        if (d12_1 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
        f32 inv = 1.0f / (d12_1 + d12_2);
        v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2;
    }
    void solve3() {
        V2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
        V2 e12 = w2 - w1; f32 w1e12 = dot(w1, e12), w2e12 = dot(w2, e12); f32 d12_1 = w2e12, d12_2 = -w1e12;
        V2 e13 = w3 - w1; f32 w1e13 = dot(w1, e13), w3e13 = dot(w3, e13); f32 d13_1 = w3e13, d13_2 = -w1e13;
        V2 e23 = w3 - w2; f32 w2e23 = dot(w2, e23), w3e23 = dot(w3, e23); f32 d23_1 = w3e23, d23_2 = -w2e23;
        f32 n123 = cross(e12, e13);
        f32 d123_1 = n123 * cross(w2, w3), d123_2 = n123 * cross(w3, w1), d123_3 = n123 * cross(w1, w2);
        if (d12_2 <= 0.0f && d13_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
        if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
            f32 inv = 1.0f / (d12_1 + d12_2); v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2; return;
        }
        if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
            f32 inv = 1.0f / (d13_1 + d13_2); v[0].a = d13_1 * inv; v[2].a = d13_2 * inv; count = 2; v[1] = v[2]; return;
        }
        if (d12_1 <= 0.0f && d23_2 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
        if (d13_1 <= 0.0f && d23_1 <= 0.0f) { v[2].a = 1.0f; count = 1; v[0] = v[2]; return; }
        if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
            f32 inv = 1.0f / (d23_1 + d23_2); v[1].a = d23_1 * inv; v[2].a = d23_2 * inv; count = 2; v[0] = v[2]; return;
        }
        f32 inv = 1.0f / (d123_1 + d123_2 + d123_3);
        v[0].a = d123_1 * inv; v[1].a = d123_2 * inv; v[2].a = d123_3 * inv; count = 3;
    }
};
// b2Distance with useRadii = false; returns the distance, updates the cache.
static inline f32 distance(SimplexCache* cache, const Poly& pA, const Xf& xfA, const Poly& pB, const Xf& xfB) {
    Simplex s; s.readCache(cache, pA, xfA, pB, xfB);
    int saveA[3], saveB[3], saveCount = 0, iter = 0;
    while (iter < 20) {
        saveCount = s.count;
        for (int i = 0; i < saveCount; ++i) { saveA[i] = s.v[i].indexA; saveB[i] = s.v[i].indexB; }
        if (s.count == 2) s.solve2(); else if (s.count == 3) s.solve3();
        if (s.count == 3) break;
        V2 d = s.searchDirection();
        if (lengthSq(d) < kEps * kEps) break;
        SimplexVertex* vx = s.v + s.count;
        vx->indexA = pA.support(mulT(xfA.q, -d)); vx->wA = mul(xfA, pA.v[vx->indexA]);
        vx->indexB = pB.support(mulT(xfB.q, d)); vx->wB = mul(xfB, pB.v[vx->indexB]);
        vx->w = vx->wB - vx->wA;
        ++iter;
        bool dup = false;
        for (int i = 0; i < saveCount; ++i) if (vx->indexA == saveA[i] && vx->indexB == saveB[i]) { dup = true; break; }
        if (dup) break;
        ++s.count;
    }
    V2 a, b; s.witnessPoints(&a, &b);
    f32 dist = length(a - b);
    s.writeCache(cache);
    return dist;
}

// ------------------------------------------------------------------ b2TimeOfImpact
enum { TOI_UNKNOWN, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };
struct SepFn {
    const Poly *pA, *pB; Sweep sA, sB; int type; V2 localPoint, axis;   // type 0 points, 1 faceA, 2 faceB
    f32 initialize(const SimplexCache* cache, const Poly* A, const Sweep& swA, const Poly* B, const Sweep& swB, f32 t1) {
        pA = A; pB = B; sA = swA; sB = swB;
        Xf xfA, xfB; sA.getTransform(&xfA, t1); sB.getTransform(&xfB, t1);
        if (cache->count == 1) {
            type = 0;
            V2 pointA = mul(xfA, pA->v[cache->indexA[0]]), pointB = mul(xfB, pB->v[cache->indexB[0]]);
            axis = pointB - pointA; return normalize(axis);
        } else if (cache->indexA[0] == cache->indexA[1]) {
            type = 2;
            V2 b1 = pB->v[cache->indexB[0]], b2v = pB->v[cache->indexB[1]];
            axis = cross(b2v - b1, 1.0f); normalize(axis);
            V2 normal = mul(xfB.q, axis);
            localPoint = 0.5f * (b1 + b2v);
            V2 pointB = mul(xfB, localPoint), pointA = mul(xfA, pA->v[cache->indexA[0]]);
            f32 s = dot(pointA - pointB, normal);
            if (s < 0.0f) { axis = -axis; s = -s; }
            return s;
        } else {
            type = 1;
            V2 a1 = pA->v[cache->indexA[0]], a2 = pA->v[cache->indexA[1]];
            axis = cross(a2 - a1, 1.0f); normalize(axis);
            V2 normal = mul(xfA.q, axis);
            localPoint = 0.5f * (a1 + a2);
            V2 pointA = mul(xfA, localPoint), pointB = mul(xfB, pB->v[cache->indexB[0]]);
            f32 s = dot(pointB - pointA, normal);
            if (s < 0.0f) { axis = -axis; s = -s; }
            return s;
        }
    }
    f32 findMinSeparation(int* indexA, int* indexB, f32 t) const {
        Xf xfA, xfB; sA.getTransform(&xfA, t); sB.getTransform(&xfB, t);
        if (type == 0) {
            V2 axisA = mulT(xfA.q, axis), axisB = mulT(xfB.q, -axis);
            *indexA = pA->support(axisA); *indexB = pB->support(axisB);
            V2 pointA = mul(xfA, pA->v[*indexA]), pointB = mul(xfB, pB->v[*indexB]);
            return dot(pointB - pointA, axis);
        } else if (type == 1) {
            V2 normal = mul(xfA.q, axis), pointA = mul(xfA, localPoint);
            V2 axisB = mulT(xfB.q, -normal);
            *indexA = -1; *indexB = pB->support(axisB);
            V2 pointB = mul(xfB, pB->v[*indexB]);
            return dot(pointB - pointA, normal);
        } else {
            V2 normal = mul(xfB.q, axis), pointB = mul(xfB, localPoint);
            V2 axisA = mulT(xfA.q, -normal);
            *indexB = -1; *indexA = pA->support(axisA);
            V2 pointA = mul(xfA, pA->v[*indexA]);
            return dot(pointA - pointB, normal);
        }
    }
    f32 evaluate(int indexA, int indexB, f32 t) const {
        Xf xfA, xfB; sA.getTransform(&xfA, t); sB.getTransform(&xfB, t);
        if (type == 0) {
            V2 pointA = mul(xfA, pA->v[indexA]), pointB = mul(xfB, pB->v[indexB]);
            return dot(pointB - pointA, axis);
        } else if (type == 1) {
            V2 normal = mul(xfA.q, axis), pointA = mul(xfA, localPoint), pointB = mul(xfB, pB->v[indexB]);
            return dot(pointB - pointA, normal);
        } else {
            V2 normal = mul(xfB.q, axis), pointB = mul(xfB, localPoint), pointA = mul(xfA, pA->v[indexA]);
            return dot(pointA - pointB, normal);
        }
    }
};
static inline void timeOfImpact(int* state, f32* tOut, const Poly& pA, Sweep sweepA, const Poly& pB, Sweep sweepB, f32 tMax) {
    *state = TOI_UNKNOWN; *tOut = tMax;
    sweepA.normalize(); sweepB.normalize();
    f32 totalRadius = pA.radius + pB.radius;
    f32 target = fmax2(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    f32 tolerance = 0.25f * kLinearSlop;
    f32 t1 = 0.0f; int iter = 0;
    SimplexCache cache; cache.count = 0; cache.metric = 0.0f;
    for (;;) {
        Xf xfA, xfB; sweepA.getTransform(&xfA, t1); sweepB.getTransform(&xfB, t1);
        f32 dist = distance(&cache, pA, xfA, pB, xfB);
        if (dist <= 0.0f) { *state = TOI_OVERLAPPED; *tOut = 0.0f; break; }
        if (dist < target + tolerance) { *state = TOI_TOUCHING; *tOut = t1; break; }
        SepFn fcn; fcn.initialize(&cache, &pA, sweepA, &pB, sweepB, t1);
        bool done = false; f32 t2 = tMax; int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            f32 s2 = fcn.findMinSeparation(&indexA, &indexB, t2);
            if (s2 > target + tolerance) { *state = TOI_SEPARATED; *tOut = tMax; done = true; break; }
            if (s2 > target - tolerance) { t1 = t2; break; }
            f32 s1 = fcn.evaluate(indexA, indexB, t1);
            if (s1 < target - tolerance) { *state = TOI_FAILED; *tOut = t1; done = true; break; }
            if (s1 <= target + tolerance) { *state = TOI_TOUCHING; *tOut = t1; done = true; break; }
            int rootIter = 0; f32 a1 = t1, a2 = t2;
            for (;;) {
                f32 t;
                if (rootIter & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                ++rootIter;
                f32 s = fcn.evaluate(indexA, indexB, t);
                if (fabsf(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                if (rootIter == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == kMaxPolygonVertices) break;
        }
        ++iter;
        if (done) break;
        if (iter == 20) { *state = TOI_FAILED; *tOut = t1; break; }
    }
}

// ------------------------------------------------------------------ world
struct Wall { Xf xf; f32 angle; Poly poly; AABB fat; };

struct Contact {
    int wall; bool touching, enabled, toiFlag, islandFlag; int toiCount; f32 toi;
    Manifold m;
};

struct Listener {   // callbacks; implemented by the oracle's CarCollisionListener restatement
    virtual void beginContact(int wall, V2 normal) = 0;
    virtual void endContact(int wall) = 0;
    virtual void postSolve(int wall, int count, const f32* normalImpulses) = 0;
    virtual ~Listener() {}
};

struct VCPoint { V2 rA, rB; f32 normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias; };
struct VelocityConstraint {
    VCPoint points[2]; V2 normal; f32 nm[4]; f32 K[4];   // Mat22 stored ex.x, ex.y, ey.x, ey.y
    f32 friction, restitution; int pointCount; int contactIndex;
};
struct PositionConstraint { V2 localPoints[2], localNormal, localPoint; int type, pointCount; int wall; };

struct World {
    // the car body
    Poly carPoly; Xf xf; Sweep sweep; V2 v; f32 w; V2 force; f32 torque;
    f32 invMass, invI, sleepTime; bool awake;
    AABB carFat;
    std::vector<Wall> walls;
    std::vector<Contact> contacts;   // index 0 = head of the contact list (newest)
    f32 inv_dt0; bool stepComplete;
    f32 friction, restitution;
    Listener* listener;
    // per-wall alpha0 of the static sweeps touched during SolveTOI (all other walls are 0)
    std::vector<f32> wallAlpha0;

    World() : listener(nullptr) {}

    void createCar(f32 x, f32 y, f32 angle, f32 hx, f32 hy, f32 mass, f32 inertia, f32 fricCar, f32 restCar, f32 fricWall,
                   f32 restWall) {
        carPoly.setAsBox(hx, hy);
        xf.p = mk(x, y); xf.q.set(angle);
        sweep.localCenter = mk(0.0f, 0.0f); sweep.c0 = sweep.c = xf.p; sweep.a0 = sweep.a = angle; sweep.alpha0 = 0.0f;
        v = mk(0.0f, 0.0f); w = 0.0f; force = mk(0.0f, 0.0f); torque = 0.0f;
        invMass = 1.0f / mass;
        f32 I = inertia - mass * dot(sweep.localCenter, sweep.localCenter);
        invI = 1.0f / I;
        sleepTime = 0.0f; awake = true;
        proxyMoved = true; newFixture = true;   // proxies are buffered as moved at creation (e_newFixture)
        AABB a = carPoly.computeAABB(xf);
        V2 r = mk(kAabbExtension, kAabbExtension);
        carFat.lo = a.lo - r; carFat.hi = a.hi + r;
        inv_dt0 = 0.0f; stepComplete = true;
        friction = sqrtf(fricCar * fricWall);
        restitution = restCar > restWall ? restCar : restWall;
        contacts.clear();
    }
    void addWall(f32 px, f32 py, f32 angle, f32 hx, f32 hy) {
        Wall wl; wl.xf.p = mk(px, py); wl.xf.q.set(angle); wl.angle = angle; wl.poly.setAsBox(hx, hy);
        AABB a = wl.poly.computeAABB(wl.xf);
        V2 r = mk(kAabbExtension, kAabbExtension);
        wl.fat.lo = a.lo - r; wl.fat.hi = a.hi + r;
        walls.push_back(wl); wallAlpha0.push_back(0.0f);
    }
    // b2Body::SetAwake
    void setAwake(bool flag) {
        if (flag) { if (!awake) { awake = true; sleepTime = 0.0f; } }
        else { awake = false; sleepTime = 0.0f; v = mk(0.0f, 0.0f); w = 0.0f; force = mk(0.0f, 0.0f); torque = 0.0f; }
    }
    // b2Body::ApplyForce / ApplyForceToCenter / ApplyTorque with wake=True
    void applyForce(V2 f, V2 point) { if (!awake) setAwake(true); force = force + f; torque += cross(point - sweep.c, f); }
    void applyForceToCenter(V2 f) { if (!awake) setAwake(true); force = force + f; }
    void applyTorque(f32 t) { if (!awake) setAwake(true); torque += t; }
    V2 worldVector(V2 lv) const { return mul(xf.q, lv); }
    V2 worldPoint(V2 lp) const { return mul(xf, lp); }
    // b2Body::SetTransform (body.position = p; body.angle = a are two SetTransform calls, net effect below)
    void setTransform(V2 p, f32 angle) {
        xf.q.set(angle); xf.p = p;
        sweep.c = mul(xf, sweep.localCenter); sweep.a = angle; sweep.c0 = sweep.c; sweep.a0 = angle;
        AABB a1 = carPoly.computeAABB(xf);
        moveProxy(a1, mk(0.0f, 0.0f));
    }
    void setLinearVelocity(V2 nv) { if (dot(nv, nv) > 0.0f) setAwake(true); v = nv; }
    void setAngularVelocity(f32 nw) { if (nw * nw > 0.0f) setAwake(true); w = nw; }

    bool proxyMoved, newFixture;
    void moveProxy(const AABB& aabb, V2 displacement) {
        if (contains(carFat, aabb)) return;
        AABB b = aabb; V2 r = mk(kAabbExtension, kAabbExtension);
        b.lo = b.lo - r; b.hi = b.hi + r;
        V2 d = kAabbMultiplier * displacement;
        if (d.x < 0.0f) b.lo.x += d.x; else b.hi.x += d.x;
        if (d.y < 0.0f) b.lo.y += d.y; else b.hi.y += d.y;
        carFat = b; proxyMoved = true;
    }
    void synchronizeTransform() { xf.q.set(sweep.a); xf.p = sweep.c - mul(xf.q, sweep.localCenter); }
    void synchronizeFixtures() {
        Xf xf1; xf1.q.set(sweep.a0); xf1.p = sweep.c0 - mul(xf1.q, sweep.localCenter);
        AABB a1 = carPoly.computeAABB(xf1), a2 = carPoly.computeAABB(xf);
        AABB c; c.lo = vmin(a1.lo, a2.lo); c.hi = vmax(a1.hi, a2.hi);
        moveProxy(c, xf.p - xf1.p);
    }
    void advanceBody(f32 alpha) {
        sweep.advance(alpha); sweep.c = sweep.c0; sweep.a = sweep.a0;
        xf.q.set(sweep.a); xf.p = sweep.c - mul(xf.q, sweep.localCenter);
    }
    bool hasContact(int wall) const { for (auto& c : contacts) if (c.wall == wall) return true; return false; }
    // b2ContactManager::FindNewContacts -> b2BroadPhase::UpdatePairs (only the car proxy ever moves)
    void findNewContacts() {
        if (!proxyMoved) return;
        proxyMoved = false;
        for (int i = 0; i < (int)walls.size(); ++i) {
            if (!overlap(carFat, walls[i].fat)) continue;
            if (hasContact(i)) continue;
            Contact c; c.wall = i; c.touching = false; c.enabled = true; c.toiFlag = false; c.islandFlag = false;
            c.toiCount = 0; c.toi = 1.0f; c.m.pointCount = 0; c.m.type = FACE_A;
            contacts.insert(contacts.begin(), c);
        }
    }
    // b2Contact::Update
    void updateContact(Contact& c) {
        Manifold old = c.m;
        c.enabled = true;
        bool wasTouching = c.touching;
        const Wall& wl = walls[c.wall];
        collidePolygons(&c.m, carPoly, xf, wl.poly, wl.xf);
        bool touching = c.m.pointCount > 0;
        for (int i = 0; i < c.m.pointCount; ++i) {
            MPoint* mp2 = c.m.points + i; mp2->normalImpulse = 0.0f; mp2->tangentImpulse = 0.0f;
            for (int j = 0; j < old.pointCount; ++j) {
                if (old.points[j].key == mp2->key) {
                    mp2->normalImpulse = old.points[j].normalImpulse; mp2->tangentImpulse = old.points[j].tangentImpulse; break;
                }
            }
        }
        if (touching != wasTouching) setAwake(true);
        c.touching = touching;
        if (!wasTouching && touching && listener) {
            WorldManifold wm; wm.initialize(&c.m, xf, carPoly.radius, wl.xf, wl.poly.radius);
            listener->beginContact(c.wall, wm.normal);
        }
        if (wasTouching && !touching && listener) listener->endContact(c.wall);
    }
    // b2ContactManager::Collide
    void collide() {
        for (size_t i = 0; i < contacts.size();) {
            Contact& c = contacts[i];
            if (!awake) { ++i; continue; }
            if (!overlap(carFat, walls[c.wall].fat)) {
                if (c.touching && listener) listener->endContact(c.wall);
                contacts.erase(contacts.begin() + i);
                continue;
            }
            updateContact(c);
            ++i;
        }
    }

    // ---- contact solver over the island {car} + contacts[idx...]
    std::vector<VelocityConstraint> vcs; std::vector<PositionConstraint> pcs; std::vector<int> islandContacts;
    V2 pc_c; f32 pc_a; V2 pv; f32 pw;   // car position / velocity used by the solver

    void solverInit(bool warmStarting, f32 dtRatio) {
        vcs.resize(islandContacts.size()); pcs.resize(islandContacts.size());
        for (size_t i = 0; i < islandContacts.size(); ++i) {
            Contact& c = contacts[islandContacts[i]];
            VelocityConstraint& vc = vcs[i]; PositionConstraint& pc = pcs[i];
            vc.friction = friction; vc.restitution = restitution; vc.contactIndex = islandContacts[i];
            vc.pointCount = c.m.pointCount;
            for (int k = 0; k < 4; ++k) { vc.K[k] = 0.0f; vc.nm[k] = 0.0f; }
            pc.localNormal = c.m.localNormal; pc.localPoint = c.m.localPoint; pc.pointCount = c.m.pointCount;
            pc.type = c.m.type; pc.wall = c.wall;
            for (int j = 0; j < c.m.pointCount; ++j) {
                VCPoint& p = vc.points[j];
                if (warmStarting) { p.normalImpulse = dtRatio * c.m.points[j].normalImpulse; p.tangentImpulse = dtRatio * c.m.points[j].tangentImpulse; }
                else { p.normalImpulse = 0.0f; p.tangentImpulse = 0.0f; }
                p.rA = mk(0, 0); p.rB = mk(0, 0); p.normalMass = 0.0f; p.tangentMass = 0.0f; p.velocityBias = 0.0f;
                pc.localPoints[j] = c.m.points[j].localPoint;
            }
        }
    }
    void initializeVelocityConstraints() {
        for (size_t i = 0; i < vcs.size(); ++i) {
            VelocityConstraint& vc = vcs[i]; PositionConstraint& pc = pcs[i];
            const Wall& wl = walls[pc.wall];
            Contact& c = contacts[vc.contactIndex];
            f32 mA = invMass, mB = 0.0f, iA = invI, iB = 0.0f;
            V2 cA = pc_c; f32 aA = pc_a; V2 vA = pv; f32 wA = pw;
            V2 cB = wl.xf.p; f32 aB = wl.angle; V2 vB = mk(0, 0); f32 wB = 0.0f;
            Xf xfA, xfB; xfA.q.set(aA); xfB.q.set(aB);
            xfA.p = cA - mul(xfA.q, sweep.localCenter); xfB.p = cB - mul(xfB.q, mk(0.0f, 0.0f));
            WorldManifold wm; wm.initialize(&c.m, xfA, carPoly.radius, xfB, wl.poly.radius);
            vc.normal = wm.normal;
            for (int j = 0; j < vc.pointCount; ++j) {
                VCPoint& p = vc.points[j];
                p.rA = wm.points[j] - cA; p.rB = wm.points[j] - cB;
                f32 rnA = cross(p.rA, vc.normal), rnB = cross(p.rB, vc.normal);
                f32 kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                p.normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
                V2 tangent = cross(vc.normal, 1.0f);
                f32 rtA = cross(p.rA, tangent), rtB = cross(p.rB, tangent);
                f32 kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
                p.tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
                p.velocityBias = 0.0f;
                f32 vRel = dot(vc.normal, vB + cross(wB, p.rB) - vA - cross(wA, p.rA));
                if (vRel < -kVelocityThreshold) p.velocityBias = -vc.restitution * vRel;
            }
            if (vc.pointCount == 2) {
                VCPoint& p1 = vc.points[0]; VCPoint& p2 = vc.points[1];
                f32 rn1A = cross(p1.rA, vc.normal), rn1B = cross(p1.rB, vc.normal);
                f32 rn2A = cross(p2.rA, vc.normal), rn2B = cross(p2.rB, vc.normal);
                f32 k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
                f32 k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
                f32 k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
                const f32 k_maxConditionNumber = 1000.0f;
                if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
                    vc.K[0] = k11; vc.K[1] = k12; vc.K[2] = k12; vc.K[3] = k22;
                    f32 a = k11, b = k12, cc = k12, d = k22;
                    f32 det = a * d - b * cc;
                    if (det != 0.0f) det = 1.0f / det;
                    vc.nm[0] = det * d; vc.nm[2] = -det * b; vc.nm[1] = -det * cc; vc.nm[3] = det * a;
                } else {
                    vc.pointCount = 1;
                }
            }
        }
    }
    void warmStart() {
        for (auto& vc : vcs) {
            f32 mA = invMass, iA = invI;
            V2 tangent = cross(vc.normal, 1.0f);
            for (int j = 0; j < vc.pointCount; ++j) {
                VCPoint& p = vc.points[j];
                V2 P = p.normalImpulse * vc.normal + p.tangentImpulse * tangent;
                pw -= iA * cross(p.rA, P);
                pv = pv - mA * P;
            }
        }
    }
    void solveVelocityConstraints() {
        for (auto& vc : vcs) {
            f32 mA = invMass, iA = invI;
            V2 vA = pv; f32 wA = pw; V2 vB = mk(0, 0); f32 wB = 0.0f;
            V2 normal = vc.normal; V2 tangent = cross(normal, 1.0f); f32 fr = vc.friction;
            for (int j = 0; j < vc.pointCount; ++j) {
                VCPoint& p = vc.points[j];
                V2 dv = vB + cross(wB, p.rB) - vA - cross(wA, p.rA);
                f32 vt = dot(dv, tangent) - 0.0f;
                f32 lambda = p.tangentMass * (-vt);
                f32 maxFriction = fr * p.normalImpulse;
                f32 newImpulse = clampf(p.tangentImpulse + lambda, -maxFriction, maxFriction);
                lambda = newImpulse - p.tangentImpulse; p.tangentImpulse = newImpulse;
                V2 P = lambda * tangent;
                vA = vA - mA * P; wA -= iA * cross(p.rA, P);
            }
            if (vc.pointCount == 1) {
                VCPoint& p = vc.points[0];
                V2 dv = vB + cross(wB, p.rB) - vA - cross(wA, p.rA);
                f32 vn = dot(dv, normal);
                f32 lambda = -p.normalMass * (vn - p.velocityBias);
                f32 newImpulse = fmax2(p.normalImpulse + lambda, 0.0f);
                lambda = newImpulse - p.normalImpulse; p.normalImpulse = newImpulse;
                V2 P = lambda * normal;
                vA = vA - mA * P; wA -= iA * cross(p.rA, P);
            } else if (vc.pointCount == 2) {
                VCPoint& cp1 = vc.points[0]; VCPoint& cp2 = vc.points[1];
                V2 a = mk(cp1.normalImpulse, cp2.normalImpulse);
                V2 dv1 = vB + cross(wB, cp1.rB) - vA - cross(wA, cp1.rA);
                V2 dv2 = vB + cross(wB, cp2.rB) - vA - cross(wA, cp2.rA);
                f32 vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
                V2 b = mk(vn1 - cp1.velocityBias, vn2 - cp2.velocityBias);
                // b -= K a   (K columns: ex = (K0,K1), ey = (K2,K3))
                b = b - mk(vc.K[0] * a.x + vc.K[2] * a.y, vc.K[1] * a.x + vc.K[3] * a.y);
                for (;;) {
                    V2 x = -mk(vc.nm[0] * b.x + vc.nm[2] * b.y, vc.nm[1] * b.x + vc.nm[3] * b.y);
                    if (x.x >= 0.0f && x.y >= 0.0f) {
                        V2 d = x - a; V2 P1 = d.x * normal, P2 = d.y * normal;
                        vA = vA - mA * (P1 + P2); wA -= iA * (cross(cp1.rA, P1) + cross(cp2.rA, P2));
                        cp1.normalImpulse = x.x; cp2.normalImpulse = x.y; break;
                    }
                    x.x = -cp1.normalMass * b.x; x.y = 0.0f;
                    vn1 = 0.0f; vn2 = vc.K[1] * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) {
                        V2 d = x - a; V2 P1 = d.x * normal, P2 = d.y * normal;
                        vA = vA - mA * (P1 + P2); wA -= iA * (cross(cp1.rA, P1) + cross(cp2.rA, P2));
                        cp1.normalImpulse = x.x; cp2.normalImpulse = x.y; break;
                    }
                    x.x = 0.0f; x.y = -cp2.normalMass * b.y;
                    vn1 = vc.K[2] * x.y + b.x; vn2 = 0.0f;
                    if (x.y >= 0.0f && vn1 >= 0.0f) {
                        V2 d = x - a; V2 P1 = d.x * normal, P2 = d.y * normal;
                        vA = vA - mA * (P1 + P2); wA -= iA * (cross(cp1.rA, P1) + cross(cp2.rA, P2));
                        cp1.normalImpulse = x.x; cp2.normalImpulse = x.y; break;
                    }
                    x.x = 0.0f; x.y = 0.0f; vn1 = b.x; vn2 = b.y;
                    if (vn1 >= 0.0f && vn2 >= 0.0f) {
                        V2 d = x - a; V2 P1 = d.x * normal, P2 = d.y * normal;
                        vA = vA - mA * (P1 + P2); wA -= iA * (cross(cp1.rA, P1) + cross(cp2.rA, P2));
                        cp1.normalImpulse = x.x; cp2.normalImpulse = x.y; break;
                    }
                    break;
                }
            }
            pv = vA; pw = wA;
        }
    }
    void storeImpulses() {
        for (auto& vc : vcs) {
            Manifold& m = contacts[vc.contactIndex].m;
            for (int j = 0; j < vc.pointCount; ++j) {
                m.points[j].normalImpulse = vc.points[j].normalImpulse; m.points[j].tangentImpulse = vc.points[j].tangentImpulse;
            }
        }
    }
    // SolvePositionConstraints / SolveTOIPositionConstraints (the car is always a TOI body)
    bool solvePositionConstraints(f32 baumgarte, f32 okFactor) {
        f32 minSeparation = 0.0f;
        for (auto& pc : pcs) {
            const Wall& wl = walls[pc.wall];
            f32 mA = invMass, iA = invI, mB = 0.0f, iB = 0.0f;
            V2 cA = pc_c; f32 aA = pc_a; V2 cB = wl.xf.p; f32 aB = wl.angle;
            for (int j = 0; j < pc.pointCount; ++j) {
                Xf xfA, xfB; xfA.q.set(aA); xfB.q.set(aB);
                xfA.p = cA - mul(xfA.q, sweep.localCenter); xfB.p = cB - mul(xfB.q, mk(0.0f, 0.0f));
                V2 normal, point; f32 separation;
                if (pc.type == FACE_A) {
                    normal = mul(xfA.q, pc.localNormal);
                    V2 planePoint = mul(xfA, pc.localPoint);
                    V2 clipPoint = mul(xfB, pc.localPoints[j]);
                    separation = dot(clipPoint - planePoint, normal) - carPoly.radius - wl.poly.radius;
                    point = clipPoint;
                } else {
                    normal = mul(xfB.q, pc.localNormal);
                    V2 planePoint = mul(xfB, pc.localPoint);
                    V2 clipPoint = mul(xfA, pc.localPoints[j]);
                    separation = dot(clipPoint - planePoint, normal) - carPoly.radius - wl.poly.radius;
                    point = clipPoint;
                    normal = -normal;
                }
                V2 rA = point - cA, rB = point - cB;
                minSeparation = fmin2(minSeparation, separation);
                f32 C = clampf(baumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                f32 rnA = cross(rA, normal), rnB = cross(rB, normal);
                f32 K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                f32 impulse = K > 0.0f ? -C / K : 0.0f;
                V2 P = impulse * normal;
                cA = cA - mA * P; aA -= iA * cross(rA, P);
            }
            pc_c = cA; pc_a = aA;
        }
        return minSeparation >= okFactor * kLinearSlop;
    }
    void report() {
        if (!listener) return;
        for (auto& vc : vcs) {
            f32 ni[2] = {0.0f, 0.0f};
            for (int j = 0; j < vc.pointCount; ++j) ni[j] = vc.points[j].normalImpulse;
            listener->postSolve(contacts[vc.contactIndex].wall, vc.pointCount, ni);
        }
    }
    void integratePositions(f32 h) {
        V2 translation = h * pv;
        if (dot(translation, translation) > kMaxTranslationSq) { f32 ratio = kMaxTranslation / length(translation); pv = ratio * pv; }
        f32 rotation = h * pw;
        if (rotation * rotation > kMaxRotationSq) { f32 ratio = kMaxRotation / fabsf(rotation); pw *= ratio; }
        pc_c = pc_c + h * pv; pc_a += h * pw;
    }

    // b2World::Solve + b2Island::Solve
    void solve(f32 h, int velIters, int posIters, f32 dtRatio) {
        if (awake) {
            islandContacts.clear();
            for (size_t i = 0; i < contacts.size(); ++i) if (contacts[i].enabled && contacts[i].touching) islandContacts.push_back((int)i);
            pc_c = sweep.c; pc_a = sweep.a; pv = v; pw = w;
            sweep.c0 = sweep.c; sweep.a0 = sweep.a;
            pv = pv + h * (1.0f * mk(0.0f, 0.0f) + invMass * force);
            pw += h * invI * torque;
            pv = (1.0f / (1.0f + h * 0.0f)) * pv;
            pw *= 1.0f / (1.0f + h * 0.0f);
            solverInit(true, dtRatio);
            initializeVelocityConstraints();
            warmStart();
            for (int i = 0; i < velIters; ++i) solveVelocityConstraints();
            storeImpulses();
            integratePositions(h);
            bool positionSolved = false;
            for (int i = 0; i < posIters; ++i) {
                bool ok = solvePositionConstraints(kBaumgarte, -3.0f);
                if (ok) { positionSolved = true; break; }
            }
            sweep.c = pc_c; sweep.a = pc_a; v = pv; w = pw;
            synchronizeTransform();
            report();
            // sleep
            {
                f32 minSleepTime = kMaxFloat;
                const f32 linTolSqr = kLinearSleepTol * kLinearSleepTol, angTolSqr = kAngularSleepTol * kAngularSleepTol;
                if (w * w > angTolSqr || dot(v, v) > linTolSqr) { sleepTime = 0.0f; minSleepTime = 0.0f; }
                else { sleepTime += h; minSleepTime = fmin2(minSleepTime, sleepTime); }
                if (minSleepTime >= kTimeToSleep && positionSolved) setAwake(false);
            }
            synchronizeFixtures();
        }
        findNewContacts();
    }

    // b2World::SolveTOI + b2Island::SolveTOI
    void solveTOI(f32 stepDt, int velIters) {
        if (stepComplete) {
            sweep.alpha0 = 0.0f;
            for (auto& a : wallAlpha0) a = 0.0f;
            for (auto& c : contacts) { c.toiFlag = false; c.islandFlag = false; c.toiCount = 0; c.toi = 1.0f; }
        }
        for (;;) {
            int minContact = -1; f32 minAlpha = 1.0f;
            for (size_t i = 0; i < contacts.size(); ++i) {
                Contact& c = contacts[i];
                if (!c.enabled) continue;
                if (c.toiCount > kMaxSubSteps) continue;
                f32 alpha = 1.0f;
                if (c.toiFlag) alpha = c.toi;
                else {
                    if (!awake) continue;   // activeA false and the wall is static
                    f32 alpha0 = sweep.alpha0;
                    f32& wa = wallAlpha0[c.wall];
                    if (sweep.alpha0 < wa) { alpha0 = wa; sweep.advance(alpha0); }
                    else if (wa < sweep.alpha0) { alpha0 = sweep.alpha0; wa = alpha0; }
                    const Wall& wl = walls[c.wall];
                    Sweep sB; sB.localCenter = mk(0, 0); sB.c0 = sB.c = wl.xf.p; sB.a0 = sB.a = wl.angle; sB.alpha0 = wa;
                    int state; f32 t;
                    timeOfImpact(&state, &t, carPoly, sweep, wl.poly, sB, 1.0f);
                    f32 beta = t;
                    if (state == TOI_TOUCHING) alpha = fmin2(alpha0 + (1.0f - alpha0) * beta, 1.0f); else alpha = 1.0f;
                    c.toi = alpha; c.toiFlag = true;
                }
                if (alpha < minAlpha) { minContact = (int)i; minAlpha = alpha; }
            }
            if (minContact < 0 || 1.0f - 10.0f * kEps < minAlpha) { stepComplete = true; break; }
            Contact& mc = contacts[minContact];
            Sweep backup1 = sweep; f32 backupWall = wallAlpha0[mc.wall];
            advanceBody(minAlpha); wallAlpha0[mc.wall] = minAlpha;
            updateContact(mc);
            mc.toiFlag = false; ++mc.toiCount;
            if (!mc.enabled || !mc.touching) {
                mc.enabled = false; sweep = backup1; wallAlpha0[mc.wall] = backupWall; synchronizeTransform();
                continue;
            }
            setAwake(true);
            islandContacts.clear(); islandContacts.push_back(minContact);
            mc.islandFlag = true;
            for (size_t i = 0; i < contacts.size(); ++i) {
                if ((int)islandContacts.size() == kMaxTOIContacts) break;
                Contact& c = contacts[i];
                if (c.islandFlag) continue;
                f32 backupW = wallAlpha0[c.wall];
                wallAlpha0[c.wall] = minAlpha;          // other->Advance(minAlpha) on a static body
                updateContact(c);
                if (!c.enabled) { wallAlpha0[c.wall] = backupW; continue; }
                if (!c.touching) { wallAlpha0[c.wall] = backupW; continue; }
                c.islandFlag = true; islandContacts.push_back((int)i);
            }
            f32 subDt = (1.0f - minAlpha) * stepDt;
            // b2Island::SolveTOI
            pc_c = sweep.c; pc_a = sweep.a; pv = v; pw = w;
            solverInit(false, 1.0f);
            for (int i = 0; i < 20; ++i) { if (solvePositionConstraints(kToiBaumgarte, -1.5f)) break; }
            sweep.c0 = pc_c; sweep.a0 = pc_a;
            initializeVelocityConstraints();
            for (int i = 0; i < velIters; ++i) solveVelocityConstraints();
            integratePositions(subDt);
            sweep.c = pc_c; sweep.a = pc_a; v = pv; w = pw;
            synchronizeTransform();
            report();
            synchronizeFixtures();
            for (auto& c : contacts) { c.toiFlag = false; c.islandFlag = false; }
            findNewContacts();
        }
    }

    // b2World::Step
    void step(f32 dt, int velIters, int posIters) {
        f32 inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
        f32 dtRatio = inv_dt0 * dt;
        if (newFixture) { findNewContacts(); newFixture = false; }
        collide();
        if (stepComplete && dt > 0.0f) solve(dt, velIters, posIters, dtRatio);
        if (dt > 0.0f) solveTOI(dt, velIters);
        if (dt > 0.0f) inv_dt0 = inv_dt;
        force = mk(0.0f, 0.0f); torque = 0.0f;
    }

    // b2World::RayCast restricted to wall fixtures; returns the final (nearest) reported fraction or 1.
    f32 rayCastWalls(V2 p1, V2 p2) const {
        f32 maxFraction = 1.0f;
        for (const Wall& wl : walls) {
            f32 fr;
            if (wl.poly.rayCast(p1, p2, maxFraction, wl.xf, &fr)) {
                if (fr == 0.0f) return 0.0f;   // callback returned 0: query terminated
                maxFraction = fr;
            }
        }
        return maxFraction;
    }
};

// ------------------------------------------------------------------ optional shared world (default off)
// The reference gives every car a private b2World (src/car_env.py:389-394), so cars never touch each other.  This is the
// Box2D semantics of putting the C cars of an env into ONE world instead: car-vs-car contacts between dynamic boxes, solved
// with full two-body rows together with each car's wall contacts, islands built over touching car-car contacts.  There is no
// reference behaviour to pin; the arithmetic is b2ContactSolver's general form and the orderings that Box2D leaves to its
// contact lists are fixed here as follows (the CUDA engine does the same):
//   * a car-car contact exists while the two fat AABBs overlap, is created after Solve / after the TOI phase like
//     b2ContactManager::FindNewContacts does, and is destroyed in Collide;
//   * an island = the cars connected by touching car-car contacts; its constraint order is every member's wall contacts
//     (members in ascending car index, each car's own list order), then the car-car contacts in ascending (i, j);
//   * TOI: Box2D skips contacts between two non-bullet dynamic bodies, so the continuous phase stays per car against walls.
// Listener: a car-car contact reports to both cars like a wall contact does (key = -1 - other car; the normal points away
// from the reporting car).
struct PairContact { int i, j; bool touching; Manifold m; };
struct PairVC { VelocityConstraint vc; PositionConstraint pc; int i, j; };

struct SharedWorld {
    std::vector<World*> cars;
    std::vector<PairContact> pairs;             // ascending (i, j)
    f32 friction, restitution;                  // b2MixFriction / b2MixRestitution of two car fixtures

    void clear() { pairs.clear(); }
    int find(int i, int j) const { for (size_t k = 0; k < pairs.size(); ++k) if (pairs[k].i == i && pairs[k].j == j) return (int)k; return -1; }
    // b2ContactManager::FindNewContacts for the car proxies against each other
    void findNewPairs() {
        int n = (int)cars.size();
        for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) {
            if (!overlap(cars[i]->carFat, cars[j]->carFat)) continue;
            if (find(i, j) >= 0) continue;
            PairContact pc; pc.i = i; pc.j = j; pc.touching = false; pc.m.pointCount = 0; pc.m.type = FACE_A;
            size_t pos = 0; while (pos < pairs.size() && (pairs[pos].i < i || (pairs[pos].i == i && pairs[pos].j < j))) ++pos;
            pairs.insert(pairs.begin() + pos, pc);
        }
    }
    // b2Contact::Update for a car-car contact
    void updatePair(PairContact& p) {
        World& A = *cars[p.i]; World& B = *cars[p.j];
        Manifold old = p.m;
        bool was = p.touching;
        collidePolygons(&p.m, A.carPoly, A.xf, B.carPoly, B.xf);
        bool touching = p.m.pointCount > 0;
        for (int a = 0; a < p.m.pointCount; ++a) {
            MPoint* mp2 = p.m.points + a; mp2->normalImpulse = 0.0f; mp2->tangentImpulse = 0.0f;
            for (int b = 0; b < old.pointCount; ++b) if (old.points[b].key == mp2->key) {
                mp2->normalImpulse = old.points[b].normalImpulse; mp2->tangentImpulse = old.points[b].tangentImpulse; break;
            }
        }
        if (touching != was) { A.setAwake(true); B.setAwake(true); }
        p.touching = touching;
        if (!was && touching) {
            WorldManifold wm; wm.initialize(&p.m, A.xf, A.carPoly.radius, B.xf, B.carPoly.radius);
            if (A.listener) A.listener->beginContact(-1 - p.j, wm.normal);
            if (B.listener) B.listener->beginContact(-1 - p.i, -wm.normal);
        }
        if (was && !touching) { if (A.listener) A.listener->endContact(-1 - p.j); if (B.listener) B.listener->endContact(-1 - p.i); }
    }
    void collidePairs() {
        for (size_t k = 0; k < pairs.size();) {
            PairContact& p = pairs[k];
            World& A = *cars[p.i]; World& B = *cars[p.j];
            if (!A.awake && !B.awake) { ++k; continue; }
            if (!overlap(A.carFat, B.carFat)) {
                if (p.touching) { if (A.listener) A.listener->endContact(-1 - p.j); if (B.listener) B.listener->endContact(-1 - p.i); }
                pairs.erase(pairs.begin() + k);
                continue;
            }
            updatePair(p);
            ++k;
        }
    }
    // ---- two-body rows (b2ContactSolver, both bodies dynamic)
    std::vector<PairVC> rows;
    void rowsInit(const std::vector<int>& isl, f32 dtRatio) {
        rows.clear();
        for (int k : isl) {
            PairContact& p = pairs[k]; PairVC r; r.i = p.i; r.j = p.j;
            VelocityConstraint& vc = r.vc; PositionConstraint& pc = r.pc;
            vc.friction = friction; vc.restitution = restitution; vc.contactIndex = k; vc.pointCount = p.m.pointCount;
            for (int q = 0; q < 4; ++q) { vc.K[q] = 0.0f; vc.nm[q] = 0.0f; }
            pc.localNormal = p.m.localNormal; pc.localPoint = p.m.localPoint; pc.pointCount = p.m.pointCount; pc.type = p.m.type; pc.wall = -1;
            for (int a = 0; a < p.m.pointCount; ++a) {
                VCPoint& cp = vc.points[a];
                cp.normalImpulse = dtRatio * p.m.points[a].normalImpulse; cp.tangentImpulse = dtRatio * p.m.points[a].tangentImpulse;
                cp.rA = mk(0, 0); cp.rB = mk(0, 0); cp.normalMass = 0.0f; cp.tangentMass = 0.0f; cp.velocityBias = 0.0f;
                pc.localPoints[a] = p.m.points[a].localPoint;
            }
            rows.push_back(r);
        }
    }
    void rowsInitVelocity() {
        for (PairVC& r : rows) {
            World& A = *cars[r.i]; World& B = *cars[r.j];
            VelocityConstraint& vc = r.vc; PairContact& p = pairs[vc.contactIndex];
            f32 mA = A.invMass, mB = B.invMass, iA = A.invI, iB = B.invI;
            V2 cA = A.pc_c, cB = B.pc_c; f32 aA = A.pc_a, aB = B.pc_a; V2 vA = A.pv, vB = B.pv; f32 wA = A.pw, wB = B.pw;
            Xf xfA, xfB; xfA.q.set(aA); xfB.q.set(aB);
            xfA.p = cA - mul(xfA.q, A.sweep.localCenter); xfB.p = cB - mul(xfB.q, B.sweep.localCenter);
            WorldManifold wm; wm.initialize(&p.m, xfA, A.carPoly.radius, xfB, B.carPoly.radius);
            vc.normal = wm.normal;
            for (int a = 0; a < vc.pointCount; ++a) {
                VCPoint& cp = vc.points[a];
                cp.rA = wm.points[a] - cA; cp.rB = wm.points[a] - cB;
                f32 rnA = cross(cp.rA, vc.normal), rnB = cross(cp.rB, vc.normal);
                f32 kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                cp.normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
                V2 tangent = cross(vc.normal, 1.0f);
                f32 rtA = cross(cp.rA, tangent), rtB = cross(cp.rB, tangent);
                f32 kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
                cp.tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
                cp.velocityBias = 0.0f;
                f32 vRel = dot(vc.normal, vB + cross(wB, cp.rB) - vA - cross(wA, cp.rA));
                if (vRel < -kVelocityThreshold) cp.velocityBias = -vc.restitution * vRel;
            }
            if (vc.pointCount == 2) {
                VCPoint& p1 = vc.points[0]; VCPoint& p2 = vc.points[1];
                f32 rn1A = cross(p1.rA, vc.normal), rn1B = cross(p1.rB, vc.normal);
                f32 rn2A = cross(p2.rA, vc.normal), rn2B = cross(p2.rB, vc.normal);
                f32 k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
                f32 k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
                f32 k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
                if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
                    vc.K[0] = k11; vc.K[1] = k12; vc.K[2] = k12; vc.K[3] = k22;
                    f32 det = k11 * k22 - k12 * k12;
                    if (det != 0.0f) det = 1.0f / det;
                    vc.nm[0] = det * k22; vc.nm[2] = -det * k12; vc.nm[1] = -det * k12; vc.nm[3] = det * k11;
                } else vc.pointCount = 1;
            }
        }
    }
    static void applyPair(World& A, World& B, V2 P, V2 rA, V2 rB) {
        A.pv = A.pv - A.invMass * P; A.pw -= A.invI * cross(rA, P);
        B.pv = B.pv + B.invMass * P; B.pw += B.invI * cross(rB, P);
    }
    void rowsWarmStart() {
        for (PairVC& r : rows) {
            World& A = *cars[r.i]; World& B = *cars[r.j]; VelocityConstraint& vc = r.vc;
            V2 tangent = cross(vc.normal, 1.0f);
            for (int a = 0; a < vc.pointCount; ++a) {
                VCPoint& cp = vc.points[a];
                V2 P = cp.normalImpulse * vc.normal + cp.tangentImpulse * tangent;
                applyPair(A, B, P, cp.rA, cp.rB);
            }
        }
    }
    void rowsSolveVelocity() {
        for (PairVC& r : rows) {
            World& A = *cars[r.i]; World& B = *cars[r.j]; VelocityConstraint& vc = r.vc;
            V2 normal = vc.normal, tangent = cross(normal, 1.0f);
            for (int a = 0; a < vc.pointCount; ++a) {
                VCPoint& cp = vc.points[a];
                V2 dv = B.pv + cross(B.pw, cp.rB) - A.pv - cross(A.pw, cp.rA);
                f32 vt = dot(dv, tangent) - 0.0f;
                f32 lambda = cp.tangentMass * (-vt);
                f32 maxFriction = vc.friction * cp.normalImpulse;
                f32 newImpulse = clampf(cp.tangentImpulse + lambda, -maxFriction, maxFriction);
                lambda = newImpulse - cp.tangentImpulse; cp.tangentImpulse = newImpulse;
                applyPair(A, B, lambda * tangent, cp.rA, cp.rB);
            }
            if (vc.pointCount == 1) {
                VCPoint& cp = vc.points[0];
                V2 dv = B.pv + cross(B.pw, cp.rB) - A.pv - cross(A.pw, cp.rA);
                f32 vn = dot(dv, normal);
                f32 lambda = -cp.normalMass * (vn - cp.velocityBias);
                f32 newImpulse = fmax2(cp.normalImpulse + lambda, 0.0f);
                lambda = newImpulse - cp.normalImpulse; cp.normalImpulse = newImpulse;
                applyPair(A, B, lambda * normal, cp.rA, cp.rB);
            } else if (vc.pointCount == 2) {
                VCPoint& cp1 = vc.points[0]; VCPoint& cp2 = vc.points[1];
                V2 a = mk(cp1.normalImpulse, cp2.normalImpulse);
                V2 dv1 = B.pv + cross(B.pw, cp1.rB) - A.pv - cross(A.pw, cp1.rA);
                V2 dv2 = B.pv + cross(B.pw, cp2.rB) - A.pv - cross(A.pw, cp2.rA);
                f32 vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
                V2 b = mk(vn1 - cp1.velocityBias, vn2 - cp2.velocityBias);
                b = b - mk(vc.K[0] * a.x + vc.K[2] * a.y, vc.K[1] * a.x + vc.K[3] * a.y);
                V2 x; bool ok = false;
                for (;;) {
                    x = -mk(vc.nm[0] * b.x + vc.nm[2] * b.y, vc.nm[1] * b.x + vc.nm[3] * b.y);
                    if (x.x >= 0.0f && x.y >= 0.0f) { ok = true; break; }
                    x.x = -cp1.normalMass * b.x; x.y = 0.0f; vn2 = vc.K[1] * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) { ok = true; break; }
                    x.x = 0.0f; x.y = -cp2.normalMass * b.y; vn1 = vc.K[2] * x.y + b.x;
                    if (x.y >= 0.0f && vn1 >= 0.0f) { ok = true; break; }
                    x.x = 0.0f; x.y = 0.0f; vn1 = b.x; vn2 = b.y;
                    if (vn1 >= 0.0f && vn2 >= 0.0f) { ok = true; break; }
                    break;
                }
                if (ok) {
                    V2 d = x - a; V2 P1 = d.x * normal, P2 = d.y * normal;
                    A.pv = A.pv - A.invMass * (P1 + P2); A.pw -= A.invI * (cross(cp1.rA, P1) + cross(cp2.rA, P2));
                    B.pv = B.pv + B.invMass * (P1 + P2); B.pw += B.invI * (cross(cp1.rB, P1) + cross(cp2.rB, P2));
                    cp1.normalImpulse = x.x; cp2.normalImpulse = x.y;
                }
            }
        }
    }
    void rowsStore() {
        for (PairVC& r : rows) { Manifold& m = pairs[r.vc.contactIndex].m; for (int a = 0; a < r.vc.pointCount; ++a) { m.points[a].normalImpulse = r.vc.points[a].normalImpulse; m.points[a].tangentImpulse = r.vc.points[a].tangentImpulse; } }
    }
    bool rowsSolvePosition() {
        f32 minSeparation = 0.0f;
        for (PairVC& r : rows) {
            World& A = *cars[r.i]; World& B = *cars[r.j]; PositionConstraint& pc = r.pc;
            f32 mA = A.invMass, iA = A.invI, mB = B.invMass, iB = B.invI;
            V2 cA = A.pc_c, cB = B.pc_c; f32 aA = A.pc_a, aB = B.pc_a;
            for (int a = 0; a < pc.pointCount; ++a) {
                Xf xfA, xfB; xfA.q.set(aA); xfB.q.set(aB);
                xfA.p = cA - mul(xfA.q, A.sweep.localCenter); xfB.p = cB - mul(xfB.q, B.sweep.localCenter);
                V2 normal, point; f32 separation;
                if (pc.type == FACE_A) {
                    normal = mul(xfA.q, pc.localNormal);
                    V2 planePoint = mul(xfA, pc.localPoint), clipPoint = mul(xfB, pc.localPoints[a]);
                    separation = dot(clipPoint - planePoint, normal) - A.carPoly.radius - B.carPoly.radius;
                    point = clipPoint;
                } else {
                    normal = mul(xfB.q, pc.localNormal);
                    V2 planePoint = mul(xfB, pc.localPoint), clipPoint = mul(xfA, pc.localPoints[a]);
                    separation = dot(clipPoint - planePoint, normal) - A.carPoly.radius - B.carPoly.radius;
                    point = clipPoint; normal = -normal;
                }
                V2 rA = point - cA, rB = point - cB;
                minSeparation = fmin2(minSeparation, separation);
                f32 C = clampf(kBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                f32 rnA = cross(rA, normal), rnB = cross(rB, normal);
                f32 K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                f32 impulse = K > 0.0f ? -C / K : 0.0f;
                V2 P = impulse * normal;
                cA = cA - mA * P; aA -= iA * cross(rA, P);
                cB = cB + mB * P; aB += iB * cross(rB, P);
            }
            A.pc_c = cA; A.pc_a = aA; B.pc_c = cB; B.pc_a = aB;
        }
        return minSeparation >= -3.0f * kLinearSlop;
    }
    void rowsReport() {
        for (PairVC& r : rows) {
            f32 ni[2] = {0.0f, 0.0f};
            for (int a = 0; a < r.vc.pointCount; ++a) ni[a] = r.vc.points[a].normalImpulse;
            if (cars[r.i]->listener) cars[r.i]->listener->postSolve(-1 - r.j, r.vc.pointCount, ni);
            if (cars[r.j]->listener) cars[r.j]->listener->postSolve(-1 - r.i, r.vc.pointCount, ni);
        }
    }
    // b2Island::Solve for the cars `mem` (ascending) joined by the touching pair contacts `isl` (ascending)
    void solveIsland(const std::vector<int>& mem, const std::vector<int>& isl, f32 h, int velIters, int posIters, f32 dtRatio) {
        for (int k : mem) {
            World& W = *cars[k];
            W.setAwake(true);
            W.islandContacts.clear();
            for (size_t c = 0; c < W.contacts.size(); ++c) if (W.contacts[c].enabled && W.contacts[c].touching) W.islandContacts.push_back((int)c);
            W.pc_c = W.sweep.c; W.pc_a = W.sweep.a; W.pv = W.v; W.pw = W.w;
            W.sweep.c0 = W.sweep.c; W.sweep.a0 = W.sweep.a;
            W.pv = W.pv + h * (1.0f * mk(0.0f, 0.0f) + W.invMass * W.force);
            W.pw += h * W.invI * W.torque;
            W.pv = (1.0f / (1.0f + h * 0.0f)) * W.pv;
            W.pw *= 1.0f / (1.0f + h * 0.0f);
            W.solverInit(true, dtRatio);
        }
        rowsInit(isl, dtRatio);
        for (int k : mem) cars[k]->initializeVelocityConstraints();
        rowsInitVelocity();
        for (int k : mem) cars[k]->warmStart();
        rowsWarmStart();
        for (int it = 0; it < velIters; ++it) { for (int k : mem) cars[k]->solveVelocityConstraints(); rowsSolveVelocity(); }
        for (int k : mem) cars[k]->storeImpulses();
        rowsStore();
        for (int k : mem) cars[k]->integratePositions(h);
        bool positionSolved = false;
        for (int it = 0; it < posIters; ++it) {
            bool ok = true;
            for (int k : mem) ok = cars[k]->solvePositionConstraints(kBaumgarte, -3.0f) && ok;
            ok = rowsSolvePosition() && ok;
            if (ok) { positionSolved = true; break; }
        }
        f32 minSleepTime = kMaxFloat;
        for (int k : mem) {
            World& W = *cars[k];
            W.sweep.c = W.pc_c; W.sweep.a = W.pc_a; W.v = W.pv; W.w = W.pw;
            W.synchronizeTransform();
        }
        for (int k : mem) cars[k]->report();
        rowsReport();
        const f32 linTolSqr = kLinearSleepTol * kLinearSleepTol, angTolSqr = kAngularSleepTol * kAngularSleepTol;
        for (int k : mem) {
            World& W = *cars[k];
            if (W.w * W.w > angTolSqr || dot(W.v, W.v) > linTolSqr) { W.sleepTime = 0.0f; minSleepTime = 0.0f; }
            else { W.sleepTime += h; minSleepTime = fmin2(minSleepTime, W.sleepTime); }
        }
        if (minSleepTime >= kTimeToSleep && positionSolved) for (int k : mem) cars[k]->setAwake(false);
        for (int k : mem) { cars[k]->synchronizeFixtures(); cars[k]->findNewContacts(); }
    }
    // b2World::Step of the shared world
    void step(f32 dt, int velIters, int posIters) {
        int n = (int)cars.size();
        f32 inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
        bool anyNew = false;
        for (World* W : cars) if (W->newFixture) { W->findNewContacts(); W->newFixture = false; anyNew = true; }
        if (anyNew) findNewPairs();
        for (World* W : cars) W->collide();
        collidePairs();
        // islands: union of cars over touching car-car contacts
        std::vector<int> root(n); for (int i = 0; i < n; ++i) root[i] = i;
        auto findr = [&](int x) { while (root[x] != x) x = root[x]; return x; };
        for (PairContact& p : pairs) if (p.touching) { int a = findr(p.i), b = findr(p.j); if (a != b) root[a > b ? a : b] = (a > b ? b : a); }
        std::vector<bool> done(n, false);
        for (int i = 0; i < n; ++i) {
            if (done[i]) continue;
            std::vector<int> mem, isl;
            for (int k = i; k < n; ++k) if (!done[k] && findr(k) == findr(i)) mem.push_back(k);
            for (int k : mem) done[k] = true;
            for (size_t q = 0; q < pairs.size(); ++q) if (pairs[q].touching && findr(pairs[q].i) == findr(i)) isl.push_back((int)q);
            if (isl.empty()) {
                World& W = *cars[i];
                if (W.stepComplete && dt > 0.0f) W.solve(dt, velIters, posIters, W.inv_dt0 * dt);
                continue;
            }
            bool anyAwake = false; for (int k : mem) anyAwake = anyAwake || cars[k]->awake;
            if (!anyAwake) { for (int k : mem) cars[k]->findNewContacts(); continue; }
            solveIsland(mem, isl, dt, velIters, posIters, cars[mem[0]]->inv_dt0 * dt);
        }
        findNewPairs();
        for (World* W : cars) { if (dt > 0.0f) W->solveTOI(dt, velIters); }
        findNewPairs();
        for (World* W : cars) { if (dt > 0.0f) W->inv_dt0 = inv_dt; W->force = mk(0.0f, 0.0f); W->torque = 0.0f; }
    }
};

}  // namespace b2
