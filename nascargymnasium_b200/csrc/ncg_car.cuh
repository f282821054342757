// ncg_car.cuh -- one car-step of NascarGymnasium's CarEnv for one car record, as run by one GPU thread
// (scalar phases) plus the lane-cooperative ray phase.  float32 state and arithmetic.
//
// Reference path restated here (file:line in /root/reference/src):
//   car.py:311-892 (engine, brake, drag, rolling, acceleration window, lateral force, damping, banking,
//   steering, friction bookkeeping), tyre_manager.py:99-222, tyre.py:68-228, car_physics.py:341-441 and
//   631-864 (step, banking lookup, collision listener), box2d-py b2World.Step (ncg_b2.cuh + World below),
//   distance_sensor.py:71-117, lap_timer.py:95-273, car_env.py:575-676, 805-1158, 1544-1611.
//
// The same source is compiled for the device by nvcc (the product) and, for CPU-side debugging of the
// scalar phases only, by tests/hostcheck (test infrastructure; never shipped, never a fallback).
#pragma once
#include "ncg_b2.cuh"

namespace ncg {

// ------------------------------------------------------------------ track table view
// Blob layout is produced by nascargymnasium_b200/track.py::build_track_table.
struct Track {
    const float* hdr; const float* segs; const double* seg64; const float* walls; const float* aabb;
    const uint32_t* cells; const uint16_t* items;     // per cell: first block | n_blocks << 16; items in blocks of 4 u16
    const uint16_t* segmask;                          // per cell: segments whose chord can be nearest to a point of the cell
    int n_walls, n_segs, gnx, gny, has_bank;
    float gx0, gy0, inv_cell, cell, ltot, min_lap, slx0, sly0, sldx, sldy, sllen2, slhalfw, half_ltot;
};
enum { TH_NWALLS = 0, TH_NSEGS, TH_GNX, TH_GNY, TH_HASBANK, TH_WORDS, TH_OFF_SEGS, TH_OFF_WALLS, TH_OFF_AABB, TH_OFF_CELLS,
       TH_OFF_ITEMS, TH_NITEMS, TH_GX0, TH_GY0, TH_INVCELL, TH_CELL, TH_LTOT, TH_MINLAP, TH_SLX0, TH_SLY0, TH_SLDX,
       TH_SLDY, TH_SLLEN2, TH_SLHALFW, TH_HALF_LTOT, TH_STAGE_WORDS, TH_OFF_SEG64, TH_OFF_SEGMASK };
enum { SEG_STRIDE = 12, WALL_STRIDE = 8, SEG64_STRIDE = 5 };
// `staged` points at the table the view should read (the copy in shared memory, or the blob in global memory); the whole
// table is staged, so `global` (the blob it was copied from) is only kept for callers that want to re-point a member.
NCG_HD Track track_view(const float* staged, const float* global) {
    (void)global;
    Track t; t.hdr = staged;
    t.n_walls = (int)f2u(staged[TH_NWALLS]); t.n_segs = (int)f2u(staged[TH_NSEGS]);
    t.gnx = (int)f2u(staged[TH_GNX]); t.gny = (int)f2u(staged[TH_GNY]); t.has_bank = (int)f2u(staged[TH_HASBANK]);
    t.segs = staged + f2u(staged[TH_OFF_SEGS]); t.walls = staged + f2u(staged[TH_OFF_WALLS]);
    t.seg64 = (const double*)(staged + f2u(staged[TH_OFF_SEG64]));
    t.cells = (const uint32_t*)(staged + f2u(staged[TH_OFF_CELLS])); t.items = (const uint16_t*)(staged + f2u(staged[TH_OFF_ITEMS]));
    t.segmask = (const uint16_t*)(staged + f2u(staged[TH_OFF_SEGMASK]));
    t.aabb = staged + f2u(staged[TH_OFF_AABB]);
    t.gx0 = staged[TH_GX0]; t.gy0 = staged[TH_GY0]; t.inv_cell = staged[TH_INVCELL]; t.cell = staged[TH_CELL];
    t.ltot = staged[TH_LTOT]; t.min_lap = staged[TH_MINLAP]; t.slx0 = staged[TH_SLX0]; t.sly0 = staged[TH_SLY0];
    t.sldx = staged[TH_SLDX]; t.sldy = staged[TH_SLDY]; t.sllen2 = staged[TH_SLLEN2]; t.slhalfw = staged[TH_SLHALFW];
    t.half_ltot = staged[TH_HALF_LTOT];
    return t;
}
// the k-th wall of a grid cell's list (lists are padded to blocks of 4 by repeating a block's first wall, so a wall
// can appear twice: every consumer is insensitive to duplicates)
NCG_HD int cell_count_max(const Track& T, int cell) { return (int)(T.cells[cell] >> 16) * 4; }
NCG_HD int cell_item(const Track& T, int cell, int k) {
    NCG_CHECK(cell >= 0 && cell < T.gnx * T.gny && k >= 0 && k < cell_count_max(T, cell), "grid cell / list index");
    return (int)T.items[(T.cells[cell] & 0xFFFFu) * 4u + (uint32_t)k];
}
NCG_HD void wall_get(const Track& T, int i, Xf* xf, Box* b) {
    NCG_CHECK(i >= 0 && i < T.n_walls, "wall index");
    const float* w = T.walls + i * WALL_STRIDE;
    xf->p = mk(w[0], w[1]); xf->q.c = w[2]; xf->q.s = w[3]; b->hx = w[7]; b->hy = w[5];      // w[4] is the rays' half-length
}
NCG_HD float wall_angle(const Track& T, int i) { return T.walls[i * WALL_STRIDE + 6]; }
NCG_HD AABB wall_fat(const Track& T, int i) { NCG_CHECK(i >= 0 && i < T.n_walls, "wall AABB index"); const float* a = T.aabb + i * 4; AABB r; r.lx = a[0]; r.ly = a[1]; r.ux = a[2]; r.uy = a[3]; return r; }

// ------------------------------------------------------------------ the per-car Box2D world
struct Contact { int wall; bool touching, enabled, toiFlag, island; int toiCount; float toi; Manifold m;
                 float sep0; };      // lower bound of the car-wall distance at the pose Collide saw this step, or < 0: unknown
struct VCPoint { V2 rA; float ni, ti, nm, tm, bias; };
struct VC { VCPoint p[2]; V2 normal; float K[4], nmat[4]; int pc, ci; };
struct PC { V2 lp[2], localNormal, localPoint; int type, pc, wall; };

// The car's b2Body plus the listener scalars: small enough to live in registers on the contact-free fast path.
struct Body {
    Xf xf; Sweep sweep; V2 v; float w; V2 force; float torque; float sleepTime, inv_dt0;
    bool awake, proxyMoved, newFixture, overflow;
    AABB fat;
    float impulse; bool hasKey;          // CarCollisionListener.car_collision_impulses[car]
};
// Everything else Box2D keeps for a car that has broad-phase contacts; only the (rare) contact path builds it.
struct World {
    Body b;
    int nc; Contact c[NCG_MAX_CONTACTS];
    float wallAlpha[NCG_MAX_CONTACTS];          // alpha0 of the static sweeps of contacted walls (SolveTOI)
    int na; int awall[NCG_MAX_ACTIVE]; float anx[NCG_MAX_ACTIVE], any[NCG_MAX_ACTIVE];   // active_collisions
    int ni; int isl[NCG_MAX_TOUCHING]; VC vc[NCG_MAX_TOUCHING]; PC pcs[NCG_MAX_TOUCHING];
    V2 pc_c; float pc_a; V2 pv; float pw;
    unsigned toi_events;
    bool v230;                                  // b2CollidePolygons as in Box2D 2.3.0 (see collide_boxes)
};
NCG_HD Box car_box() { Box b; b.hx = NCG_CAR_HALF_LENGTH; b.hy = NCG_CAR_HALF_WIDTH; return b; }
#define NCG_INV_MASS (1.0f / NCG_CAR_MASS)
#define NCG_INV_I (1.0f / NCG_CAR_MOI)

NCG_HD void b_set_awake(Body& B, bool flag) {
    if (flag) { if (!B.awake) { B.awake = true; B.sleepTime = 0.0f; } }
    else { B.awake = false; B.sleepTime = 0.0f; B.v = mk(0.0f, 0.0f); B.w = 0.0f; B.force = mk(0.0f, 0.0f); B.torque = 0.0f; }
}
NCG_HD void b_apply_force(Body& B, V2 f, V2 point) { if (!B.awake) b_set_awake(B, true); B.force = B.force + f; B.torque += cross(point - B.sweep.c, f); }
NCG_HD void b_apply_force_center(Body& B, V2 f) { if (!B.awake) b_set_awake(B, true); B.force = B.force + f; }
NCG_HD void b_apply_torque(Body& B, float t) { if (!B.awake) b_set_awake(B, true); B.torque += t; }
NCG_HD void b_sync_transform(Body& B) { B.xf.q = rot(B.sweep.a); B.xf.p = B.sweep.c - mul(B.xf.q, mk(0.0f, 0.0f)); }
NCG_HD void b_move_proxy(Body& B, const AABB& aabb, V2 disp) {
    if (aabb_contains(B.fat, aabb)) return;
    AABB b = aabb;
    b.lx -= NCG_B2_AABB_EXT; b.ly -= NCG_B2_AABB_EXT; b.ux += NCG_B2_AABB_EXT; b.uy += NCG_B2_AABB_EXT;
    V2 d = NCG_B2_AABB_MULT * disp;
    if (d.x < 0.0f) b.lx += d.x; else b.ux += d.x;
    if (d.y < 0.0f) b.ly += d.y; else b.uy += d.y;
    B.fat = b; B.proxyMoved = true;
}
NCG_HD void b_set_transform(Body& B, V2 p, float angle) {
    B.xf.q = rot(angle); B.xf.p = p;
    B.sweep.c = mul(B.xf, mk(0.0f, 0.0f)); B.sweep.a = angle; B.sweep.c0 = B.sweep.c; B.sweep.a0 = angle;
    b_move_proxy(B, box_aabb(car_box(), B.xf), mk(0.0f, 0.0f));
}
// b2Body::SynchronizeFixtures with the start-of-step transform xf1 given (it equals rot(a0), c0 when no TOI advance
// has touched the sweep, which saves a sincosf on the contact-free path)
NCG_HD void b_sync_fixtures_from(Body& B, const Xf& xf1) {
    AABB a1 = box_aabb(car_box(), xf1), a2 = box_aabb(car_box(), B.xf);
    AABB c; c.lx = fminb(a1.lx, a2.lx); c.ly = fminb(a1.ly, a2.ly); c.ux = fmaxb(a1.ux, a2.ux); c.uy = fmaxb(a1.uy, a2.uy);
    b_move_proxy(B, c, B.xf.p - xf1.p);
}
NCG_HD void b_sync_fixtures(Body& B) {
    Xf xf1; xf1.q = rot(B.sweep.a0); xf1.p = B.sweep.c0 - mul(xf1.q, mk(0.0f, 0.0f));
    AABB a1 = box_aabb(car_box(), xf1), a2 = box_aabb(car_box(), B.xf);
    AABB c; c.lx = fminb(a1.lx, a2.lx); c.ly = fminb(a1.ly, a2.ly); c.ux = fmaxb(a1.ux, a2.ux); c.uy = fmaxb(a1.uy, a2.uy);
    b_move_proxy(B, c, B.xf.p - xf1.p);
}
NCG_HD void b_advance(Body& B, float alpha) {
    sweep_advance(B.sweep, alpha); B.sweep.c = B.sweep.c0; B.sweep.a = B.sweep.a0;
    B.xf.q = rot(B.sweep.a); B.xf.p = B.sweep.c - mul(B.xf.q, mk(0.0f, 0.0f));
}
// integrate velocities (forces) -- b2Island::Solve prologue for a body with no joints, gravity or damping
NCG_HD void b_integrate_velocity(Body& B, float h) {
    B.v = B.v + h * (1.0f * mk(0.0f, 0.0f) + NCG_INV_MASS * B.force);
    B.w += h * NCG_INV_I * B.torque;
}
NCG_HDN float max_translation_ratio(V2 tr) { return NCG_B2_MAX_TRANSLATION / length(tr); }      // out of line: see collision_angle
// integrate positions with b2_maxTranslation / b2_maxRotation clamps, on (c, a, v, w)
NCG_HD void integrate_position(V2& c, float& a, V2& v, float& w, float h) {
    V2 tr = h * v;
    if (__builtin_expect(dot(tr, tr) > NCG_B2_MAX_TRANSLATION * NCG_B2_MAX_TRANSLATION, 0)) v = max_translation_ratio(tr) * v;     // > 2 m per step: never at racing speeds
    float ro = h * w;
    if (__builtin_expect(ro * ro > NCG_B2_MAX_ROTATION * NCG_B2_MAX_ROTATION, 0)) { float ratio = NCG_B2_MAX_ROTATION / fabsf(ro); w *= ratio; }
    c = c + h * v; a += h * w;
}
// sleep bookkeeping at the end of b2Island::Solve
NCG_HD void b_sleep(Body& B, float h, bool positionSolved) {
    float minSleep = NCG_B2_MAXFLOAT;
    if (B.w * B.w > NCG_B2_ANG_SLEEP_TOL * NCG_B2_ANG_SLEEP_TOL || dot(B.v, B.v) > NCG_B2_LIN_SLEEP_TOL * NCG_B2_LIN_SLEEP_TOL) { B.sleepTime = 0.0f; minSleep = 0.0f; }
    else { B.sleepTime += h; minSleep = fminb(minSleep, B.sleepTime); }
    if (minSleep >= NCG_B2_TIME_TO_SLEEP && positionSolved) b_set_awake(B, false);
}
#define w_set_awake(W, f) b_set_awake((W).b, f)
#define w_sync_transform(W) b_sync_transform((W).b)
#define w_sync_fixtures(W) b_sync_fixtures((W).b)
#define w_advance(W, a) b_advance((W).b, a)
// listener callbacks
NCG_HD void l_begin(World& W, int wall, V2 n) {
    bool found = false;
    for (int i = 0; i < W.na; ++i) if (W.awall[i] == wall) { W.anx[i] = n.x; W.any[i] = n.y; found = true; break; }
    if (!found) { if (W.na < NCG_MAX_ACTIVE) { W.awall[W.na] = wall; W.anx[W.na] = n.x; W.any[W.na] = n.y; ++W.na; } else W.b.overflow = true; }
    if (!W.b.hasKey) { W.b.hasKey = true; W.b.impulse = 0.0f; }
}
NCG_HD void l_end(World& W, int wall) {
    for (int i = 0; i < W.na; ++i) if (W.awall[i] == wall) {
        for (int j = i; j + 1 < W.na; ++j) { W.awall[j] = W.awall[j + 1]; W.anx[j] = W.anx[j + 1]; W.any[j] = W.any[j + 1]; }
        --W.na; break;
    }
    if (W.na == 0) { W.b.impulse = 0.0f; W.b.hasKey = true; }
}
NCG_HD void l_post_solve(World& W, int count, float n0, float n1) {
    if (count > 0) {
        float total = n0; if (count > 1) total = total + n1;
        if (W.b.hasKey) W.b.impulse = total > W.b.impulse ? total : W.b.impulse;
    }
}
// b2Contact::Update
NCG_HDN void w_update_contact(World& W, const Track& T, Contact& c) {
    Manifold old = c.m;
    c.enabled = true;
    bool was = c.touching;
    Xf xfB; Box bB; wall_get(T, c.wall, &xfB, &bB);
    collide_boxes(&c.m, car_box(), W.b.xf, bB, xfB, W.v230, &c.sep0);
    bool touching = c.m.pc > 0;
    for (int i = 0; i < c.m.pc; ++i) {
        c.m.ni[i] = 0.0f; c.m.ti[i] = 0.0f;
        for (int j = 0; j < old.pc; ++j) if (old.key[j] == c.m.key[i]) { c.m.ni[i] = old.ni[j]; c.m.ti[i] = old.ti[j]; break; }
    }
    if (touching != was) w_set_awake(W, true);
    c.touching = touching;
    if (!was && touching) { V2 n, pts[2]; world_manifold(&n, pts, c.m, W.b.xf, xfB); l_begin(W, c.wall, n); }
    if (was && !touching) l_end(W, c.wall);
}
// b2ContactManager::Collide
NCG_HDN void w_collide(World& W, const Track& T) {
    int i = 0;
    while (i < W.nc) {
        if (!W.b.awake) { ++i; continue; }
        Contact& c = W.c[i];
        if (!aabb_overlap(W.b.fat, wall_fat(T, c.wall))) {
            if (c.touching) l_end(W, c.wall);
            for (int j = i; j + 1 < W.nc; ++j) W.c[j] = W.c[j + 1];
            --W.nc; continue;
        }
        w_update_contact(W, T, c);
        ++i;
    }
}
// b2ContactManager::FindNewContacts: pairs = walls whose fat AABB overlaps the car's, ascending wall index,
// each head-inserted.  Candidates come from the uniform grid (cells overlapped by the car's fat AABB).
NCG_HDN void w_find_new_contacts(World& W, const Track& T) {
    if (!W.b.proxyMoved) return;
    W.b.proxyMoved = false;
    int ix0 = (int)floorf((W.b.fat.lx - T.gx0) * T.inv_cell), ix1 = (int)floorf((W.b.fat.ux - T.gx0) * T.inv_cell);
    int iy0 = (int)floorf((W.b.fat.ly - T.gy0) * T.inv_cell), iy1 = (int)floorf((W.b.fat.uy - T.gy0) * T.inv_cell);
    ix0 = ix0 < 0 ? 0 : ix0; iy0 = iy0 < 0 ? 0 : iy0; ix1 = ix1 >= T.gnx ? T.gnx - 1 : ix1; iy1 = iy1 >= T.gny ? T.gny - 1 : iy1;
    int found[NCG_MAX_CONTACTS]; int nf = 0;
    for (int iy = iy0; iy <= iy1; ++iy) for (int ix = ix0; ix <= ix1; ++ix) {
        int cell = iy * T.gnx + ix;
        const int nk = cell_count_max(T, cell);
        for (int k = 0; k < nk; ++k) {
            int wi = cell_item(T, cell, k);
            if (!aabb_overlap(W.b.fat, wall_fat(T, wi))) continue;
            bool have = false;
            for (int j = 0; j < W.nc; ++j) if (W.c[j].wall == wi) { have = true; break; }
            for (int j = 0; j < nf && !have; ++j) if (found[j] == wi) have = true;
            if (have) continue;
            if (nf < NCG_MAX_CONTACTS) found[nf++] = wi; else W.b.overflow = true;
        }
    }
    // ascending wall index, then head-insert one by one
    for (int i = 1; i < nf; ++i) { int key = found[i], j = i - 1; while (j >= 0 && found[j] > key) { found[j + 1] = found[j]; --j; } found[j + 1] = key; }
    for (int i = 0; i < nf; ++i) {
        if (W.nc >= NCG_MAX_CONTACTS) { W.b.overflow = true; break; }
        for (int j = W.nc; j > 0; --j) { W.c[j] = W.c[j - 1]; W.wallAlpha[j] = W.wallAlpha[j - 1]; }
        Contact& c = W.c[0];
        c.wall = found[i]; c.touching = false; c.enabled = true; c.toiFlag = false; c.island = false; c.toiCount = 0; c.toi = 1.0f;
        c.m.pc = 0; c.m.type = FACE_A; c.sep0 = -1.0f;
        W.wallAlpha[0] = 0.0f;
        ++W.nc;
    }
}
// ---- b2ContactSolver over the island {car} + isl[]
// (`#pragma unroll 1` on the loops over contacts / manifold points of the contact path: their trip counts are 1-2 at run time, the
// unrolled copies only made the code a contact step has to fetch larger -- 16 KB per kernel; the driving distribution gained 3 %)
NCG_HDN void s_init(World& W, bool warm, float dtRatio) {
#pragma unroll 1
    for (int i = 0; i < W.ni; ++i) {
        Contact& c = W.c[W.isl[i]]; VC& vc = W.vc[i]; PC& pc = W.pcs[i];
        vc.ci = W.isl[i]; vc.pc = c.m.pc;
        for (int k = 0; k < 4; ++k) { vc.K[k] = 0.0f; vc.nmat[k] = 0.0f; }
        pc.localNormal = c.m.localNormal; pc.localPoint = c.m.localPoint; pc.pc = c.m.pc; pc.type = c.m.type; pc.wall = c.wall;
        for (int j = 0; j < c.m.pc; ++j) {
            VCPoint& p = vc.p[j];
            if (warm) { p.ni = dtRatio * c.m.ni[j]; p.ti = dtRatio * c.m.ti[j]; } else { p.ni = 0.0f; p.ti = 0.0f; }
            p.rA = mk(0.0f, 0.0f); p.nm = 0.0f; p.tm = 0.0f; p.bias = 0.0f;
            pc.lp[j] = c.m.lp[j];
        }
    }
}
NCG_HDN void s_init_velocity(World& W, const Track& T) {
    const float mA = NCG_INV_MASS, iA = NCG_INV_I, restitution = NCG_WALL_RESTITUTION;   // max(0.1, 0.25)
#pragma unroll 1
    for (int i = 0; i < W.ni; ++i) {
        VC& vc = W.vc[i]; PC& pc = W.pcs[i]; Contact& c = W.c[vc.ci];
        Xf xfB; Box bB; wall_get(T, pc.wall, &xfB, &bB);
        V2 cA = W.pc_c; float aA = W.pc_a; V2 vA = W.pv; float wA = W.pw;
        Xf xfA; xfA.q = rot(aA); xfA.p = cA - mul(xfA.q, mk(0.0f, 0.0f));
        V2 pts[2]; world_manifold(&vc.normal, pts, c.m, xfA, xfB);
#pragma unroll 1
        for (int j = 0; j < vc.pc; ++j) {
            VCPoint& p = vc.p[j];
            p.rA = pts[j] - cA;
            float rnA = cross(p.rA, vc.normal);
            float kNormal = mA + iA * rnA * rnA;
            p.nm = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
            V2 tangent = cross(vc.normal, 1.0f);
            float rtA = cross(p.rA, tangent);
            float kTangent = mA + iA * rtA * rtA;
            p.tm = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
            p.bias = 0.0f;
            float vRel = dot(vc.normal, (-vA) - cross(wA, p.rA));
            if (vRel < -NCG_B2_VEL_THRESHOLD) p.bias = -restitution * vRel;
        }
        if (vc.pc == 2) {
            float rn1A = cross(vc.p[0].rA, vc.normal), rn2A = cross(vc.p[1].rA, vc.normal);
            float k11 = mA + iA * rn1A * rn1A, k22 = mA + iA * rn2A * rn2A, k12 = mA + iA * rn1A * rn2A;
            if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
                vc.K[0] = k11; vc.K[1] = k12; vc.K[2] = k12; vc.K[3] = k22;
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = 1.0f / det;
                vc.nmat[0] = det * k22; vc.nmat[2] = -det * k12; vc.nmat[1] = -det * k12; vc.nmat[3] = det * k11;
            } else vc.pc = 1;
        }
    }
}
NCG_HDN void s_warm_start(World& W) {
    const float mA = NCG_INV_MASS, iA = NCG_INV_I;
#pragma unroll 1
    for (int i = 0; i < W.ni; ++i) {
        VC& vc = W.vc[i];
        V2 tangent = cross(vc.normal, 1.0f);
#pragma unroll 1
        for (int j = 0; j < vc.pc; ++j) {
            V2 P = vc.p[j].ni * vc.normal + vc.p[j].ti * tangent;
            W.pw -= iA * cross(vc.p[j].rA, P);
            W.pv = W.pv - mA * P;
        }
    }
}
NCG_HD void s_apply2(V2& vA, float& wA, const VC& vc, V2 x, V2 a) {
    const float mA = NCG_INV_MASS, iA = NCG_INV_I;
    V2 d = x - a; V2 P1 = d.x * vc.normal, P2 = d.y * vc.normal;
    vA = vA - mA * (P1 + P2); wA -= iA * (cross(vc.p[0].rA, P1) + cross(vc.p[1].rA, P2));
}
NCG_HDN void s_solve_velocity(World& W) {
    const float mA = NCG_INV_MASS, iA = NCG_INV_I;
    const float friction = sqrtf(NCG_CAR_FRICTION * NCG_WALL_FRICTION);
#pragma unroll 1
    for (int i = 0; i < W.ni; ++i) {
        VC& vc = W.vc[i];
        V2 vA = W.pv; float wA = W.pw;
        V2 normal = vc.normal, tangent = cross(normal, 1.0f);
#pragma unroll 1
        for (int j = 0; j < vc.pc; ++j) {
            VCPoint& p = vc.p[j];
            V2 dv = (-vA) - cross(wA, p.rA);
            float vt = dot(dv, tangent) - 0.0f;
            float lambda = p.tm * (-vt);
            float maxF = friction * p.ni;
            float ni = clampb(p.ti + lambda, -maxF, maxF);
            lambda = ni - p.ti; p.ti = ni;
            V2 P = lambda * tangent;
            vA = vA - mA * P; wA -= iA * cross(p.rA, P);
        }
        if (vc.pc == 1) {
            VCPoint& p = vc.p[0];
            V2 dv = (-vA) - cross(wA, p.rA);
            float vn = dot(dv, normal);
            float lambda = -p.nm * (vn - p.bias);
            float ni = fmaxb(p.ni + lambda, 0.0f);
            lambda = ni - p.ni; p.ni = ni;
            V2 P = lambda * normal;
            vA = vA - mA * P; wA -= iA * cross(p.rA, P);
        } else if (vc.pc == 2) {
            VCPoint& c1 = vc.p[0]; VCPoint& c2 = vc.p[1];
            V2 a = mk(c1.ni, c2.ni);
            V2 dv1 = (-vA) - cross(wA, c1.rA), dv2 = (-vA) - cross(wA, c2.rA);
            float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
            V2 b = mk(vn1 - c1.bias, vn2 - c2.bias);
            b = b - mk(vc.K[0] * a.x + vc.K[2] * a.y, vc.K[1] * a.x + vc.K[3] * a.y);
            for (;;) {
                V2 x = -mk(vc.nmat[0] * b.x + vc.nmat[2] * b.y, vc.nmat[1] * b.x + vc.nmat[3] * b.y);
                if (x.x >= 0.0f && x.y >= 0.0f) { s_apply2(vA, wA, vc, x, a); c1.ni = x.x; c2.ni = x.y; break; }
                x.x = -c1.nm * b.x; x.y = 0.0f; vn2 = vc.K[1] * x.x + b.y;
                if (x.x >= 0.0f && vn2 >= 0.0f) { s_apply2(vA, wA, vc, x, a); c1.ni = x.x; c2.ni = x.y; break; }
                x.x = 0.0f; x.y = -c2.nm * b.y; vn1 = vc.K[2] * x.y + b.x;
                if (x.y >= 0.0f && vn1 >= 0.0f) { s_apply2(vA, wA, vc, x, a); c1.ni = x.x; c2.ni = x.y; break; }
                x.x = 0.0f; x.y = 0.0f; vn1 = b.x; vn2 = b.y;
                if (vn1 >= 0.0f && vn2 >= 0.0f) { s_apply2(vA, wA, vc, x, a); c1.ni = x.x; c2.ni = x.y; break; }
                break;
            }
        }
        W.pv = vA; W.pw = wA;
    }
}
NCG_HD void s_store_impulses(World& W) {
    for (int i = 0; i < W.ni; ++i) { VC& vc = W.vc[i]; Manifold& m = W.c[vc.ci].m; for (int j = 0; j < vc.pc; ++j) { m.ni[j] = vc.p[j].ni; m.ti[j] = vc.p[j].ti; } }
}
NCG_HDN bool s_solve_position(World& W, const Track& T, float baumgarte, float okFactor) {
    const float mA = NCG_INV_MASS, iA = NCG_INV_I;
    float minSep = 0.0f;
#pragma unroll 1
    for (int i = 0; i < W.ni; ++i) {
        PC& pc = W.pcs[i];
        Xf xfB; Box bB; wall_get(T, pc.wall, &xfB, &bB);
        V2 cA = W.pc_c; float aA = W.pc_a;
        for (int j = 0; j < pc.pc; ++j) {
            Xf xfA; xfA.q = rot(aA); xfA.p = cA - mul(xfA.q, mk(0.0f, 0.0f));
            V2 normal, point; float separation;
            if (pc.type == FACE_A) {
                normal = mul(xfA.q, pc.localNormal);
                V2 planePoint = mul(xfA, pc.localPoint), clip = mul(xfB, pc.lp[j]);
                separation = dot(clip - planePoint, normal) - NCG_B2_POLY_RADIUS - NCG_B2_POLY_RADIUS;
                point = clip;
            } else {
                normal = mul(xfB.q, pc.localNormal);
                V2 planePoint = mul(xfB, pc.localPoint), clip = mul(xfA, pc.lp[j]);
                separation = dot(clip - planePoint, normal) - NCG_B2_POLY_RADIUS - NCG_B2_POLY_RADIUS;
                point = clip; normal = -normal;
            }
            V2 rA = point - cA;
            minSep = fminb(minSep, separation);
            float C = clampb(baumgarte * (separation + NCG_B2_LINEAR_SLOP), -NCG_B2_MAX_LIN_CORR, 0.0f);
            float rnA = cross(rA, normal);
            float K = mA + iA * rnA * rnA;
            float impulse = K > 0.0f ? -C / K : 0.0f;
            V2 P = impulse * normal;
            cA = cA - mA * P; aA -= iA * cross(rA, P);
        }
        W.pc_c = cA; W.pc_a = aA;
    }
    return minSep >= okFactor * NCG_B2_LINEAR_SLOP;
}
NCG_HD void s_report(World& W) {
    for (int i = 0; i < W.ni; ++i) { VC& vc = W.vc[i]; l_post_solve(W, vc.pc, vc.p[0].ni, vc.pc > 1 ? vc.p[1].ni : 0.0f); }
}
NCG_HD void s_integrate(World& W, float h) { integrate_position(W.pc_c, W.pc_a, W.pv, W.pw, h); }
// b2World::Solve + b2Island::Solve for a car that has contacts in its list
NCG_HDN void w_solve(World& W, const Track& T, float h, float dtRatio) {
    Body& B = W.b;
    if (B.awake) {
        W.ni = 0;
        for (int i = 0; i < W.nc; ++i) if (W.c[i].enabled && W.c[i].touching) { if (W.ni < NCG_MAX_TOUCHING) W.isl[W.ni++] = i; else B.overflow = true; }
        B.sweep.c0 = B.sweep.c; B.sweep.a0 = B.sweep.a;
        b_integrate_velocity(B, h);
        W.pc_c = B.sweep.c; W.pc_a = B.sweep.a; W.pv = B.v; W.pw = B.w;
        bool positionSolved;
        if (W.ni > 0) {
            s_init(W, true, dtRatio); s_init_velocity(W, T); s_warm_start(W);
            for (int i = 0; i < NCG_VEL_ITERS; ++i) s_solve_velocity(W);
            s_store_impulses(W);
            s_integrate(W, h);
            positionSolved = false;
            for (int i = 0; i < NCG_POS_ITERS; ++i) if (s_solve_position(W, T, NCG_B2_BAUMGARTE, -3.0f)) { positionSolved = true; break; }
        } else { s_integrate(W, h); positionSolved = true; }
        B.sweep.c = W.pc_c; B.sweep.a = W.pc_a; B.v = W.pv; B.w = W.pw;
        b_sync_transform(B);
        if (W.ni > 0) s_report(W);
        b_sleep(B, h, positionSolved);
        b_sync_fixtures(B);
    }
}
// b2World::SolveTOI + b2Island::SolveTOI
// A lower bound of the distance between the car box and a (static) wall box over the WHOLE sweep c0,a0 -> c,a, on the wall's
// own face normals as separating axes.  Along the unit axis u the separation is d(t) - r(t) - h: d = u.(car centre - wall
// centre) is linear in t, so >= min(d0, d1); r = the car box's support radius along u, a maximum of sinusoids in the heading
// of amplitude R (half diagonal), so over the heading interval r <= max(r0, r1) + R da^2 / 8.  The distance of two convex
// shapes is at least their separation on any axis.
NCG_HD float sweep_face_bound(const Sweep& s, const Rot q1, const Xf& xfB, const Box& bB, float enough) {
    const Rot q0 = rot(s.a0);
    const float da = s.a - s.a0, curv = 2.7114f * da * da * 0.125f;
    const V2 r0 = s.c0 - xfB.p, r1 = s.c - xfB.p;
    float best = -NCG_B2_MAXFLOAT;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const V2 u = k == 0 ? mk(xfB.q.c, xfB.q.s) : mk(-xfB.q.s, xfB.q.c);
        const float h = k == 0 ? bB.hx : bB.hy;
        const float d0 = dot(u, r0), d1 = dot(u, r1);
        const V2 m0 = mulT(q0, u), m1 = mulT(q1, u);                   // the axis in the car's frame at both ends
        const float sup0 = fabsf(m0.x) * NCG_CAR_HALF_LENGTH + fabsf(m0.y) * NCG_CAR_HALF_WIDTH;
        const float sup1 = fabsf(m1.x) * NCG_CAR_HALF_LENGTH + fabsf(m1.y) * NCG_CAR_HALF_WIDTH;
        const float r = fmaxb(sup0, sup1) + curv + h;
        best = fmaxb(best, fmaxb(fminb(d0, d1) - r, -fmaxb(d0, d1) - r));
    }
    if (best > enough) return best;                                    // (most scraping contacts: decided on the wall's axes)
#ifndef NCG_NO_TOI_CAR_AXES
    // the car's own face normals (a wall corner against the car's side: the inner wall of a bend).  In the car's frame the
    // wall centre is D(t) = R(a(t))^T (p - c(t)), per component |D''| <= da^2 |p - c|max + 2 |da| |dc|, and the wall box's
    // support radius along a car axis is again a maximum of sinusoids in the heading, of amplitude <= hx + hy.
    const V2 D0 = mulT(q0, -1.0f * r0), D1 = mulT(q1, -1.0f * r1);
    const float ada = fabsf(da);
    const float wmax = fmaxb(length(r0), length(r1)), dd = (da * da * wmax + 2.0f * ada * length(s.c - s.c0)) * 0.125f;
    const V2 ex0 = mulT(q0, mk(xfB.q.c, xfB.q.s)), ex1 = mulT(q1, mk(xfB.q.c, xfB.q.s));      // the wall's x axis in the car's frame; its y axis is (-ex.y, ex.x)
    const float scurv = (bB.hx + bB.hy) * da * da * 0.125f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float a0 = k == 0 ? D0.x : D0.y, a1 = k == 0 ? D1.x : D1.y;
        const float s0 = k == 0 ? fabsf(ex0.x) * bB.hx + fabsf(ex0.y) * bB.hy : fabsf(ex0.y) * bB.hx + fabsf(ex0.x) * bB.hy;
        const float s1 = k == 0 ? fabsf(ex1.x) * bB.hx + fabsf(ex1.y) * bB.hy : fabsf(ex1.y) * bB.hx + fabsf(ex1.x) * bB.hy;
        const float r = fmaxb(s0, s1) + scurv + dd + (k == 0 ? NCG_CAR_HALF_LENGTH : NCG_CAR_HALF_WIDTH);
        best = fmaxb(best, fmaxb(fminb(a0, a1) - r, -fmaxb(a0, a1) - r));
    }
#endif
    return best;
}
#if !defined(__CUDA_ARCH__)
static unsigned long long g_toi_full = 0, g_toi_skip_reach = 0, g_toi_skip_face = 0;      // host compile: how the queries were answered
#endif
NCG_HDN void w_solve_toi(World& W, const Track& T, float stepDt) {
    W.b.sweep.alpha0 = 0.0f;
    for (int i = 0; i < W.nc; ++i) { W.wallAlpha[i] = 0.0f; Contact& c = W.c[i]; c.toiFlag = false; c.island = false; c.toiCount = 0; c.toi = 1.0f; }
    // Most broad-phase contacts of a step are walls the car merely passes: b2TimeOfImpact can only answer e_touching (the one
    // answer that yields alpha < 1) if some separation it evaluates along the sweep comes down to `target`; every separation it
    // evaluates is at least the distance at the start pose minus what a car point can travel in the step.  Collide measured a
    // lower bound of that distance on a separating axis this very step (Contact::sep0, at the pose the sweep starts from), so
    // while nothing has advanced the sweep yet a contact with  sep0 > target + tolerance + 2 * (|dc| + |da| * R) + 1 cm  keeps
    // alpha = 1 without the GJK / root-finder query -- the same outcome, contact by contact (the parity tests compare the TOI
    // counters and contact lists exactly).  R = the car's half diagonal.
    const float toi_reach = 2.0f * (length(W.b.sweep.c - W.b.sweep.c0) + fabsf(W.b.sweep.a - W.b.sweep.a0) * 2.7114f)
                            + (NCG_B2_LINEAR_SLOP + 0.25f * NCG_B2_LINEAR_SLOP) + 0.01f;
    bool pristine = true;                                   // no TOI event has advanced the sweep or touched the contacts yet
    for (;;) {
        int minC = -1; float minAlpha = 1.0f;
#pragma unroll 1
        for (int i = 0; i < W.nc; ++i) {
            Contact& c = W.c[i];
            if (!c.enabled) continue;
            if (c.toiCount > NCG_B2_MAX_SUBSTEPS) continue;
            float alpha = 1.0f;
            if (c.toiFlag) alpha = c.toi;
            else {
                if (!W.b.awake) continue;
#ifndef NCG_NO_TOI_SHORTCUT      /* (tests/test_hostcheck.py builds the host compile both ways and demands identical records) */
                if (pristine && c.sep0 > toi_reach) {
#if !defined(__CUDA_ARCH__)
                    ++g_toi_skip_reach;
#endif
                    c.toi = 1.0f; c.toiFlag = true; continue;
                }
                // A car that scrapes along a wall keeps ~1.5 cm between the core shapes (the position solver's resting
                // penetration) and does not come closer during the step: the same argument with a sharper bound -- the
                // separation on the wall's face normals over the whole sweep -- answers those without the query too.  The
                // margin covers the rounding of both computations (coordinates of ~1e3 m carry 1e-4 m).
                if (pristine) {
                    Xf xfW; Box bW; wall_get(T, c.wall, &xfW, &bW);
                    const float enough = NCG_B2_LINEAR_SLOP + 0.25f * NCG_B2_LINEAR_SLOP + 1e-3f;
                    if (sweep_face_bound(W.b.sweep, W.b.xf.q, xfW, bW, enough) > enough) {
#if !defined(__CUDA_ARCH__)
                        ++g_toi_skip_face;
#endif
                        c.toi = 1.0f; c.toiFlag = true; continue;
                    }
                }
#endif
#if !defined(__CUDA_ARCH__)
                ++g_toi_full;
#endif
                float alpha0 = W.b.sweep.alpha0;
                if (W.b.sweep.alpha0 < W.wallAlpha[i]) { alpha0 = W.wallAlpha[i]; sweep_advance(W.b.sweep, alpha0); }
                else if (W.wallAlpha[i] < W.b.sweep.alpha0) { alpha0 = W.b.sweep.alpha0; W.wallAlpha[i] = alpha0; }
                Xf xfB; Box bB; wall_get(T, c.wall, &xfB, &bB);
                Sweep sB; sB.c0 = xfB.p; sB.c = xfB.p; sB.a0 = wall_angle(T, c.wall); sB.a = sB.a0; sB.alpha0 = W.wallAlpha[i];
                int state; float t;
                time_of_impact(&state, &t, car_box(), W.b.sweep, bB, sB, 1.0f);
                if (state == TOI_TOUCHING) alpha = fminb(alpha0 + (1.0f - alpha0) * t, 1.0f); else alpha = 1.0f;
                c.toi = alpha; c.toiFlag = true;
            }
            if (alpha < minAlpha) { minC = i; minAlpha = alpha; }
        }
        if (minC < 0 || 1.0f - 10.0f * NCG_B2_EPS < minAlpha) break;
        Contact& mc = W.c[minC];
        pristine = false;
        Sweep backup = W.b.sweep; float backupWall = W.wallAlpha[minC];
        w_advance(W, minAlpha); W.wallAlpha[minC] = minAlpha;
        w_update_contact(W, T, mc);
        mc.toiFlag = false; ++mc.toiCount;
        if (!mc.enabled || !mc.touching) { mc.enabled = false; W.b.sweep = backup; W.wallAlpha[minC] = backupWall; w_sync_transform(W); continue; }
        w_set_awake(W, true);
        W.ni = 0; W.isl[W.ni++] = minC; mc.island = true;
#pragma unroll 1
        for (int i = 0; i < W.nc; ++i) {
            Contact& c = W.c[i];
            if (c.island) continue;
            float bw = W.wallAlpha[i];
            W.wallAlpha[i] = minAlpha;
            w_update_contact(W, T, c);
            if (!c.enabled || !c.touching) { W.wallAlpha[i] = bw; continue; }
            if (W.ni < NCG_MAX_TOUCHING) { c.island = true; W.isl[W.ni++] = i; } else W.b.overflow = true;
        }
        float subDt = (1.0f - minAlpha) * stepDt;
        W.pc_c = W.b.sweep.c; W.pc_a = W.b.sweep.a; W.pv = W.b.v; W.pw = W.b.w;
        s_init(W, false, 1.0f);
        for (int i = 0; i < 20; ++i) if (s_solve_position(W, T, NCG_B2_TOI_BAUMGARTE, -1.5f)) break;
        W.b.sweep.c0 = W.pc_c; W.b.sweep.a0 = W.pc_a;
        s_init_velocity(W, T);
        for (int i = 0; i < NCG_VEL_ITERS; ++i) s_solve_velocity(W);
        s_integrate(W, subDt);
        W.b.sweep.c = W.pc_c; W.b.sweep.a = W.pc_a; W.b.v = W.pv; W.b.w = W.pw;
        w_sync_transform(W);
        s_report(W);
        w_sync_fixtures(W);
        for (int i = 0; i < W.nc; ++i) { W.c[i].toiFlag = false; W.c[i].island = false; }
        w_find_new_contacts(W, T);
        ++W.toi_events;
    }
}
// ------------------------------------------------------------------ record <-> Body / World
NCG_HD void b_load(Body& B, const float* R) {
    B.sweep.c = mk(R[NCG_R_X], R[NCG_R_Y]); B.sweep.a = R[NCG_R_ANGLE]; B.sweep.c0 = B.sweep.c; B.sweep.a0 = B.sweep.a; B.sweep.alpha0 = 0.0f;
    b_sync_transform(B);
    B.v = mk(R[NCG_R_VX], R[NCG_R_VY]); B.w = R[NCG_R_OMEGA]; B.force = mk(0.0f, 0.0f); B.torque = 0.0f;
    B.sleepTime = R[NCG_R_SLEEP]; B.inv_dt0 = R[NCG_R_INV_DT0];
    uint32_t fl = f2u(R[NCG_R_FLAGS]);
    B.awake = (fl & NCG_F_AWAKE) != 0; B.proxyMoved = (fl & NCG_F_PROXY_MOVED) != 0; B.newFixture = (fl & NCG_F_NEW_FIXTURE) != 0;
    B.hasKey = (fl & NCG_F_HAS_KEY) != 0; B.overflow = false;
    B.fat.lx = R[NCG_R_FAT_LX]; B.fat.ly = R[NCG_R_FAT_LY]; B.fat.ux = R[NCG_R_FAT_UX]; B.fat.uy = R[NCG_R_FAT_UY];
    B.impulse = R[NCG_R_IMPULSE];
}
NCG_HD void b_store(const Body& B, float* R, uint32_t* flags) {
    R[NCG_R_X] = B.sweep.c.x; R[NCG_R_Y] = B.sweep.c.y; R[NCG_R_ANGLE] = B.sweep.a; R[NCG_R_VX] = B.v.x; R[NCG_R_VY] = B.v.y; R[NCG_R_OMEGA] = B.w;
    R[NCG_R_SLEEP] = B.sleepTime; R[NCG_R_INV_DT0] = B.inv_dt0;
    uint32_t fl = *flags & ~(uint32_t)(NCG_F_AWAKE | NCG_F_PROXY_MOVED | NCG_F_NEW_FIXTURE | NCG_F_HAS_KEY);
    if (B.awake) fl |= NCG_F_AWAKE;
    if (B.proxyMoved) fl |= NCG_F_PROXY_MOVED;
    if (B.newFixture) fl |= NCG_F_NEW_FIXTURE;
    if (B.hasKey) fl |= NCG_F_HAS_KEY;
    if (B.overflow) fl |= NCG_F_OVERFLOW;
    *flags = fl;
    R[NCG_R_FAT_LX] = B.fat.lx; R[NCG_R_FAT_LY] = B.fat.ly; R[NCG_R_FAT_UX] = B.fat.ux; R[NCG_R_FAT_UY] = B.fat.uy;
    R[NCG_R_IMPULSE] = B.impulse;
}
NCG_HD void w_load_contacts(World& W, const float* R) {
    W.toi_events = 0;
    uint32_t ncw = f2u(R[NCG_R_NCONTACT]);
    W.nc = (int)(ncw & 255u); W.na = (int)((ncw >> 8) & 255u);
    uint32_t tmask = (ncw >> 16) & 0xFFFu, pcw = f2u(R[NCG_R_MANIFOLD_PC]);
    int k = 0;
#pragma unroll 1
    for (int i = 0; i < W.nc; ++i) {
        Contact& c = W.c[i];
        uint32_t ww = f2u(R[NCG_R_CONTACT_WALL + (i >> 1)]);
        c.wall = (int)((i & 1) ? (ww >> 16) : (ww & 0xFFFFu));
        c.touching = ((tmask >> i) & 1u) != 0; c.enabled = true; c.toiFlag = false; c.island = false; c.toiCount = 0; c.toi = 1.0f; c.sep0 = -1.0f;
        c.m.pc = 0; c.m.type = FACE_A; c.m.localNormal = mk(0.0f, 0.0f); c.m.localPoint = mk(0.0f, 0.0f);
        W.wallAlpha[i] = 0.0f;
        if (c.touching && k < NCG_MAX_TOUCHING) {
            const float* M = R + NCG_R_MANIFOLD + 6 * k;
            c.m.pc = (int)((pcw >> (2 * k)) & 3u);
            c.m.key[0] = f2u(M[0]); c.m.key[1] = f2u(M[1]); c.m.ni[0] = M[2]; c.m.ti[0] = M[3]; c.m.ni[1] = M[4]; c.m.ti[1] = M[5];
            c.m.lp[0] = mk(0.0f, 0.0f); c.m.lp[1] = mk(0.0f, 0.0f);
            ++k;
        }
    }
    for (int i = 0; i < W.na; ++i) { W.awall[i] = (int)f2u(R[NCG_R_ACTIVE + 3 * i]); W.anx[i] = R[NCG_R_ACTIVE + 3 * i + 1]; W.any[i] = R[NCG_R_ACTIVE + 3 * i + 2]; }
}
NCG_HD void w_store_contacts(World& W, float* R) {
    uint32_t tmask = 0, pcw = 0; int k = 0;
    uint32_t ww[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int i = 0; i < W.nc; ++i) {
        const Contact& c = W.c[i];
        ww[i >> 1] |= ((uint32_t)c.wall & 0xFFFFu) << ((i & 1) * 16);
        if (c.touching && k < NCG_MAX_TOUCHING) {
            tmask |= 1u << i; pcw |= ((uint32_t)c.m.pc & 3u) << (2 * k);
            float* M = R + NCG_R_MANIFOLD + 6 * k;
            M[0] = u2f(c.m.key[0]); M[1] = u2f(c.m.pc > 1 ? c.m.key[1] : 0u); M[2] = c.m.ni[0]; M[3] = c.m.ti[0];
            M[4] = c.m.pc > 1 ? c.m.ni[1] : 0.0f; M[5] = c.m.pc > 1 ? c.m.ti[1] : 0.0f;
            ++k;
        } else if (c.touching) W.b.overflow = true;
    }
    for (int i = 0; i < 6; ++i) R[NCG_R_CONTACT_WALL + i] = u2f(ww[i]);
    R[NCG_R_NCONTACT] = u2f((uint32_t)W.nc | ((uint32_t)W.na << 8) | (tmask << 16));
    R[NCG_R_MANIFOLD_PC] = u2f(pcw);
    for (int i = 0; i < W.na; ++i) { R[NCG_R_ACTIVE + 3 * i] = u2f((uint32_t)W.awall[i]); R[NCG_R_ACTIVE + 3 * i + 1] = W.anx[i]; R[NCG_R_ACTIVE + 3 * i + 2] = W.any[i]; }
}
// does any wall's fat AABB overlap the car's?  (what b2BroadPhase::UpdatePairs would turn into new contacts)
// One block of four walls per iteration (the grid lists are blocks of four, a short one padded by repeating a wall):
// the four AABB rows are independent loads and the four tests are branch-free, so a check is a few load latencies
// instead of one dependent load chain per wall -- some lane of the physics warp runs this on almost every step.
NCG_HD bool aabb_apart(const AABB& a, const F4 b) {      // the two early-outs of b2TestOverlap, b = {lx, ly, ux, uy}
    return (b.x - a.ux > 0.0f) | (b.y - a.uy > 0.0f) | (a.lx - b.z > 0.0f) | (a.ly - b.w > 0.0f);
}
NCG_HD bool any_wall_overlap(const Track& T, const AABB& fat) {
    int ix0 = (int)floorf((fat.lx - T.gx0) * T.inv_cell), ix1 = (int)floorf((fat.ux - T.gx0) * T.inv_cell);
    int iy0 = (int)floorf((fat.ly - T.gy0) * T.inv_cell), iy1 = (int)floorf((fat.uy - T.gy0) * T.inv_cell);
    ix0 = ix0 < 0 ? 0 : ix0; iy0 = iy0 < 0 ? 0 : iy0; ix1 = ix1 >= T.gnx ? T.gnx - 1 : ix1; iy1 = iy1 >= T.gny ? T.gny - 1 : iy1;
    for (int iy = iy0; iy <= iy1; ++iy) for (int ix = ix0; ix <= ix1; ++ix) {
        const uint32_t h = T.cells[iy * T.gnx + ix];
        for (uint32_t k = h & 0xFFFFu, e = k + (h >> 16); k < e; ++k) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(T.items + 4u * k);
            const uint32_t lo = q[0], hi = q[1];
            const F4 a0 = *reinterpret_cast<const F4*>(T.aabb + 4u * (lo & 0xFFFFu)), a1 = *reinterpret_cast<const F4*>(T.aabb + 4u * (lo >> 16));
            const F4 a2 = *reinterpret_cast<const F4*>(T.aabb + 4u * (hi & 0xFFFFu)), a3 = *reinterpret_cast<const F4*>(T.aabb + 4u * (hi >> 16));
            if (!(aabb_apart(fat, a0) & aabb_apart(fat, a1) & aabb_apart(fat, a2) & aabb_apart(fat, a3))) return true;
        }
    }
    return false;
}
// The general b2World::Step for a car with (or about to get) contacts.  `resume`: the fast path has already done
// Collide (nothing to do) and Solve for a car whose contact list was empty, and found that the moved proxy now
// overlaps a wall; continue from FindNewContacts.
// (Track by value and the caller passes copies of its Body / Counters: nothing of the inlined fast path has its
// address taken, so the fast path keeps all of it in registers.)
NCG_HDN void step_with_contacts(Body& B, float* R, const Track T, float dt, bool resume, bool v230, Counters* cnt) {
    World W; W.b = B; W.v230 = v230;
    w_load_contacts(W, R);
    float dtRatio = B.inv_dt0 * dt;
    if (!resume) {
        if (W.b.newFixture) { w_find_new_contacts(W, T); W.b.newFixture = false; }
        w_collide(W, T);
        w_solve(W, T, dt, dtRatio);
    }
    w_find_new_contacts(W, T);
    w_solve_toi(W, T, dt);
    w_store_contacts(W, R);
    B = W.b;
    int nt = 0; for (int i = 0; i < W.nc; ++i) nt += W.c[i].touching ? 1 : 0;
    if (nt) cnt->contact_steps++;
    cnt->toi_events += W.toi_events;
}
// b2World::Step
// contacts: 0 = contact-free integrator, 1 = Box2D 2.3.1+ collision, 2 = Box2D 2.3.0 collision (NcgConfig.contacts)
NCG_HD void body_step(Body& B, float* R, const Track& T, float dt, int contacts, Counters* cnt) {
    const float inv_dt = 1.0f / dt;
    const int nc = (int)(f2u(R[NCG_R_NCONTACT]) & 255u);
    bool slow = contacts && nc > 0;
    if (!slow && contacts && B.newFixture) {           // first Step of a fresh world: pairs of the initial proxy
        if (any_wall_overlap(T, B.fat)) slow = true; else { B.newFixture = false; B.proxyMoved = false; }
    }
    int solver = slow ? 1 : 0;                          // 1: the whole Step with contacts, 2: only FindNewContacts + TOI after the lone-body step
    if (!slow) {
        if (B.awake) {                                 // b2Island::Solve of a lone body
            const Xf xf0 = B.xf;                       // == (rot(a0), c0): b_load synchronised it from the sweep
            B.sweep.c0 = B.sweep.c; B.sweep.a0 = B.sweep.a;
            b_integrate_velocity(B, dt);
            integrate_position(B.sweep.c, B.sweep.a, B.v, B.w, dt);
            b_sync_transform(B);
            b_sleep(B, dt, true);
            b_sync_fixtures_from(B, xf0);
        }
        if (contacts && B.proxyMoved) {                // FindNewContacts
            if (any_wall_overlap(T, B.fat)) solver = 2; else B.proxyMoved = false;
        }
    }
    // one cold block for both ways in (a copy of the body goes in and comes back: see step_with_contacts)
    if (__builtin_expect(solver != 0, 0)) { Body b2 = B; Counters c2 = *cnt; step_with_contacts(b2, R, T, dt, solver == 2, contacts == 2, &c2); B = b2; *cnt = c2; }
    B.inv_dt0 = inv_dt;
    B.force = mk(0.0f, 0.0f); B.torque = 0.0f;
}

// (x, y, a): CarEnv(start_position=..., start_angle=...) (car_env.py:114-115, 391, 398); float32 as Box2D takes them
struct StartPose { float x, y, a; };

// ------------------------------------------------------------------ optional shared world: the cars of an env collide
// (NcgConfig.car_contacts, default off).  The reference gives every car a private b2World, so there is no
// reference behaviour: this is Box2D's semantics for ONE world holding the env's cars -- car-vs-car contacts between dynamic
// boxes, two-body solver rows, islands over touching car-car contacts -- restated in oracle/b2lite.h (SharedWorld), whose
// orderings this follows line by line.  An env that has car-car contacts is stepped by ONE lane (its first car's) over the
// per-car Worlds, which then live in a global-memory scratch; envs without any keep the per-lane path.
// Pair table of an env: NCG_CC_STRIDE words; pair (i < j) at 8 * pair_index: {flags | pc << 8, key0, key1, ni0, ti0, ni1, ti1, -},
// word NCG_CC_COUNT = number of existing pairs.
#define NCG_MAX_PAIRS 45
#define NCG_CC_COUNT (8 * NCG_MAX_PAIRS)
#define NCG_CC_STRIDE (8 * NCG_MAX_PAIRS + 8)
#define NCG_MAX_PAIR_ROWS 12          /* touching car-car contacts solved per env and step (more: counted as overflow) */
enum { PAIR_EXISTS = 1u, PAIR_TOUCHING = 2u };
NCG_HD int pair_index(int i, int j, int C) { return i * (2 * C - i - 1) / 2 + (j - i - 1); }
struct PVCPoint { V2 rA, rB; float ni, ti, nm, tm, bias; };
struct PairRow { PVCPoint p[2]; V2 normal; float K[4], nmat[4]; int pc, i, j, pid; Manifold m; };
// CarEnv start grid of the shared world: car k at (-(k / 2) * dx, +-dy / 2) in the start frame
NCG_HD StartPose cc_start_pose(const StartPose sp, int k, float gdx, float gdy) {
    const double lx = -(double)(k / 2) * (double)gdx, ly = (k % 2 == 0 ? 0.5 : -0.5) * (double)gdy;
    const double c = cos((double)sp.a), s_ = sin((double)sp.a);
    StartPose o; o.a = sp.a;
    o.x = (float)((double)sp.x + c * lx - s_ * ly);
    o.y = (float)((double)sp.y + s_ * lx + c * ly);
    return o;
}
NCG_HDN void cc_find_new_pairs(const AABB* fat, int C, float* PT) {
    uint32_t n = f2u(PT[NCG_CC_COUNT]);
    for (int i = 0; i < C; ++i) for (int j = i + 1; j < C; ++j) {
        float* P = PT + 8 * pair_index(i, j, C);
        if (f2u(P[0]) & PAIR_EXISTS) continue;
        if (!aabb_overlap(fat[i], fat[j])) continue;
        P[0] = u2f(PAIR_EXISTS); P[1] = u2f(0u); P[2] = u2f(0u); P[3] = P[4] = P[5] = P[6] = 0.0f;
        ++n;
    }
    PT[NCG_CC_COUNT] = u2f(n);
}
NCG_HD void cc_apply(World& A, World& B, V2 P, V2 rA, V2 rB) {
    A.pv = A.pv - NCG_INV_MASS * P; A.pw -= NCG_INV_I * cross(rA, P);
    B.pv = B.pv + NCG_INV_MASS * P; B.pw += NCG_INV_I * cross(rB, P);
}
// b2Island::Solve of the cars mem[0..nm) (ascending) joined by rows[ridx[0..nr)] (ascending pair index)
NCG_HDN void cc_solve_island(World* Ws, const Track& T, const int* mem, int nm, PairRow* rows, const int* ridx, int nr, float* PT, float h) {
    const float mA = NCG_INV_MASS, mB = NCG_INV_MASS, iA = NCG_INV_I, iB = NCG_INV_I;
    const float friction = sqrtf(NCG_CAR_FRICTION * NCG_CAR_FRICTION), restitution = NCG_CAR_RESTITUTION;
    const float dtRatio = Ws[mem[0]].b.inv_dt0 * h;
    for (int q = 0; q < nm; ++q) {
        World& W = Ws[mem[q]]; Body& B = W.b;
        b_set_awake(B, true);
        W.ni = 0;
        for (int c = 0; c < W.nc; ++c) if (W.c[c].enabled && W.c[c].touching) { if (W.ni < NCG_MAX_TOUCHING) W.isl[W.ni++] = c; else B.overflow = true; }
        B.sweep.c0 = B.sweep.c; B.sweep.a0 = B.sweep.a;
        b_integrate_velocity(B, h);
        W.pc_c = B.sweep.c; W.pc_a = B.sweep.a; W.pv = B.v; W.pw = B.w;
        s_init(W, true, dtRatio);
    }
    for (int q = 0; q < nr; ++q) {          // rowsInit: warm-start impulses
        PairRow& r = rows[ridx[q]];
        r.pc = r.m.pc;
        for (int k = 0; k < 4; ++k) { r.K[k] = 0.0f; r.nmat[k] = 0.0f; }
        for (int a = 0; a < r.m.pc; ++a) { r.p[a].ni = dtRatio * r.m.ni[a]; r.p[a].ti = dtRatio * r.m.ti[a]; r.p[a].nm = r.p[a].tm = r.p[a].bias = 0.0f; }
    }
    for (int q = 0; q < nm; ++q) s_init_velocity(Ws[mem[q]], T);
    for (int q = 0; q < nr; ++q) {          // rowsInitVelocity
        PairRow& r = rows[ridx[q]]; World& A = Ws[r.i]; World& B = Ws[r.j];
        Xf xfA, xfB; xfA.q = rot(A.pc_a); xfA.p = A.pc_c - mul(xfA.q, mk(0.0f, 0.0f)); xfB.q = rot(B.pc_a); xfB.p = B.pc_c - mul(xfB.q, mk(0.0f, 0.0f));
        V2 pts[2]; world_manifold(&r.normal, pts, r.m, xfA, xfB);
        for (int a = 0; a < r.pc; ++a) {
            PVCPoint& cp = r.p[a];
            cp.rA = pts[a] - A.pc_c; cp.rB = pts[a] - B.pc_c;
            float rnA = cross(cp.rA, r.normal), rnB = cross(cp.rB, r.normal);
            float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
            cp.nm = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
            V2 tangent = cross(r.normal, 1.0f);
            float rtA = cross(cp.rA, tangent), rtB = cross(cp.rB, tangent);
            float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
            cp.tm = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
            cp.bias = 0.0f;
            float vRel = dot(r.normal, B.pv + cross(B.pw, cp.rB) - A.pv - cross(A.pw, cp.rA));
            if (vRel < -NCG_B2_VEL_THRESHOLD) cp.bias = -restitution * vRel;
        }
        if (r.pc == 2) {
            float rn1A = cross(r.p[0].rA, r.normal), rn1B = cross(r.p[0].rB, r.normal), rn2A = cross(r.p[1].rA, r.normal), rn2B = cross(r.p[1].rB, r.normal);
            float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B, k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
            float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
            if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
                r.K[0] = k11; r.K[1] = k12; r.K[2] = k12; r.K[3] = k22;
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = 1.0f / det;
                r.nmat[0] = det * k22; r.nmat[2] = -det * k12; r.nmat[1] = -det * k12; r.nmat[3] = det * k11;
            } else r.pc = 1;
        }
    }
    for (int q = 0; q < nm; ++q) s_warm_start(Ws[mem[q]]);
    for (int q = 0; q < nr; ++q) {
        PairRow& r = rows[ridx[q]]; V2 tangent = cross(r.normal, 1.0f);
        for (int a = 0; a < r.pc; ++a) cc_apply(Ws[r.i], Ws[r.j], r.p[a].ni * r.normal + r.p[a].ti * tangent, r.p[a].rA, r.p[a].rB);
    }
    for (int it = 0; it < NCG_VEL_ITERS; ++it) {
        for (int q = 0; q < nm; ++q) s_solve_velocity(Ws[mem[q]]);
        for (int q = 0; q < nr; ++q) {      // rowsSolveVelocity
            PairRow& r = rows[ridx[q]]; World& A = Ws[r.i]; World& B = Ws[r.j];
            V2 normal = r.normal, tangent = cross(normal, 1.0f);
            for (int a = 0; a < r.pc; ++a) {
                PVCPoint& cp = r.p[a];
                V2 dv = B.pv + cross(B.pw, cp.rB) - A.pv - cross(A.pw, cp.rA);
                float vt = dot(dv, tangent) - 0.0f;
                float lambda = cp.tm * (-vt);
                float maxF = friction * cp.ni;
                float ni = clampb(cp.ti + lambda, -maxF, maxF);
                lambda = ni - cp.ti; cp.ti = ni;
                cc_apply(A, B, lambda * tangent, cp.rA, cp.rB);
            }
            if (r.pc == 1) {
                PVCPoint& cp = r.p[0];
                V2 dv = B.pv + cross(B.pw, cp.rB) - A.pv - cross(A.pw, cp.rA);
                float vn = dot(dv, normal);
                float lambda = -cp.nm * (vn - cp.bias);
                float ni = fmaxb(cp.ni + lambda, 0.0f);
                lambda = ni - cp.ni; cp.ni = ni;
                cc_apply(A, B, lambda * normal, cp.rA, cp.rB);
            } else if (r.pc == 2) {
                PVCPoint& c1 = r.p[0]; PVCPoint& c2 = r.p[1];
                V2 a = mk(c1.ni, c2.ni);
                V2 dv1 = B.pv + cross(B.pw, c1.rB) - A.pv - cross(A.pw, c1.rA), dv2 = B.pv + cross(B.pw, c2.rB) - A.pv - cross(A.pw, c2.rA);
                float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
                V2 b = mk(vn1 - c1.bias, vn2 - c2.bias);
                b = b - mk(r.K[0] * a.x + r.K[2] * a.y, r.K[1] * a.x + r.K[3] * a.y);
                V2 x; bool ok = false;
                for (;;) {
                    x = -mk(r.nmat[0] * b.x + r.nmat[2] * b.y, r.nmat[1] * b.x + r.nmat[3] * b.y);
                    if (x.x >= 0.0f && x.y >= 0.0f) { ok = true; break; }
                    x.x = -c1.nm * b.x; x.y = 0.0f; vn2 = r.K[1] * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) { ok = true; break; }
                    x.x = 0.0f; x.y = -c2.nm * b.y; vn1 = r.K[2] * x.y + b.x;
                    if (x.y >= 0.0f && vn1 >= 0.0f) { ok = true; break; }
                    x.x = 0.0f; x.y = 0.0f; vn1 = b.x; vn2 = b.y;
                    if (vn1 >= 0.0f && vn2 >= 0.0f) { ok = true; break; }
                    break;
                }
                if (ok) {
                    V2 d = x - a; V2 P1 = d.x * normal, P2 = d.y * normal;
                    A.pv = A.pv - mA * (P1 + P2); A.pw -= iA * (cross(c1.rA, P1) + cross(c2.rA, P2));
                    B.pv = B.pv + mB * (P1 + P2); B.pw += iB * (cross(c1.rB, P1) + cross(c2.rB, P2));
                    c1.ni = x.x; c2.ni = x.y;
                }
            }
        }
    }
    for (int q = 0; q < nm; ++q) s_store_impulses(Ws[mem[q]]);
    for (int q = 0; q < nr; ++q) {          // rowsStore: warm-start impulses of the next step
        PairRow& r = rows[ridx[q]]; float* P = PT + 8 * r.pid;
        for (int a = 0; a < r.pc; ++a) { P[3 + 2 * a] = r.p[a].ni; P[4 + 2 * a] = r.p[a].ti; }
    }
    for (int q = 0; q < nm; ++q) s_integrate(Ws[mem[q]], h);
    bool positionSolved = false;
    for (int it = 0; it < NCG_POS_ITERS; ++it) {
        bool ok = true;
        for (int q = 0; q < nm; ++q) ok = s_solve_position(Ws[mem[q]], T, NCG_B2_BAUMGARTE, -3.0f) && ok;
        float minSep = 0.0f;
        for (int q = 0; q < nr; ++q) {      // rowsSolvePosition
            PairRow& r = rows[ridx[q]]; World& A = Ws[r.i]; World& B = Ws[r.j];
            V2 cA = A.pc_c, cB = B.pc_c; float aA = A.pc_a, aB = B.pc_a;
            for (int a = 0; a < r.m.pc; ++a) {
                Xf xfA, xfB; xfA.q = rot(aA); xfA.p = cA - mul(xfA.q, mk(0.0f, 0.0f)); xfB.q = rot(aB); xfB.p = cB - mul(xfB.q, mk(0.0f, 0.0f));
                V2 normal, point; float separation;
                if (r.m.type == FACE_A) {
                    normal = mul(xfA.q, r.m.localNormal);
                    V2 planePoint = mul(xfA, r.m.localPoint), clip = mul(xfB, r.m.lp[a]);
                    separation = dot(clip - planePoint, normal) - NCG_B2_POLY_RADIUS - NCG_B2_POLY_RADIUS;
                    point = clip;
                } else {
                    normal = mul(xfB.q, r.m.localNormal);
                    V2 planePoint = mul(xfB, r.m.localPoint), clip = mul(xfA, r.m.lp[a]);
                    separation = dot(clip - planePoint, normal) - NCG_B2_POLY_RADIUS - NCG_B2_POLY_RADIUS;
                    point = clip; normal = -normal;
                }
                V2 rA = point - cA, rB = point - cB;
                minSep = fminb(minSep, separation);
                float Cc = clampb(NCG_B2_BAUMGARTE * (separation + NCG_B2_LINEAR_SLOP), -NCG_B2_MAX_LIN_CORR, 0.0f);
                float rnA = cross(rA, normal), rnB = cross(rB, normal);
                float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                float impulse = K > 0.0f ? -Cc / K : 0.0f;
                V2 P = impulse * normal;
                cA = cA - mA * P; aA -= iA * cross(rA, P);
                cB = cB + mB * P; aB += iB * cross(rB, P);
            }
            A.pc_c = cA; A.pc_a = aA; B.pc_c = cB; B.pc_a = aB;
        }
        ok = (minSep >= -3.0f * NCG_B2_LINEAR_SLOP) && ok;
        if (ok) { positionSolved = true; break; }
    }
    for (int q = 0; q < nm; ++q) { World& W = Ws[mem[q]]; Body& B = W.b; B.sweep.c = W.pc_c; B.sweep.a = W.pc_a; B.v = W.pv; B.w = W.pw; b_sync_transform(B); }
    for (int q = 0; q < nm; ++q) if (Ws[mem[q]].ni > 0) s_report(Ws[mem[q]]);
    for (int q = 0; q < nr; ++q) {          // rowsReport: both cars' listeners
        PairRow& r = rows[ridx[q]];
        l_post_solve(Ws[r.i], r.pc, r.p[0].ni, r.pc > 1 ? r.p[1].ni : 0.0f);
        l_post_solve(Ws[r.j], r.pc, r.p[0].ni, r.pc > 1 ? r.p[1].ni : 0.0f);
    }
    float minSleep = NCG_B2_MAXFLOAT;
    for (int q = 0; q < nm; ++q) {
        Body& B = Ws[mem[q]].b;
        if (B.w * B.w > NCG_B2_ANG_SLEEP_TOL * NCG_B2_ANG_SLEEP_TOL || dot(B.v, B.v) > NCG_B2_LIN_SLEEP_TOL * NCG_B2_LIN_SLEEP_TOL) { B.sleepTime = 0.0f; minSleep = 0.0f; }
        else { B.sleepTime += h; minSleep = fminb(minSleep, B.sleepTime); }
    }
    if (minSleep >= NCG_B2_TIME_TO_SLEEP && positionSolved) for (int q = 0; q < nm; ++q) b_set_awake(Ws[mem[q]].b, false);
    for (int q = 0; q < nm; ++q) { b_sync_fixtures(Ws[mem[q]].b); w_find_new_contacts(Ws[mem[q]], T); }
}
// b2World::Step of the shared world of one env: Ws[0..C) hold the cars' bodies (forces applied) and wall contacts
NCG_HDN void shared_world_step(World* Ws, int C, float* PT, const Track T, float dt, Counters* cnt) {
    const bool v230 = Ws[0].v230;
    for (int k = 0; k < C; ++k) { Ws[k].toi_events = 0; if (Ws[k].b.newFixture) { w_find_new_contacts(Ws[k], T); Ws[k].b.newFixture = false; } }
    for (int k = 0; k < C; ++k) w_collide(Ws[k], T);
    // ---- Collide of the car-car contacts (ascending pair index = ascending (i, j))
    PairRow rows[NCG_MAX_PAIR_ROWS]; int nrows = 0;
    uint32_t count = f2u(PT[NCG_CC_COUNT]);
    for (int i = 0; i < C; ++i) for (int j = i + 1; j < C; ++j) {
        const int pid = pair_index(i, j, C);
        float* P = PT + 8 * pid;
        const uint32_t fl = f2u(P[0]);
        if (!(fl & PAIR_EXISTS)) continue;
        World& A = Ws[i]; World& B = Ws[j];
        const bool was = (fl & PAIR_TOUCHING) != 0;
        if (!A.b.awake && !B.b.awake) continue;
        if (!aabb_overlap(A.b.fat, B.b.fat)) {
            if (was) { l_end(A, -1 - j); l_end(B, -1 - i); }
            P[0] = u2f(0u); --count;
            continue;
        }
        Manifold m;
        collide_boxes(&m, car_box(), A.b.xf, car_box(), B.b.xf, v230);
        const int opc = (int)((fl >> 8) & 3u);
        const uint32_t okey[2] = {f2u(P[1]), f2u(P[2])}; const float oni[2] = {P[3], P[5]}, oti[2] = {P[4], P[6]};
        for (int a = 0; a < m.pc; ++a) {
            m.ni[a] = 0.0f; m.ti[a] = 0.0f;
            for (int b = 0; b < opc; ++b) if (okey[b] == m.key[a]) { m.ni[a] = oni[b]; m.ti[a] = oti[b]; break; }
        }
        const bool touching = m.pc > 0;
        if (touching != was) { b_set_awake(A.b, true); b_set_awake(B.b, true); }
        if (!was && touching) { V2 n, pts[2]; world_manifold(&n, pts, m, A.b.xf, B.b.xf); l_begin(A, -1 - j, n); l_begin(B, -1 - i, -n); }
        if (was && !touching) { l_end(A, -1 - j); l_end(B, -1 - i); }
        P[0] = u2f(PAIR_EXISTS | (touching ? PAIR_TOUCHING : 0u) | ((uint32_t)m.pc << 8));
        P[1] = u2f(m.pc > 0 ? m.key[0] : 0u); P[2] = u2f(m.pc > 1 ? m.key[1] : 0u);
        P[3] = m.pc > 0 ? m.ni[0] : 0.0f; P[4] = m.pc > 0 ? m.ti[0] : 0.0f; P[5] = m.pc > 1 ? m.ni[1] : 0.0f; P[6] = m.pc > 1 ? m.ti[1] : 0.0f;
        if (touching) {
            if (nrows < NCG_MAX_PAIR_ROWS) { PairRow& r = rows[nrows++]; r.m = m; r.i = i; r.j = j; r.pid = pid; }
            else A.b.overflow = true;
        }
    }
    PT[NCG_CC_COUNT] = u2f(count);
    // ---- Solve: islands = cars connected by touching car-car contacts (root = the lowest car index)
    int root[NCG_MAX_CARS];
    for (int k = 0; k < C; ++k) root[k] = k;
    for (int q = 0; q < nrows; ++q) {
        int a = rows[q].i, b = rows[q].j;
        while (root[a] != a) a = root[a];
        while (root[b] != b) b = root[b];
        if (a != b) { if (a > b) root[a] = b; else root[b] = a; }
    }
    for (int k = 0; k < C; ++k) { int r = k; while (root[r] != r) r = root[r]; root[k] = r; }
    for (int i = 0; i < C; ++i) {
        if (root[i] != i) continue;
        int mem[NCG_MAX_CARS], nm = 0, ridx[NCG_MAX_PAIR_ROWS], nr = 0;
        for (int k = i; k < C; ++k) if (root[k] == i) mem[nm++] = k;
        for (int q = 0; q < nrows; ++q) if (root[rows[q].i] == i) ridx[nr++] = q;
        if (nr == 0) {
            World& W = Ws[i];
            w_solve(W, T, dt, W.b.inv_dt0 * dt);
            w_find_new_contacts(W, T);
            continue;
        }
        bool anyAwake = false;
        for (int q = 0; q < nm; ++q) anyAwake = anyAwake || Ws[mem[q]].b.awake;
        if (!anyAwake) { for (int q = 0; q < nm; ++q) w_find_new_contacts(Ws[mem[q]], T); continue; }
        cc_solve_island(Ws, T, mem, nm, rows, ridx, nr, PT, dt);
    }
    // ---- continuous phase: per car against walls (Box2D skips contacts between two non-bullet dynamic bodies)
    for (int k = 0; k < C; ++k) w_solve_toi(Ws[k], T, dt);
    AABB fat[NCG_MAX_CARS];
    for (int k = 0; k < C; ++k) {
        fat[k] = Ws[k].b.fat;
        int nt = 0; for (int c = 0; c < Ws[k].nc; ++c) nt += Ws[k].c[c].touching ? 1 : 0;
        if (nt) cnt->contact_steps++;
        cnt->toi_events += Ws[k].toi_events;
    }
    cc_find_new_pairs(fat, C, PT);
}

// ------------------------------------------------------------------ tyres (tyre.py, tyre_manager.py)
NCG_HD float tyre_grip(float T, float wear) {
    float tg;
    if (85.0f <= T && T <= 105.0f) tg = 1.5f;
    else { float dev = T < 85.0f ? 85.0f - T : T - 105.0f; tg = fmaxf(0.8f, 1.5f - dev * 0.02f); }
    return tg * (1.0f - (wear * 0.01f) * 0.5f);
}
NCG_HD float total_grip(const float* R) {
    float tg = 0.0f, tw = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { float l = R[NCG_R_TYRE_LOAD + i]; tg += tyre_grip(R[NCG_R_TYRE_TEMP + i], R[NCG_R_TYRE_WEAR + i]) * l; tw += l; }
    return tw > 0.0f ? tg / tw : 0.0f;
}
NCG_HD void weight_transfer(float along, float alat, float speed, float* loads) {
    float down = 0.0f;
    if (speed > 50.0f) { float sf = (speed * 0.02f) * (speed * 0.02f); down = fminf(0.12f * sf * NCG_WEIGHT, 1.5f * NCG_WEIGHT); }
    float base_front = NCG_WEIGHT * 0.5f + down * (1.0f - 0.6f), base_rear = NCG_WEIGHT * 0.5f + down * 0.6f;
    float tew = NCG_WEIGHT + down;
    float raw_long = along * tew * 0.02f;
    float lt = raw_long > 0.0f ? fminf(raw_long, base_front * 0.95f) : fmaxf(raw_long, -(base_rear * 0.95f));
    float front = base_front - lt, rear = base_rear + lt;
    float raw_lat = alat * tew * 0.01f;
    float ml = fminf(front / 2.0f - 200.0f, rear / 2.0f - 200.0f);
    float latt = ml > 0.0f ? fmaxf(-ml, fminf(ml, raw_lat)) : 0.0f;
    float raw[4] = {front / 2.0f - latt / 2.0f, front / 2.0f + latt / 2.0f, rear / 2.0f - latt / 2.0f, rear / 2.0f + latt / 2.0f};
    float deficit = 0.0f, excess = 0.0f, con[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (raw[i] < 50.0f) { con[i] = 50.0f; deficit += 50.0f - raw[i]; } else { con[i] = raw[i]; excess += raw[i] - 50.0f; } }
    if (deficit > 0.0f && excess > 0.0f) {
        float rf = deficit / excess;
#pragma unroll
        for (int i = 0; i < 4; ++i) loads[i] = raw[i] >= 50.0f ? fmaxf(50.0f, con[i] - (con[i] - 50.0f) * rf) : con[i];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) loads[i] = con[i];
    }
}
NCG_HD void tyre_update(float* T, float* wear, float load, float ff, float speed, float alat, float slip) {
    const float dt = NCG_DT;
    float ns = fminf(speed * (1.0f / NCG_CAR_MAX_SPEED), 2.0f);
    float fp = fabsf(ff) * 0.040f * (1.0f + ns * ns);
    float aero = speed > 50.0f ? speed * speed * 0.0002f : 0.0f;
    float heat = (fp + aero) * (dt * (1.0f / 125.0f));
    float ce = 0.01f;
    if (speed > 50.0f) ce *= (1.0f - fminf(0.8f, speed * 0.01f) * 0.5f);
    float t = *T;
    t += heat - (t - 25.0f) * ce * dt;
    t = fmaxf(25.0f, fminf(120.0f, t));
    *T = t;
    float tm = (85.0f <= t && t <= 105.0f) ? 1.0f : (t > 105.0f ? 1.0f + (t - 105.0f) * 0.05f : 1.0f + (85.0f - t) * (1.0f / 30.0f));
    float lm = fmaxf(0.5f, load * (1.0f / NCG_STATIC_TYRE_LOAD));
    float wf = fmaxf(1.0f, t * 0.0125f);
    float sw = fminf(1.0f + (speed * (3.6f / 400.0f)) * 3.0f, 3.0f);
    float lg = fabsf(alat) * (1.0f / 9.81f);
    float cw = lg > 2.0f ? fminf(1.0f + (lg - 2.0f) * 1.5f, 2.5f) : 1.0f;
    float kw = fminf(1.0f + (fabsf(slip) * (1.0f / 45.0f)) * 2.0f, 3.0f);
    float wr = fabsf(ff) * 0.00001f * (tm * lm * wf * sw * cw * kw);
    *wear = fminf(100.0f, *wear + wr * dt);
}

// friction forces for tyre heating: car.py:702-830
NCG_HD void friction_forces(float* ff, const float* R, float driving, float throttle, float brake, float steer_angle, float sp) {
    float rear, front;
    if (throttle > 0.01f && sp > 50.0f) { float faf = fminf(0.3f, sp * 0.005f); rear = driving * (1.0f - faf); front = driving * faf / 2.0f; }
    else { rear = throttle > 0.01f ? driving : 0.0f; front = 0.0f; }
    float bf = 0.0f;
    if (brake > 0.01f) { float base = (NCG_CAR_MASS * 14.0f * 0.25f) * brake; bf = sp <= 1.0f ? base * (0.05f + 0.95f * sp) : base; }
    float rf = sp > 0.1f ? (0.015f * NCG_WEIGHT) / 4.0f : 0.0f;
    ff[0] = ff[1] = front + bf + rf; ff[2] = ff[3] = rear / 2.0f + bf + rf;
    if (sp < 2.0f) return;
    float slip = R[NCG_R_SLIP], m = 1.0f;
    if (slip > 5.0f) { float ex = slip - 5.0f; m = fminf(2.2f + 0.04f * ex * ex, 8.0f); }
    float base = R[NCG_R_FLAT] * 0.05f * m;
    float fl = base * 0.5f / 2.0f, rl = fl;
    if (fabsf(steer_angle) > 0.01f && sp > 3.0f) {
        int of = steer_angle < 0.0f ? 0 : 1, orr = of + 2, inf = 1 - of, inr = inf + 2;
        ff[of] += fl * 1.5f + R[NCG_R_TYRE_LOAD + of] * 0.001f; ff[orr] += rl * 1.5f + R[NCG_R_TYRE_LOAD + orr] * 0.001f;
        ff[inf] += fl * 0.5f; ff[inr] += rl * 0.5f;
    } else { ff[0] += fl; ff[1] += fl; ff[2] += rl; ff[3] += rl; }
}

// chord-nearest segment (first strict minimum): car_physics.py:631-672 (banking) and car_env.py:1544-1611 (progress).
// The reference does this search in float64.  Tracks whose last segment overshoots the start (closure gaps of
// SURVEY App. F) have two chords that coincide to ~1e-14 m near the start line, and the reference's answer there
// hangs on float64 digits -- so a float32 scan picks the winner only when it is clear-cut, and otherwise the
// search is redone in float64 with the reference's exact expressions.
NCG_HDN void nearest_segment64(const Track T, float xf, float yf, float* banking, float* progress) {
    const double x = (double)xf, y = (double)yf;
    double best = INFINITY; int bi = 0; double bcx = 0.0, bcy = 0.0;
    for (int i = 0; i < T.n_segs; ++i) {
        const double* s = T.seg64 + i * SEG64_STRIDE;
        double sx = s[0], sy = s[1], dx = s[2] - s[0], dy = s[3] - s[1];
        double l2 = dx * dx + dy * dy, cx, cy;
        if (l2 < 1e-6) { cx = sx; cy = sy; }
        else {
            double t = ((x - sx) * dx + (y - sy) * dy) / l2;
            t = t > 1.0 ? 1.0 : t; t = t < 0.0 ? 0.0 : t;
            cx = sx + t * dx; cy = sy + t * dy;
        }
        double d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
        if (d2 < best) { best = d2; bi = i; bcx = cx; bcy = cy; }
    }
    const double* s = T.seg64 + bi * SEG64_STRIDE;
    *banking = T.segs[bi * SEG_STRIDE + 9];
    *progress = (float)(s[4] + sqrt((bcx - s[0]) * (bcx - s[0]) + (bcy - s[1]) * (bcy - s[1])));
}
NCG_HD void nearest_segment(const Track& T, float x, float y, float* banking, float* progress) {
    float best = INFINITY, second = INFINITY; int bi = 0; float bcx = 0.0f, bcy = 0.0f;
    // candidate chords of this grid cell (all of them outside the grid): a superset of every chord that can be the
    // minimum or fall inside the float64 tie band below, visited in ascending index like the reference's loop
    uint32_t cand = (1u << T.n_segs) - 1u;
    {
        const int ix = (int)floorf((x - T.gx0) * T.inv_cell), iy = (int)floorf((y - T.gy0) * T.inv_cell);
        if (ix >= 0 && iy >= 0 && ix < T.gnx && iy < T.gny) cand = T.segmask[iy * T.gnx + ix];
    }
    while (cand) {
        const int i = ctz32(cand); cand &= cand - 1u;
        const float* s = T.segs + i * SEG_STRIDE;
        float sx = s[0], sy = s[1], dx = s[4], dy = s[5], l2 = s[6], cx, cy;
        if (l2 < 1e-6f) { cx = sx; cy = sy; }
        else { float t = fmaxf(0.0f, fminf(1.0f, ((x - sx) * dx + (y - sy) * dy) * s[7])); cx = sx + t * dx; cy = sy + t * dy; }
        float d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
        if (d2 < best) { second = best; best = d2; bi = i; bcx = cx; bcy = cy; }
        else if (d2 < second) second = d2;
    }
    // float32 projection error is ~1e-4 m at these coordinates: demand a clear margin in distance, else go to float64
    float db = sqrtf(best), ds = sqrtf(second);
    if (ds - db < 2e-3f + 1e-4f * ds) { float b64, p64; nearest_segment64(T, x, y, &b64, &p64); *banking = b64; *progress = p64; return; }
    const float* s = T.segs + bi * SEG_STRIDE;
    *banking = s[9];
    float px = bcx - s[0], py = bcy - s[1];
    *progress = s[8] + sqrtf(px * px + py * py);
}
// lap_timer.py:205-241
NCG_HD bool on_startline(const Track& T, float x, float y) {
    if (T.slhalfw < 0.0f) return false;
    float d;
    if (T.sllen2 < 1e-6f) { float ax = x - T.slx0, ay = y - T.sly0; d = sqrtf(ax * ax + ay * ay); }
    else {
        float t = fmaxf(0.0f, fminf(1.0f, ((x - T.slx0) * T.sldx + (y - T.sly0) * T.sldy) / T.sllen2));
        float cx = T.slx0 + t * T.sldx, cy = T.sly0 + t * T.sldy;
        d = sqrtf((x - cx) * (x - cx) + (y - cy) * (y - cy));
    }
    return d <= T.slhalfw;
}
// car_physics.py:470-524 (AABB query of +-0.5 m, TestPoint, corner distance)
NCG_HDN bool on_track(const Track T, float x, float y) {
    const float radius = 0.5f;
    AABB q; q.lx = x - radius; q.ly = y - radius; q.ux = x + radius; q.uy = y + radius;
    int ix0 = (int)floorf((q.lx - T.gx0) * T.inv_cell), ix1 = (int)floorf((q.ux - T.gx0) * T.inv_cell);
    int iy0 = (int)floorf((q.ly - T.gy0) * T.inv_cell), iy1 = (int)floorf((q.uy - T.gy0) * T.inv_cell);
    ix0 = ix0 < 0 ? 0 : ix0; iy0 = iy0 < 0 ? 0 : iy0; ix1 = ix1 >= T.gnx ? T.gnx - 1 : ix1; iy1 = iy1 >= T.gny ? T.gny - 1 : iy1;
    for (int iy = iy0; iy <= iy1; ++iy) for (int ix = ix0; ix <= ix1; ++ix) {
        int cell = iy * T.gnx + ix;
        const int nk = cell_count_max(T, cell);
        for (int k = 0; k < nk; ++k) {
            int wi = cell_item(T, cell, k);
            if (!aabb_overlap(q, wall_fat(T, wi))) continue;
            Xf xf; Box b; wall_get(T, wi, &xf, &b);
            V2 pl = mulT(xf.q, mk(x, y) - xf.p);
            bool inside = true;
            for (int i = 0; i < 4; ++i) if (dot(box_n(i), pl - box_v(b, i)) > 0.0f) { inside = false; break; }
            if (inside) return false;
            for (int i = 0; i < 4; ++i) { V2 v = mul(xf, box_v(b, i)); float dx = x - v.x, dy = y - v.y; if (sqrtf(dx * dx + dy * dy) < radius) return false; }
        }
    }
    return true;
}

// ------------------------------------------------------------------ reset (car_env.py:316-535)
// fresh: new Box2D world (CarPhysics.__init__); otherwise CarPhysics.reset_car + Car.reset on the existing world.
NCG_HD void reset_record(float* R, const Track& T, bool fresh, uint32_t track_id, const StartPose sp) {
    uint32_t fl;
    if (fresh) {
        for (int i = 0; i < NCG_RECORD_WORDS; ++i) R[i] = 0.0f;
        Xf xf; xf.p = mk(sp.x, sp.y); xf.q = rot(sp.a);
        R[NCG_R_X] = sp.x; R[NCG_R_Y] = sp.y; R[NCG_R_ANGLE] = sp.a;
        AABB a = box_aabb(car_box(), xf);
        R[NCG_R_FAT_LX] = a.lx - NCG_B2_AABB_EXT; R[NCG_R_FAT_LY] = a.ly - NCG_B2_AABB_EXT; R[NCG_R_FAT_UX] = a.ux + NCG_B2_AABB_EXT; R[NCG_R_FAT_UY] = a.uy + NCG_B2_AABB_EXT;
        fl = NCG_F_AWAKE | NCG_F_PROXY_MOVED | NCG_F_NEW_FIXTURE;
    } else {
        Body B; b_load(B, R);
        V2 p = mk(sp.x, sp.y);
        b_set_transform(B, p, B.sweep.a); b_set_transform(B, p, sp.a);
        B.v = mk(0.0f, 0.0f); B.w = 0.0f;     // SetLinearVelocity(0)/SetAngularVelocity(0) do not wake
        B.impulse = 0.0f; B.hasKey = false;
        fl = f2u(R[NCG_R_FLAGS]) & (NCG_F_OVERFLOW);
        b_store(B, R, &fl);
        R[NCG_R_NCONTACT] = u2f(f2u(R[NCG_R_NCONTACT]) & ~0xFF00u);      // active_collisions.clear(); contacts persist
        fl &= (NCG_F_AWAKE | NCG_F_PROXY_MOVED | NCG_F_NEW_FIXTURE | NCG_F_OVERFLOW);
        for (int i = NCG_R_RPM; i < NCG_R_USED; ++i) if (i != NCG_R_BANK) R[i] = 0.0f;
    }
    fl |= NCG_F_FIRST_STEP;
    R[NCG_R_FLAGS] = u2f(fl);
    R[NCG_R_RPM] = 1000.0f;
    for (int i = 0; i < 4; ++i) { R[NCG_R_TYRE_TEMP + i] = 80.0f; R[NCG_R_TYRE_WEAR + i] = 0.0f; R[NCG_R_TYRE_LOAD + i] = NCG_STATIC_TYRE_LOAD; }
    R[NCG_R_ACC_N] = u2f(0u); R[NCG_R_STUCK_STEPS] = u2f(0u); R[NCG_R_LAP_START] = u2f(0u); R[NCG_R_LAP_COUNT] = u2f(0u);
    R[NCG_R_STEP] = u2f(0u); R[NCG_R_TRACK] = u2f(track_id);
    float bank, prog; nearest_segment(T, R[NCG_R_X], R[NCG_R_Y], &bank, &prog);
    R[NCG_R_PROGRESS_PREV] = prog; R[NCG_R_PREV_X] = R[NCG_R_X]; R[NCG_R_PREV_Y] = R[NCG_R_Y];
}

// ------------------------------------------------------------------ observation words 0..21 (car_env.py:891-946)
NCG_HD float clip1(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
// Observation word k (< 22) = clip(raw_k * scale_k, lo_k, 1).  observe_raw gathers the 22 raw values on the physics
// warp; the scaling and clipping can then be done by whoever stores the row (the ray warps, spread over lanes).
#define NCG_OBS_STATE_DIM 22
NCG_HD float obs_scale(int k) {
    const float t[NCG_OBS_STATE_DIM] = {1e-4f, 1e-4f, 1.0f / 111.1f, 1.0f / 111.1f, 1.0f / 111.1f, 1.0f / 3.14159265358979f, 0.1f,
                                        1.0f / 29430.0f, 1.0f / 29430.0f, 1.0f / 29430.0f, 1.0f / 29430.0f, 0.005f, 0.005f, 0.005f, 0.005f,
                                        0.01f, 0.01f, 0.01f, 0.01f, 1.0f / 50000.0f, 1.0f / 3.14159265358979f, 1.0f / 250000.0f};
    return k < NCG_OBS_STATE_DIM ? t[k] : 1.0f;
}
NCG_HD float obs_lo(int k) { return (k <= 3 || k == 5 || k == 6 || k == 20) ? -1.0f : 0.0f; }
NCG_HD float obs_word(float raw, float scale, float lo) { return clip1(raw * scale, lo, 1.0f); }
// direction of the last collision relative to the heading, wrapped to [-pi, pi]; out of line: it runs on the few steps
// with a hard impact, and inline its atan2f is 120 instructions the physics warp has to jump over on every other step
NCG_HDN float collision_angle(float ny, float nx, float heading) {
    float ca = atan2f(ny, nx) - heading;
    while (ca > 3.14159265358979f) ca -= 6.28318530717959f;
    while (ca < -3.14159265358979f) ca += 6.28318530717959f;
    return ca;
}
NCG_HD void observe_raw(const float* R, float* raw) {
    float vx = R[NCG_R_VX], vy = R[NCG_R_VY];
    raw[0] = R[NCG_R_X]; raw[1] = R[NCG_R_Y]; raw[2] = vx; raw[3] = vy; raw[4] = sqrtf(vx * vx + vy * vy);
    raw[5] = R[NCG_R_ANGLE]; raw[6] = R[NCG_R_OMEGA];
#pragma unroll
    for (int k = 0; k < 4; ++k) { raw[7 + k] = R[NCG_R_TYRE_LOAD + k]; raw[11 + k] = R[NCG_R_TYRE_TEMP + k]; raw[15 + k] = R[NCG_R_TYRE_WEAR + k]; }
    float ci = R[NCG_R_IMPULSE], ca = 0.0f;
    if (ci < 100.0f) ci = 0.0f;
    else if (((f2u(R[NCG_R_NCONTACT]) >> 8) & 255u) > 0) ca = collision_angle(R[NCG_R_ACTIVE + 2], R[NCG_R_ACTIVE + 1], R[NCG_R_ANGLE]);
    raw[19] = ci; raw[20] = ca; raw[21] = R[NCG_R_CUM_IMPACT];
}
NCG_HD void observe_state(const float* R, float* obs) {
    float raw[NCG_OBS_STATE_DIM]; observe_raw(R, raw);
#pragma unroll
    for (int k = 0; k < NCG_OBS_STATE_DIM; ++k) obs[k] = obs_word(raw[k], obs_scale(k), obs_lo(k));
}

// ------------------------------------------------------------------ the scalar car phase of one step
// Runs CarPhysics.step and every per-car part of CarEnv._step_multi_car up to (not including) the env-level
// termination, in two halves so the kernel can hand the new pose to the ray warps as soon as it exists:
//   car_step_dynamics  inputs -> forces -> tyres -> b2World.Step; the record holds the new pose when it returns
//   car_step_rules     banking/progress, impact and stuck rules, lap timer, obs[0..21], reward; returns the reward,
//                      *xflags gets NCG_X_* bits for the env phase.
// the acceleration window while it is still filling (an episode's first nine steps); out of line, see collision_angle
NCG_HDN void acc_window_filling(float* R, int n, float lo, float la, float* along, float* alat) {
    R[NCG_R_ACC + 2 * n] = lo; R[NCG_R_ACC + 2 * n + 1] = la; ++n;
    R[NCG_R_ACC_N] = u2f((uint32_t)n);
    float s0 = 0.0f, s1 = 0.0f;
    for (int i = 0; i < n; ++i) { s0 += R[NCG_R_ACC + 2 * i]; s1 += R[NCG_R_ACC + 2 * i + 1]; }
    *along = s0 / (float)n; *alat = s1 / (float)n;
}
// lateral force of a banked segment (car.py:509-566): m g sin(|banking|) * 0.3
NCG_HDN float banking_force(float bank) {
    float sb, cb; sincos_heading(fabsf(bank * 0.017453292519943295f), &sb, &cb);
    return NCG_CAR_MASS * 9.81f * sb * 0.3f;
}
struct StepCtx { uint32_t fl, xf, laps_pre; bool dis_pre; float impulse; };
// car_dyn_pre: everything of CarPhysics.step before b2World.Step (inputs, forces, tyres): leaves the loaded body with its
// force / torque in *Wout.  car_dyn_post: the record takes the stepped body.  car_step_dynamics = pre, body_step, post; the
// shared-world mode (cars of an env collide, optional) steps all bodies of an env between the two halves.
struct DynPre { uint32_t fl, xf, laps_pre; bool dis_pre; };
NCG_HD void car_dyn_pre(float* R, const Track& T, float thr_in, float brk_in, float steer_in, Body* Wout, DynPre* pre) {
    uint32_t fl = f2u(R[NCG_R_FLAGS]);
    uint32_t xf = 0;
    const bool dis_pre = (fl & NCG_F_DISABLED) != 0;
    uint32_t laps_pre = f2u(R[NCG_R_LAP_COUNT]);
    if (dis_pre) { xf |= NCG_X_DIS_PRE; thr_in = 0.0f; brk_in = 0.0f; steer_in = 0.0f; }
    if (laps_pre >= 1u) xf |= NCG_X_LAP_PRE;
    // ---- Car.set_inputs + update_physics (car.py:241-387)
    float throttle = fmaxf(0.0f, fminf(1.0f, thr_in)), brake = fmaxf(0.0f, fminf(1.0f, brk_in));
    float steer = fmaxf(-1.0f, fminf(1.0f, steer_in));
    float delta = steer * 0.78539816339744831f;
    Body W; b_load(W, R);
    {   // rpm :311-327
        float rpm = R[NCG_R_RPM];
        float diff = (1000.0f + 800.0f * throttle) - rpm;
        rpm += diff * fminf(1.0f, NCG_DT * 3000.0f / fabsf(diff + 0.1f));
        R[NCG_R_RPM] = fmaxf(600.0f, fminf(9500.0f, rpm));
    }
    {   // engine :389-449
        float cs = length(W.v);
        float r = fmaxf(1000.0f, fminf(9000.0f, R[NCG_R_RPM]));
        float tf = r <= 5500.0f ? 0.7f + (r - 1000.0f) * (0.3f / 4500.0f) : 1.0f - (r - 5500.0f) * (0.6f / 3500.0f);
        float tlf = (NCG_CAR_MAX_TORQUE * tf) * throttle * (7.5f / 0.35f);
        float ef;
        if (cs > 12.0f) {
            float plf = (NCG_CAR_MAX_POWER * throttle) / cs;
            if (cs <= 25.0f) {
                float ns = (cs - 12.0f) * (1.0f / 13.0f);
                float b = fmaxf(0.05f, fminf(0.75f, 1.0f - expf(-2.0f * ns)));
                ef = tlf * (1.0f - b) + plf * b;
            } else ef = tlf * 0.25f + plf * 0.75f;
        } else ef = tlf;
        float grip = total_grip(R);
        float sf = 1.0f - fminf(0.4f, fabsf(delta) * 1.5f);
        ef = fminf(ef, NCG_WEIGHT * grip * sf);
        V2 fwd = mul(W.xf.q, mk(1.0f, 0.0f));
        float ff[4]; friction_forces(ff, R, fminf(2000.0f, fabsf(ef) / 2.0f), throttle, brake, delta, cs);
        V2 rear = mul(W.xf, mk(-NCG_CAR_WHEELBASE / 2.0f, 0.0f));
        b_apply_force(W, mk(ef * fwd.x, ef * fwd.y), rear);
        // brake :451-469
        if (brake > 0.01f) {
            float bsf = 1.0f - fminf(0.3f, fabsf(delta) * 1.5f);
            float bforce = NCG_CAR_MASS * 14.0f * bsf * brake;
            float sp = length(W.v);
            if (sp > 0.1f) {
                { float is = 1.0f / sp; b_apply_force_center(W, mk(bforce * (-W.v.x * is), bforce * (-W.v.y * is))); }
                friction_forces(ff, R, 0.0f, throttle, brake, delta, sp);
            }
        }
        // drag :471-484, rolling :486-500
        float sp = length(W.v);
        if (sp > 0.1f) {
            float mag = NCG_DRAG_CONSTANT * sp * sp, is = 1.0f / sp;
            b_apply_force_center(W, mk(mag * (-W.v.x * is), mag * (-W.v.y * is)));
            float rr = 0.015f * NCG_WEIGHT;
            b_apply_force_center(W, mk(rr * (-W.v.x * is), rr * (-W.v.y * is)));
        }
        // acceleration window :832-892
        float along, alat;
        {
            float ax = (W.v.x - R[NCG_R_PREV_VX]) * 60.0f, ay = (W.v.y - R[NCG_R_PREV_VY]) * 60.0f;
            V2 f = mul(W.xf.q, mk(1.0f, 0.0f)), l = mul(W.xf.q, mk(0.0f, 1.0f));
            float lo = fmaxf(-12.0f, fminf(12.0f, ax * f.x + ay * f.y));
            float la = fmaxf(-12.0f, fminf(12.0f, ax * l.x + ay * l.y));
            int n = (int)f2u(R[NCG_R_ACC_N]);
            if (n == 10) {
                // full window (every step but an episode's first nine): drop the oldest pair, append the new one and sum,
                // through registers -- one batch of loads, the same left-to-right float32 sums, one batch of stores
                float a[20];
#pragma unroll
                for (int i = 0; i < 18; ++i) a[i] = R[NCG_R_ACC + i + 2];
                a[18] = lo; a[19] = la;
                float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                for (int i = 0; i < 10; ++i) { s0 += a[2 * i]; s1 += a[2 * i + 1]; }
#pragma unroll
                for (int i = 0; i < 20; ++i) R[NCG_R_ACC + i] = a[i];
                along = s0 / 10.0f; alat = s1 / 10.0f;
            } else acc_window_filling(R, n, lo, la, &along, &alat);
            R[NCG_R_PREV_VX] = W.v.x; R[NCG_R_PREV_VY] = W.v.y;
        }
        // tyres :354-357
        {
            float loads[4]; weight_transfer(along, alat, sp, loads);
            float slip_prev = R[NCG_R_SLIP];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                R[NCG_R_TYRE_LOAD + i] = loads[i];
                tyre_update(&R[NCG_R_TYRE_TEMP + i], &R[NCG_R_TYRE_WEAR + i], loads[i], ff[i], sp, alat, slip_prev);
            }
        }
        // lateral alignment force :635-700
        if (sp > 0.05f) {
            V2 f = mul(W.xf.q, mk(1.0f, 0.0f));
            float is = 1.0f / sp; float vnx = W.v.x * is, vny = W.v.y * is;
            float cp = vnx * f.y - vny * f.x, dp = vnx * f.x + vny * f.y;
            R[NCG_R_SLIP] = fabsf(atan2f(fabsf(cp), dp)) * 57.295779513082323f;
            float cfx = (f.x * sp - W.v.x) * (NCG_CAR_MASS * 5.0f), cfy = (f.y * sp - W.v.y) * (NCG_CAR_MASS * 5.0f);
            float g2 = total_grip(R);
            float mf = fminf(30000.0f * g2, NCG_WEIGHT * g2);
            float fm = sqrtf(cfx * cfx + cfy * cfy);
            if (fm > mf) { float sc = mf / fm; cfx *= sc; cfy *= sc; fm = mf; }
            R[NCG_R_FLAT] = fm;
            b_apply_force_center(W, mk(cfx, cfy));
        }
        // angular damping :502-507
        b_apply_torque(W, -W.w * NCG_CAR_MASS * 4.0f);
        // banking :509-566
        float bank = R[NCG_R_BANK];
        if (!(fabsf(bank) < 0.1f) && !(sp < 1.0f)) {
            const float la = banking_force(bank);      // (out of line: flat tracks never get here)
            if (!(fabsf(la) < 1.0f) && sp > 5.0f) {
                float sg = bank < 0.0f ? -1.0f : 1.0f;
                { float is = 1.0f / sp; b_apply_force_center(W, mk((-(W.v.y * is)) * la * sg, (W.v.x * is) * la * sg)); }
            }
        }
        // steering :568-584
        if (fabsf(delta) > 0.01f && sp > 0.1f) {
            float sd, cd; sincos_heading(delta, &sd, &cd);                 // |delta| <= pi/4; tan = s / c without the library's inline slow path
            float dav = sp * (sd / cd) * (1.0f / NCG_CAR_WHEELBASE);
            b_apply_torque(W, (dav - W.w) * NCG_CAR_MASS * 0.8f);
        }
    }
    *Wout = W; pre->fl = fl; pre->xf = xf; pre->laps_pre = laps_pre; pre->dis_pre = dis_pre;
}
NCG_HD void car_dyn_post(float* R, Body& W, const DynPre& pre, StepCtx* ctx, Counters* cnt) {
    uint32_t fl = pre.fl;
    b_store(W, R, &fl);
    if (W.overflow) cnt->overflow++;
    ctx->fl = fl; ctx->xf = pre.xf; ctx->laps_pre = pre.laps_pre; ctx->dis_pre = pre.dis_pre; ctx->impulse = W.impulse;
}
NCG_HD void car_step_dynamics(float* R, const Track& T, float thr_in, float brk_in, float steer_in, int contacts, StepCtx* ctx,
                               Counters* cnt) {
    Body W; DynPre pre;
    car_dyn_pre(R, T, thr_in, brk_in, steer_in, &W, &pre);
    // ---- b2World.Step (car_physics.py:363)
    body_step(W, R, T, NCG_DT, contacts, cnt);
    car_dyn_post(R, W, pre, ctx, cnt);
}
// RAW: obs[0..21] receives observe_raw's values (the caller normalises them when it stores the row)
template <bool RAW>
NCG_HD float car_step_rules(float* R, const Track& T, const StepCtx* ctx, float* obs, uint32_t* xflags, Counters* cnt) {
    uint32_t fl = ctx->fl, xf = ctx->xf;
    const uint32_t laps_pre = ctx->laps_pre;
    const bool dis_pre = ctx->dis_pre;
    struct { float impulse; } W; W.impulse = ctx->impulse;
    const float x = R[NCG_R_X], y = R[NCG_R_Y];
    const float speed = length(mk(R[NCG_R_VX], R[NCG_R_VY]));
    // banking refresh + progress (one segment scan serves both)
    float bank_new, progress;
    nearest_segment(T, x, y, &bank_new, &progress);
    R[NCG_R_BANK] = T.has_bank ? bank_new : 0.0f;
    if (speed > R[NCG_R_MAX_SPEED]) R[NCG_R_MAX_SPEED] = speed;
    // ---- _run_single_physics_step (car_env.py:581-638)
    bool disabled = dis_pre, jd = false;
    uint32_t stuck = f2u(R[NCG_R_STUCK_STEPS]);
    if (!dis_pre) {
        float ci = W.impulse;
        if (ci > 50000.0f) { if (!disabled) { disabled = true; jd = true; } }
        if (ci > 100.0f) R[NCG_R_CUM_IMPACT] += ci;
        if (R[NCG_R_CUM_IMPACT] > 250000.0f) { if (!disabled) { disabled = true; jd = true; } }
        if (speed < 0.5f) stuck += 1u; else { stuck = 0u; fl &= ~(uint32_t)NCG_F_STUCK_POS; }
    }
    if (disabled) xf |= NCG_X_DIS_MID;
    // ---- LapTimer.update (lap_timer.py:95-203) with the pre-increment clock
    uint32_t step_pre = f2u(R[NCG_R_STEP]);
    {
        bool has_pos = (fl & NCG_F_HAS_POS) != 0;
        float lx = R[NCG_R_LAP_X], ly = R[NCG_R_LAP_Y];
        if (has_pos) { float dx = x - lx, dy = y - ly; float d = sqrtf(dx * dx + dy * dy); if (d < 50.0f) R[NCG_R_ODO] += d; }
        if (has_pos && T.slhalfw >= 0.0f) {
            bool now = on_startline(T, x, y), before = on_startline(T, lx, ly);
            if (now && !before) {
                if (fl & NCG_F_CROSSED) {
                    uint32_t lap_steps = step_pre - f2u(R[NCG_R_LAP_START]);
                    if (lap_steps >= NCG_MIN_LAP_STEPS && !(R[NCG_R_ODO] < T.min_lap)) {
                        float lt = (float)lap_steps * (1.0f / 60.0f);
                        R[NCG_R_LAST_LAP] = lt;
                        if (!(fl & NCG_F_HAS_BEST) || lt < R[NCG_R_BEST_LAP]) { R[NCG_R_BEST_LAP] = lt; fl |= NCG_F_HAS_BEST; }
                        fl |= NCG_F_HAS_LAST;
                        R[NCG_R_LAP_START] = u2f(step_pre);
                        R[NCG_R_LAP_COUNT] = u2f(laps_pre + 1u); R[NCG_R_ODO] = 0.0f;
                        xf |= NCG_X_COMPLETED; cnt->laps++;
                    }
                } else { fl |= NCG_F_CROSSED; R[NCG_R_LAP_START] = u2f(step_pre); R[NCG_R_ODO] = 0.0f; }
            }
        }
        R[NCG_R_LAP_X] = x; R[NCG_R_LAP_Y] = y; fl |= NCG_F_HAS_POS;
    }
    if (f2u(R[NCG_R_LAP_COUNT]) >= 1u) xf |= NCG_X_LAP_MID;
    R[NCG_R_STEP] = u2f(step_pre + 1u);
    // ---- _check_and_disable_cars (car_env.py:805-888)
    if (!disabled) {
        if (speed < 0.5f && stuck > 0u) {
            if (!(fl & NCG_F_STUCK_POS)) { fl |= NCG_F_STUCK_POS; R[NCG_R_STUCK_X] = x; R[NCG_R_STUCK_Y] = y; }
            float dx = x - R[NCG_R_STUCK_X], dy = y - R[NCG_R_STUCK_Y];
            float moved = sqrtf(dx * dx + dy * dy);
            if (stuck >= NCG_STUCK_STEPS) { if (moved < 1.0f || stuck >= NCG_STUCK_EXT_STEPS) { disabled = true; jd = true; } }
        } else { stuck = 0u; fl &= ~(uint32_t)NCG_F_STUCK_POS; }
    }
    R[NCG_R_STUCK_STEPS] = u2f(stuck);
    // ---- observation words 0..21 (before the end-of-step impulse reset)
    if (RAW) observe_raw(R, obs); else observe_state(R, obs);
    // ---- reward (car_env.py:980-1113)
    float reward = 0.0f;
    if (!(disabled && !jd)) {
        float r = jd ? 10.0f : 0.0f;
        if (!disabled) { r -= 0.05f; if (fabsf(W.impulse) > 0.0f) r -= 0.5f; }
        {
            float dx = x - R[NCG_R_PREV_X], dy = y - R[NCG_R_PREV_Y];
            r += sqrtf(dx * dx + dy * dy) * 0.15f;
            R[NCG_R_PREV_X] = x; R[NCG_R_PREV_Y] = y;
        }
        if (!(fl & NCG_F_FIRST_STEP)) {
            float dl = progress - R[NCG_R_PROGRESS_PREV];
            if (dl > T.half_ltot) dl -= T.ltot; else if (dl < -T.half_ltot) dl += T.ltot;
            if (dl < 0.0f) {
                float back = R[NCG_R_BACK] + fabsf(dl);
                R[NCG_R_BACK] = back;
                if (back > 200.0f) { if (!disabled) { disabled = true; jd = true; } }
                if (back > 25.0f) {
                    fl |= NCG_F_BACK_ACTIVE;
                    float nb = fmaxf(0.0f, back - 25.0f) - fmaxf(0.0f, R[NCG_R_BACK_PREV] - 25.0f);
                    if (nb > 0.0f) r -= nb * 0.05f;
                }
            } else { R[NCG_R_BACK] = 0.0f; R[NCG_R_BACK_PREV] = 0.0f; fl &= ~(uint32_t)NCG_F_BACK_ACTIVE; }
            R[NCG_R_PROGRESS_PREV] = progress;
        } else { R[NCG_R_PROGRESS_PREV] = progress; fl &= ~(uint32_t)NCG_F_FIRST_STEP; }
        if (!disabled) R[NCG_R_BACK_PREV] = R[NCG_R_BACK];
        reward = r;
    }
    if (disabled) { fl |= NCG_F_DISABLED; xf |= NCG_X_DIS_POST; }
    if (jd) xf |= NCG_X_JUST_DISABLED;
    if (R[NCG_R_CUM_REWARD] < -250.0f) xf |= NCG_X_LOW_REWARD;
    R[NCG_R_FLAGS] = u2f(fl);
    *xflags = xf;
    return reward;
}
NCG_HD float car_step(float* R, const Track& T, float thr_in, float brk_in, float steer_in, int contacts, float* obs,
                      uint32_t* xflags, Counters* cnt) {
    StepCtx ctx;
    car_step_dynamics(R, T, thr_in, brk_in, steer_in, contacts, &ctx, cnt);
    return car_step_rules<false>(R, T, &ctx, obs, xflags, cnt);
}

// ------------------------------------------------------------------ env phase (car_env.py:672-676, 1115-1158, 773-797)
// xf[0..C) are the cars' NCG_X_* words of this step, step_after the env clock after its increment.
// reason: 0 none, 1 all_cars_disabled, 2 all_active_cars_low_reward, 3 time_limit, 4 truncated.
NCG_HD void env_decide(const uint32_t* xf, int C, bool reset_on_lap, uint32_t step_after, bool* term, bool* trunc, int* reason) {
    // _lap_reset_pending: evaluated when car i's timer fires, i.e. with cars j<=i already stepped and j>i not yet
    bool pending = false;
    if (reset_on_lap) {
        for (int i = 0; i < C; ++i) {
            if (!(xf[i] & NCG_X_COMPLETED)) continue;
            int active = 0; bool all = true;
            for (int j = 0; j < C; ++j) {
                bool dis = j <= i ? (xf[j] & NCG_X_DIS_MID) != 0 : (xf[j] & NCG_X_DIS_PRE) != 0;
                bool lap = j <= i ? (xf[j] & NCG_X_LAP_MID) != 0 : (xf[j] & NCG_X_LAP_PRE) != 0;
                if (dis) continue;
                ++active; if (!lap) all = false;
            }
            if (active > 0 && all) pending = true;
        }
    }
    bool te = false, tr = false; int why = 0;
    int nd = 0, active = 0, below = 0;
    for (int i = 0; i < C; ++i) { if (xf[i] & NCG_X_DIS_POST) ++nd; else { ++active; if (xf[i] & NCG_X_LOW_REWARD) ++below; } }
    if (nd >= C) { te = true; why = 1; }
    else if (active > 0 && below == active) { te = true; why = 2; }
    else if (reset_on_lap && step_after >= NCG_TERMINATE_STEPS) { te = true; why = 3; }
    else if (step_after >= NCG_TRUNCATE_STEPS) { tr = true; why = 4; }
    if (pending) te = true;
    *term = te; *trunc = tr; *reason = why;
}
// per-car end of step: cumulative reward (float32, as numpy 2 accumulates it), listener impulse reset
NCG_HD void car_finish(float* R, float reward) {
    R[NCG_R_CUM_REWARD] = R[NCG_R_CUM_REWARD] + reward;
    R[NCG_R_IMPULSE] = 0.0f;
    R[NCG_R_FLAGS] = u2f(f2u(R[NCG_R_FLAGS]) | NCG_F_HAS_KEY);
}

NCG_HD float sensor_obs_m(float dist) { float n = dist * 0.004f; return n < 0.0f ? 0.0f : (n > 1.0f ? 1.0f : n); }
// ------------------------------------------------------------------ sensor rays (distance_sensor.py:71-117)
// 16 rays per car from the body origin, direction = heading - i*22.5 deg, length 250 m, nearest wall-box entry.
// The reference goes through b2World::RayCast -> b2PolygonShape::RayCast per fixture; here a uniform-grid DDA
// (per-track cell size, lists in the staged track table) visits the boxes near the ray and each candidate gets a
// slab test in the box frame: entry/exit distances along the unit direction, MUFU.RCP reciprocals and FFMA products.
// That is the same intersection as Box2D's half-plane clipping (a hit needs an entry crossing at t > 0, so an
// origin inside a box reports nothing for that box) evaluated to ~1e-6 relative instead of bit-for-bit: rays only
// feed obs[22..37], whose stated tolerance is 1e-3 normalised (0.25 m); tests/hostcheck compares this traversal
// with Box2D's own clipping arithmetic over all walls.
NCG_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;     // one MUFU.RCP (no range fix-up code)
#else
    return 1.0f / x;
#endif
}
// Loads of the ray loop.  SH = the track table is staged in shared memory: the table pointers are turned into 32-bit
// shared-window addresses once per call and the loop uses ld.shared (LDS) with integer address arithmetic, instead of
// the generic-address loads the compiler emits for a pointer it cannot trace to a __shared__ object.
template <bool SH> struct RayMem {
#if defined(__CUDA_ARCH__)
    unsigned walls, cells, items; const Track* t;
    __device__ __forceinline__ RayMem(const Track& T) : t(&T) {
        if (SH) { walls = (unsigned)__cvta_generic_to_shared(T.walls); cells = (unsigned)__cvta_generic_to_shared(T.cells); items = (unsigned)__cvta_generic_to_shared(T.items); }
    }
    __device__ __forceinline__ void wall(uint32_t wi, F4* a, F4* b) const {
        NCG_CHECK(wi < (uint32_t)t->n_walls, "ray loop: wall index");
        if (SH) {
            const unsigned ad = walls + wi * (WALL_STRIDE * 4u);
            asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w) : "r"(ad));
            asm("ld.shared.v2.f32 {%0,%1}, [%2+16];" : "=f"(b->x), "=f"(b->y) : "r"(ad));      // the rays' half extents only
        } else { const float* w = t->walls + wi * WALL_STRIDE; *a = *reinterpret_cast<const F4*>(w); b->x = w[4]; b->y = w[5]; }
    }
    __device__ __forceinline__ uint32_t cell(int c) const {
        NCG_CHECK(c >= 0 && c < t->gnx * t->gny, "ray loop: grid cell");
        if (SH) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(cells + 4u * (unsigned)c)); return v; }
        return t->cells[c];
    }
    __device__ __forceinline__ void block(int k, uint32_t* lo, uint32_t* hi) const {
        // (k may be one past a cell's list -- the one-block-ahead fetch -- but never past the item array plus its pad block)
        NCG_CHECK(k >= 0 && 4 * k <= (int)f2u(t->hdr[TH_NITEMS]), "ray loop: item block");
        if (SH) { asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(*lo), "=r"(*hi) : "r"(items + 8u * (unsigned)k)); return; }
        const uint32_t* q = reinterpret_cast<const uint32_t*>(t->items + 4 * k); *lo = q[0]; *hi = q[1];
    }
#else
    const Track* t;
    RayMem(const Track& T) : t(&T) {}
    void wall(uint32_t wi, F4* a, F4* b) const { const float* w = t->walls + wi * WALL_STRIDE; *a = *reinterpret_cast<const F4*>(w); b->x = w[4]; b->y = w[5]; }
    uint32_t cell(int c) const { return t->cells[c]; }
    void block(int k, uint32_t* lo, uint32_t* hi) const { const uint32_t* q = reinterpret_cast<const uint32_t*>(t->items + 4 * k); *lo = q[0]; *hi = q[1]; }
#endif
};
// entry distance (metres) of the ray origin (px,py), unit direction (dx,dy) into the wall row (wa, wb), or +inf
NCG_HD float ray_box_slab(const F4 wa, const F4 wb, float px, float py, float dx, float dy) {
    const float c = wa.z, s = wa.w;
    const float ax = wa.x - px, ay = wa.y - py;
    const float mx = fmaf(c, ax, s * ay), my = fmaf(c, ay, -(s * ax));      // box centre seen from the origin, box frame
    // direction in the box frame; the 1e-30 keeps a ray parallel to a face off 0*inf = NaN when its origin lies
    // exactly in the face's plane (every car starts at x = 0, where two wall boxes abut)
    const float ex = fmaf(c, dx, fmaf(s, dy, 1e-30f)), ey = fmaf(c, dy, fmaf(-s, dx, 1e-30f));
    const float ix = rcp_fast(ex), iy = rcp_fast(ey);
    // entry / exit along each box axis: (m -+ h) / e = m/e -+ |h/e| -- one product and two FFMAs per axis, no min/max pairs.
    // wb.x is the half-length plus 0.1 mm (the exact one, for contacts, is word 7 of the row): with m = h exactly (origin in the plane of an end face, ray parallel to it)
    // m/e - |h/e| is the rounding error of one product times 1e30, of either sign, and a ray in the plane of the joint
    // between two collinear boxes could miss both; with the margin it enters both
    const float hxi = fabsf(wb.x * ix), hyi = fabsf(wb.y * iy);
    const float tn = fmaxf(fmaf(mx, ix, -hxi), fmaf(my, iy, -hyi)), tf = fminf(fmaf(mx, ix, hxi), fmaf(my, iy, hyi));
#if defined(__CUDA_ARCH__)
    float r;                                                  // (tn > 0 && tn <= tf) ? tn : +inf as two FSETP (the second takes the first as input) and one FSEL
    asm("{ .reg .pred q, p; setp.le.f32 q, %1, %2; setp.gt.and.f32 p, %1, 0f00000000, q; selp.f32 %0, %1, 0f7F800000, p; }" : "=f"(r) : "f"(tn), "f"(tf));
    return r;
#else
    return (tn > 0.0f && tn <= tf) ? tn : INFINITY;
#endif
}
#define NCG_RAY_LEN 250.0f
// the unit direction of a sensor ray: the heading (ca, sa) rotated by the ray's constant (kc, ks), with the roundings
// pinned (one rounded product, one FMA per component) so that every kernel shape and both ray schedulers get the same
// bits whatever the compiler would otherwise contract
NCG_HD void ray_dir(float ca, float sa, float kc, float ks, float* dx, float* dy) {
#if defined(__CUDA_ARCH__)
    *dx = fmaf(ca, kc, -__fmul_rn(sa, ks)); *dy = fmaf(sa, kc, __fmul_rn(ca, ks));
#else
    const volatile float p0 = sa * ks, p1 = ca * ks;
    *dx = fmaf(ca, kc, -p0); *dy = fmaf(sa, kc, p1);
#endif
}
// One lane's RPL rays of one car (see RaySet: a lane's total work mixes along-track and across-track rays).  All RPL
// rays run in one flattened loop -- every iteration is
// "fetch the next cell if this one's list is exhausted, then test one block of four walls" -- so lanes never sit
// in different loop nests, and the four slab tests of a block are independent instruction streams that overlap
// their shared-memory and MUFU latencies.  Normalised distances go to dst[q0 + 4*j].
// cos/sin(-q*22.5 deg): the reference evaluates cos/sin(theta - i*pi/8) in float64 (distance_sensor.py:95-103);
// rotating the float32 heading by a constant keeps the axis-aligned rays of the start pose exactly axis-aligned.
NCG_HD void ray_rotation(int q, float* kcq, float* ksq) {
    const float kc[16] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f, -0.38268343236508977f,
                          -0.70710678118654752f, -0.92387953251128674f, -1.0f, -0.92387953251128674f, -0.70710678118654752f,
                          -0.38268343236508977f, 0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
    const float ks[16] = {0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.0f, -0.92387953251128674f,
                          -0.70710678118654752f, -0.38268343236508977f, 0.0f, 0.38268343236508977f, 0.70710678118654752f,
                          0.92387953251128674f, 1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
    *kcq = kc[q]; *ksq = ks[q];
}
// The rays one lane casts: RPL ray indices and their rotation constants (fixed per lane, so the kernel builds this once
// per launch).  RPL = 2: rays q and q+4; RPL = 4: rays q, q+4 and the pair opposite to the *other* diagonal,
// ((q+2)&3)+8 and +12 -- a car's two longest rays are the ones along the track, 180 deg apart (r and r+8), and this
// keeps them in different lanes.
template <int RPL> struct RaySet { int idx[RPL]; float kc[RPL], ks[RPL]; };
template <int RPL> NCG_HD RaySet<RPL> ray_set(int q) {
    RaySet<RPL> rs;
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
        rs.idx[j] = RPL == 4 ? (j < 2 ? q + 4 * j : ((q + 2) & 3) + 4 * j) : q + 4 * j;
        ray_rotation(rs.idx[j], &rs.kc[j], &rs.ks[j]);
    }
    return rs;
}
template <int RPL, bool SH>
NCG_HD void cast_rays(const Track& T, float px, float py, float angle, const RaySet<RPL>& rs, float* dst, unsigned* tests) {
    float sa, ca; sincos_heading(angle, &sa, &ca);
    float dx, dy; ray_dir(ca, sa, rs.kc[0], rs.ks[0], &dx, &dy);
    unsigned nt = 0;
    const RayMem<SH> M(T);
    const int gnx = T.gnx, gny = T.gny;
    const float gx = (px - T.gx0) * T.inv_cell, gy = (py - T.gy0) * T.inv_cell;
    const int ix0 = (int)floorf(gx), iy0 = (int)floorf(gy);
    if (ix0 < 0 || iy0 < 0 || ix0 >= gnx || iy0 >= gny) {              // origin outside the grid: scan every wall
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            ray_dir(ca, sa, rs.kc[j], rs.ks[j], &dx, &dy);
            float best = NCG_RAY_LEN;
            for (int wi = 0; wi < T.n_walls; ++wi) { F4 a, b; M.wall((uint32_t)wi, &a, &b); best = fminf(best, ray_box_slab(a, b, px, py, dx, dy)); ++nt; }
            dst[rs.idx[j]] = sensor_obs_m(best);
        }
        *tests += nt;
        return;
    }
    const float fx = gx - (float)ix0, fy = gy - (float)iy0;
    // Amanatides-Woo in metres along the unit direction
    float tdx = dx != 0.0f ? T.cell * rcp_fast(fabsf(dx)) : INFINITY, tdy = dy != 0.0f ? T.cell * rcp_fast(fabsf(dy)) : INFINITY;
    float tmx = dx != 0.0f ? (dx > 0.0f ? 1.0f - fx : fx) * tdx : INFINITY;
    float tmy = dy != 0.0f ? (dy > 0.0f ? 1.0f - fy : fy) * tdy : INFINITY;
    int sx = dx > 0.0f ? 1 : -1, sy = dy > 0.0f ? 1 : -1;
    int ix = ix0, iy = iy0;
    const uint32_t h0 = M.cell(iy0 * gnx + ix0);
    const int k0 = (int)(h0 & 0xFFFFu), e0 = k0 + (int)(h0 >> 16);   // block range of the origin cell
    int k = k0, e = e0, j = 0;
    // the wall indices of block k are fetched one block ahead (while the previous block's walls are loaded and tested),
    // which takes one shared-memory latency out of every iteration's dependent chain: small batches are bound by that
    // chain, not by issue slots.  (At the end of a cell's list the fetch reads the next list's first block: unused.)
    uint32_t blo, bhi; M.block(k0, &blo, &bhi);
    const uint32_t blo0 = blo, bhi0 = bhi;
    float best = NCG_RAY_LEN;
    for (;;) {
        if (k >= e) {                                                   // this cell's list is done: leave or finish
            const float texit = fminf(tmx, tmy);
            bool fin = best <= texit || texit >= NCG_RAY_LEN;
            if (!fin) {
                if (tmx < tmy) { ix += sx; tmx += tdx; fin = (unsigned)ix >= (unsigned)gnx; }
                else { iy += sy; tmy += tdy; fin = (unsigned)iy >= (unsigned)gny; }
                if (!fin) { const uint32_t h = M.cell(iy * gnx + ix); k = (int)(h & 0xFFFFu); e = k + (int)(h >> 16); M.block(k, &blo, &bhi); }
            }
            if (fin) {
                // j is a run-time value here: pick this ray's slot and the next ray's constants with selects, not by
                // indexing the (register-resident) arrays
                int slot = rs.idx[0]; float kcn = 0.0f, ksn = 0.0f;
#pragma unroll
                for (int m = 1; m < RPL; ++m) { if (j == m) slot = rs.idx[m]; if (j + 1 == m) { kcn = rs.kc[m]; ksn = rs.ks[m]; } }
                dst[slot] = sensor_obs_m(best);
                if (++j == RPL) break;
                // next ray of this lane, same origin cell
                ray_dir(ca, sa, kcn, ksn, &dx, &dy);
                tdx = dx != 0.0f ? T.cell * rcp_fast(fabsf(dx)) : INFINITY; tdy = dy != 0.0f ? T.cell * rcp_fast(fabsf(dy)) : INFINITY;
                sx = dx > 0.0f ? 1 : -1; sy = dy > 0.0f ? 1 : -1;
                tmx = dx != 0.0f ? (dx > 0.0f ? 1.0f - fx : fx) * tdx : INFINITY;
                tmy = dy != 0.0f ? (dy > 0.0f ? 1.0f - fy : fy) * tdy : INFINITY;
                ix = ix0; iy = iy0; k = k0; e = e0; best = NCG_RAY_LEN; blo = blo0; bhi = bhi0;
            }
        }
        if (k < e) {                                                    // one block: four walls (padding repeats wall 0, masked)
            const uint32_t lo = blo, hi = bhi; ++k; M.block(k, &blo, &bhi);
            F4 a0, b0, a1, b1, a2, b2, a3, b3;                           // a short block is padded by repeating its first wall
            M.wall(lo & 0xFFFFu, &a0, &b0); M.wall(lo >> 16, &a1, &b1); M.wall(hi & 0xFFFFu, &a2, &b2); M.wall(hi >> 16, &a3, &b3);
            const float t0 = ray_box_slab(a0, b0, px, py, dx, dy), t1 = ray_box_slab(a1, b1, px, py, dx, dy);
            const float t2 = ray_box_slab(a2, b2, px, py, dx, dy), t3 = ray_box_slab(a3, b3, px, py, dx, dy);
            best = fminf(best, fminf(fminf(t0, t1), fminf(t2, t3)));
            nt += 4u;
        }
    }
    *tests += nt;
}
// convenience form: builds the lane's ray set (the kernel hoists that out of its step loop)
template <int RPL, bool SH>
NCG_HD void cast_rays(const Track& T, float px, float py, float angle, int q0, float* dst, unsigned* tests) {
    const RaySet<RPL> rs = ray_set<RPL>(q0);
    cast_rays<RPL, SH>(T, px, py, angle, rs, dst, tests);
}
// ---- the same rays, handed out from a queue (what the step kernel runs once a batch fills every SM with resident CTAs).
// A fixed lane -> rays assignment makes a warp wait for its busiest lane: the rays along the track walk ~8x more blocks
// than the ones across it, and lane loads differ by ~2.7x.  Here every ray of a CTA's cars is one job; jobs are ordered
// longest-expected first (ray 0 and 8 -- ahead and behind -- of every car, then their neighbours, the across-track
// rays last) and each lane takes the next job when its ray ends, so the warps of a CTA drain together.  Which lane casts
// a ray does not change the ray's result.  (Ordering the jobs by what each ray cost on the previous step -- lists
// refilled through two more shared-memory atomics per job -- models 25 % fewer loop iterations and measured 14 % slower:
// the extra dependent latency sits in the divergent end-of-ray branch the whole warp waits for.)
// RayCar is what a job needs to know about its car.
struct RayCar { float px, py, ca, sa, fx, fy; int cell0; uint32_t h0; };      // 8 words; cell0 = ix0 | iy0 << 16, or -1 outside the grid
NCG_HD RayCar ray_car(const Track& T, float px, float py, float angle) {
    RayCar rc; rc.px = px; rc.py = py; sincos_heading(angle, &rc.sa, &rc.ca);
    const float gx = (px - T.gx0) * T.inv_cell, gy = (py - T.gy0) * T.inv_cell;
    const int ix0 = (int)floorf(gx), iy0 = (int)floorf(gy);
    const bool in = ix0 >= 0 && iy0 >= 0 && ix0 < T.gnx && iy0 < T.gny;
    rc.fx = gx - (float)ix0; rc.fy = gy - (float)iy0;
    rc.cell0 = in ? (ix0 | (iy0 << 16)) : -1;
    rc.h0 = in ? T.cells[iy0 * T.gnx + ix0] : 0u;
    return rc;
}
#define NCG_RAY_ORDER 0xC4B5D3A6E297F180ULL      // nibble r = the ray index of job class r: 0,8,1,15,7,9,2,14,6,10,3,13,5,11,4,12
// cars: n_cars RayCar rows; job j = (class j / n_cars, car j % n_cars), `magic` = ceil(2^17 / n_cars) turns the division
// into a multiply (exact for j < 1024, n_cars <= 64); car c's observation row is row c, or c + gap for c >= gap_at (a
// CTA with two physics warps keeps the second one's cars from slot 32); j0 = this lane's first job (>= 16 * n_cars:
// none, < 0: claim one); *ctr = the next unclaimed job.  obs22 = word 22 of row 0, rows obs_stride apart.  rot = the 16
// (cos, sin) ray rotations.
template <bool SH>
NCG_HD void cast_rays_queue(const Track& T, const float* cars, int n_cars, unsigned magic, int gap_at, int gap, int j0, int* ctr,
                            float* obs22, int obs_stride, const float* rot, unsigned* tests) {
    const RayMem<SH> M(T);
    const int gnx = T.gnx, gny = T.gny, total = 16 * n_cars;
    unsigned nt = 0;
    int jn = j0;                                    // the pre-assigned first job, then -1 = claim one from the counter
    float px = 0.0f, py = 0.0f, dx = 1.0f, dy = 0.0f, tdx = INFINITY, tdy = INFINITY, tmx = INFINITY, tmy = INFINITY, best = NCG_RAY_LEN;
    int sx = 1, sy = 1, ix = 0, iy = 0, k = 0, e = 0;
    float* out = nullptr;
    uint32_t blo = 0u, bhi = 0u;                    // the wall indices of block k, fetched one block ahead (see cast_rays)
    for (;;) {
        if (k >= e) {                                                   // this cell's list is done: leave or finish
            const float texit = fminf(tmx, tmy);
            bool fin = best <= texit || texit >= NCG_RAY_LEN;
            if (!fin) {
                if (tmx < tmy) { ix += sx; tmx += tdx; fin = (unsigned)ix >= (unsigned)gnx; }
                else { iy += sy; tmy += tdy; fin = (unsigned)iy >= (unsigned)gny; }
                if (!fin) { const uint32_t h = M.cell(iy * gnx + ix); k = (int)(h & 0xFFFFu); e = k + (int)(h >> 16); M.block(k, &blo, &bhi); }
            }
            if (fin) {
                if (out) *out = sensor_obs_m(best);
                int j = jn; jn = -1;
                if (j < 0) {
#if defined(__CUDA_ARCH__)
                    j = atomicAdd(ctr, 1);
#else
                    j = (*ctr)++;
#endif
                }
                if (j >= total) break;
                const unsigned r = ((unsigned)j * magic) >> 17;
                const int car = j - (int)r * n_cars, ray = (int)((NCG_RAY_ORDER >> (4u * r)) & 15ull);
                const F4 c0 = *reinterpret_cast<const F4*>(cars + car * 8), c1 = *reinterpret_cast<const F4*>(cars + car * 8 + 4);
                const float kc = rot[2 * ray], ks = rot[2 * ray + 1];
                px = c0.x; py = c0.y;
                ray_dir(c0.z, c0.w, kc, ks, &dx, &dy);
                out = obs22 + (car + (car >= gap_at ? gap : 0)) * obs_stride + ray;
                best = NCG_RAY_LEN;
                const int cell0 = (int)f2u(c1.z);
                if (cell0 < 0) {                                        // origin outside the grid: scan every wall
                    for (int wi = 0; wi < T.n_walls; ++wi) { F4 a, b; M.wall((uint32_t)wi, &a, &b); best = fminf(best, ray_box_slab(a, b, px, py, dx, dy)); ++nt; }
                    k = e = 0; tmx = tmy = INFINITY;                    // -> stored and replaced by the next job on the next pass
                    continue;
                }
                tdx = dx != 0.0f ? T.cell * rcp_fast(fabsf(dx)) : INFINITY; tdy = dy != 0.0f ? T.cell * rcp_fast(fabsf(dy)) : INFINITY;
                sx = dx > 0.0f ? 1 : -1; sy = dy > 0.0f ? 1 : -1;
                tmx = dx != 0.0f ? (dx > 0.0f ? 1.0f - c1.x : c1.x) * tdx : INFINITY;
                tmy = dy != 0.0f ? (dy > 0.0f ? 1.0f - c1.y : c1.y) * tdy : INFINITY;
                ix = cell0 & 0xFFFF; iy = cell0 >> 16;
                const uint32_t h0 = f2u(c1.w);
                k = (int)(h0 & 0xFFFFu); e = k + (int)(h0 >> 16);
                M.block(k, &blo, &bhi);
            }
        }
        if (k < e) {                                                    // one block: four walls (padding repeats wall 0, masked)
            const uint32_t lo = blo, hi = bhi; ++k; M.block(k, &blo, &bhi);
            F4 a0, b0, a1, b1, a2, b2, a3, b3;
            M.wall(lo & 0xFFFFu, &a0, &b0); M.wall(lo >> 16, &a1, &b1); M.wall(hi & 0xFFFFu, &a2, &b2); M.wall(hi >> 16, &a3, &b3);
            const float t0 = ray_box_slab(a0, b0, px, py, dx, dy), t1 = ray_box_slab(a1, b1, px, py, dx, dy);
            const float t2 = ray_box_slab(a2, b2, px, py, dx, dy), t3 = ray_box_slab(a3, b3, px, py, dx, dy);
            best = fminf(best, fminf(fminf(t0, t1), fminf(t2, t3)));
            nt += 4u;
        }
    }
    *tests += nt;
}
NCG_HD float sensor_obs(float dist) { return sensor_obs_m(dist); }

// ------------------------------------------------------------------ Philox4x32-10 (synthetic actions)
NCG_HD void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
NCG_HD float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// raw action -> [throttle, brake, steer] (base_env.py:201-252)
NCG_HD void action_continuous(float tb, float steer, float* thr, float* brk, float* st) {
    if (tb >= 0.0f) { *thr = tb; *brk = 0.0f; } else { *thr = 0.0f; *brk = -tb; }
    *st = steer;
}
NCG_HD void action_discrete(int a, float* thr, float* brk, float* st) {
    float tb = a == 1 ? 1.0f : (a == 2 ? -1.0f : 0.0f), s = a == 3 ? -1.0f : (a == 4 ? 1.0f : 0.0f);
    action_continuous(tb, s, thr, brk, st);
}
NCG_HD void action_synthetic(uint64_t seed, uint32_t car, uint32_t step, int mode, bool discrete, float* thr, float* brk, float* st) {
    uint32_t r[4]; philox4x32(car, step, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    if (discrete) { action_discrete((int)(r[0] % 5u), thr, brk, st); return; }
    float a0 = u01(r[0]), a1 = u01(r[1]);
    if (mode == 0) action_continuous(2.0f * a0 - 1.0f, 2.0f * a1 - 1.0f, thr, brk, st);
    else action_continuous(0.2f + 0.8f * a0, -0.2f + 0.8f * a1, thr, brk, st);
}

}  // namespace ncg
