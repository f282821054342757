// TEST INFRASTRUCTURE ONLY.  Compiles the product's per-car device code (ncg_car.cuh) for the host so the
// scalar phases can be compared with the oracle on a machine without a GPU.  This library is never shipped,
// never loaded by nascargymnasium_b200 and is not a fallback: the product path is the CUDA build only.
#include "../../nascargymnasium_b200/csrc/ncg_car.cuh"

using namespace ncg;

extern "C" {

// One env of C cars: records [C][128], actions [C][3] (throttle, brake, steer), obs [C][38], reward [C].
void hc_env_step(const float* blob, float* records, int C, const float* act3, int contacts, int reset_on_lap, float* obs,
                 float* reward, int* terminated, int* truncated, int* reason, unsigned long long* counters) {
    Track T = track_view(blob, blob);
    Counters cnt = {0, 0, 0, 0, 0};
    uint32_t xf[NCG_MAX_CARS];
    for (int i = 0; i < C; ++i) {
        float* R = records + i * NCG_RECORD_WORDS;
        reward[i] = car_step(R, T, act3[i * 3], act3[i * 3 + 1], act3[i * 3 + 2], contacts, obs + i * 38, &xf[i], &cnt);
        unsigned tests = 0;
        for (int k = 0; k < 16; ++k) cast_rays<1, false>(T, R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], k, obs + i * 38 + 22, &tests);
        cnt.ray_tests += tests;
    }
    bool te, tr; int why;
    env_decide(xf, C, reset_on_lap != 0, f2u(records[NCG_R_STEP]), &te, &tr, &why);
    for (int i = 0; i < C; ++i) car_finish(records + i * NCG_RECORD_WORDS, reward[i]);
    *terminated = te; *truncated = tr; *reason = why;
    if (counters) { counters[0] = cnt.ray_tests; counters[1] = cnt.contact_steps; counters[2] = cnt.toi_events; counters[3] = cnt.overflow; counters[4] = cnt.laps; }
}

// The same env step in the optional shared-world mode (the cars of the env collide): what the kernel's physics warp does
// with its lanes, done here one car after the other.  pairs: the env's pair table [NCG_CC_STRIDE] floats.
void hc_env_step_cc(const float* blob, float* records, int C, const float* act3, int contacts, int reset_on_lap, float* pairs,
                    float* obs, float* reward, int* terminated, int* truncated, int* reason, unsigned long long* counters) {
    Track T = track_view(blob, blob);
    Counters cnt = {0, 0, 0, 0, 0};
    uint32_t xf[NCG_MAX_CARS];
    static World Ws[NCG_MAX_CARS];
    Body W[NCG_MAX_CARS]; DynPre pre[NCG_MAX_CARS]; StepCtx ctx[NCG_MAX_CARS];
    const bool joint = f2u(pairs[NCG_CC_COUNT]) != 0u;
    for (int i = 0; i < C; ++i) car_dyn_pre(records + i * NCG_RECORD_WORDS, T, act3[i * 3], act3[i * 3 + 1], act3[i * 3 + 2], &W[i], &pre[i]);
    if (!joint) { for (int i = 0; i < C; ++i) body_step(W[i], records + i * NCG_RECORD_WORDS, T, NCG_DT, contacts, &cnt); }
    else {
        for (int i = 0; i < C; ++i) { Ws[i].b = W[i]; Ws[i].v230 = contacts == 2; w_load_contacts(Ws[i], records + i * NCG_RECORD_WORDS); }
        shared_world_step(Ws, C, pairs, T, NCG_DT, &cnt);
        for (int i = 0; i < C; ++i) { w_store_contacts(Ws[i], records + i * NCG_RECORD_WORDS); W[i] = Ws[i].b; W[i].inv_dt0 = 1.0f / NCG_DT; W[i].force = mk(0.0f, 0.0f); W[i].torque = 0.0f; }
    }
    for (int i = 0; i < C; ++i) car_dyn_post(records + i * NCG_RECORD_WORDS, W[i], pre[i], &ctx[i], &cnt);
    if (!joint) {
        AABB fat[NCG_MAX_CARS];
        for (int c = 0; c < C; ++c) { const float* Rc = records + c * NCG_RECORD_WORDS; fat[c].lx = Rc[NCG_R_FAT_LX]; fat[c].ly = Rc[NCG_R_FAT_LY]; fat[c].ux = Rc[NCG_R_FAT_UX]; fat[c].uy = Rc[NCG_R_FAT_UY]; }
        cc_find_new_pairs(fat, C, pairs);
    }
    for (int i = 0; i < C; ++i) {
        float* R = records + i * NCG_RECORD_WORDS;
        reward[i] = car_step_rules<false>(R, T, &ctx[i], obs + i * 38, &xf[i], &cnt);
        unsigned tests = 0;
        for (int k = 0; k < 16; ++k) cast_rays<1, false>(T, R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], k, obs + i * 38 + 22, &tests);
        cnt.ray_tests += tests;
    }
    bool te, tr; int why;
    env_decide(xf, C, reset_on_lap != 0, f2u(records[NCG_R_STEP]), &te, &tr, &why);
    for (int i = 0; i < C; ++i) car_finish(records + i * NCG_RECORD_WORDS, reward[i]);
    *terminated = te; *truncated = tr; *reason = why;
    if (counters) { counters[0] = cnt.ray_tests; counters[1] = cnt.contact_steps; counters[2] = cnt.toi_events; counters[3] = cnt.overflow; counters[4] = cnt.laps; }
}
void hc_env_reset_cc(const float* blob, float* records, int C, int fresh, int track_id, float gdx, float gdy, float* pairs, float* obs) {
    Track T = track_view(blob, blob);
    StartPose sp0; sp0.x = 0.0f; sp0.y = 0.0f; sp0.a = 0.0f;
    for (int i = 0; i < C; ++i) {
        float* R = records + i * NCG_RECORD_WORDS;
        reset_record(R, T, fresh != 0, (uint32_t)track_id, cc_start_pose(sp0, i, gdx, gdy));
        if (obs) {
            observe_state(R, obs + i * 38);
            unsigned tests = 0;
            for (int k = 0; k < 16; ++k) cast_rays<1, false>(T, R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], k, obs + i * 38 + 22, &tests);
        }
    }
    if (fresh) for (int w = 0; w < NCG_CC_STRIDE; ++w) pairs[w] = 0.0f;
}
int hc_cc_stride() { return NCG_CC_STRIDE; }

void hc_env_reset(const float* blob, float* records, int C, int fresh, int track_id, float* obs) {
    Track T = track_view(blob, blob);
    for (int i = 0; i < C; ++i) {
        float* R = records + i * NCG_RECORD_WORDS;
        { StartPose sp; sp.x = 0.0f; sp.y = 0.0f; sp.a = 0.0f; reset_record(R, T, fresh != 0, (uint32_t)track_id, sp); }
        if (obs) {
            observe_state(R, obs + i * 38);
            unsigned tests = 0;
            for (int k = 0; k < 16; ++k) cast_rays<1, false>(T, R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], k, obs + i * 38 + 22, &tests);
        }
    }
}

// b2PolygonShape::RayCast for a box, Box2D's own half-plane clipping loop (b2PolygonShape.cpp), float32: the
// arithmetic the reference's sensors go through.  Test-side reference for the product's slab test.
static float ray_box_generic(const float* w, V2 P1, V2 P2, float maxFraction) {
    Rot q; q.c = w[2]; q.s = w[3]; V2 pos = mk(w[0], w[1]); Box b; b.hx = w[7]; b.hy = w[5];      // the exact half-length (w[4] is the rays' own, 0.1 mm longer)
    V2 p1 = mulT(q, P1 - pos), p2 = mulT(q, P2 - pos), d = p2 - p1;
    float lower = 0.0f, upper = maxFraction; int index = -1;
    for (int i = 0; i < 4; ++i) {
        V2 n = box_n(i);
        float num = dot(n, box_v(b, i) - p1), den = dot(n, d);
        if (den == 0.0f) { if (num < 0.0f) return -1.0f; }
        else {
            if (den < 0.0f && num < lower * den) { lower = num / den; index = i; }
            else if (den > 0.0f && num < upper * den) { upper = num / den; }
        }
        if (upper < lower) return -1.0f;
    }
    return index >= 0 ? lower : -1.0f;
}
// all-walls scan with Box2D's arithmetic (metres), the way DistanceSensor sets the ray up (distance_sensor.py:95-113)
void hc_sensors_brute(const float* blob, float x, float y, float angle, float* out16) {
    Track T = track_view(blob, blob);
    for (int i = 0; i < 16; ++i) {
        double a = (double)angle - (double)i * (3.14159265358979323846 / 8.0);
        V2 P1 = mk(x, y), P2 = mk((float)((double)x + cos(a) * 250.0), (float)((double)y + sin(a) * 250.0));
        float best = 1.0f;
        for (int wi = 0; wi < T.n_walls; ++wi) { float fr = ray_box_generic(T.walls + wi * WALL_STRIDE, P1, P2, 1.0f); if (fr >= 0.0f && fr < best) best = fr; }
        out16[i] = best * 250.0f;
    }
}
// the product's grid traversal + slab test; out16 in metres
void hc_sensors_grid(const float* blob, float x, float y, float angle, float* out16, unsigned* tests) {
    Track T = track_view(blob, blob);
    float n[16];
    for (int i = 0; i < 16; ++i) cast_rays<1, false>(T, x, y, angle, i, n, tests);
    for (int i = 0; i < 16; ++i) out16[i] = n[i] * 250.0f;
}
// the 2- and 4-rays-per-lane variants the kernel uses must give the same 16 numbers as the 1-ray variant
int hc_sensors_multi_mismatches(const float* blob, float x, float y, float angle) {
    Track T = track_view(blob, blob);
    float a[16], b[16], c[16]; unsigned tests = 0; int bad = 0;
    for (int i = 0; i < 16; ++i) cast_rays<1, false>(T, x, y, angle, i, a, &tests);
    for (int q = 0; q < 8; ++q) cast_rays<2, false>(T, x, y, angle, q < 4 ? q : q + 4, b, &tests);
    for (int q = 0; q < 4; ++q) cast_rays<4, false>(T, x, y, angle, q, c, &tests);
    for (int i = 0; i < 16; ++i) if (f2u(a[i]) != f2u(b[i]) || f2u(a[i]) != f2u(c[i])) ++bad;
    // the queued form (three cars so the job -> (class, car) split is exercised; the middle one is the pose under test)
    float cars[3 * 8], rot[32], obs[3 * 41]; int ctr = 0;
    for (int i = 0; i < 16; ++i) ray_rotation(i, &rot[2 * i], &rot[2 * i + 1]);
    for (int cidx = 0; cidx < 3; ++cidx) {
        const RayCar rc = ray_car(T, x + 3.0f * (float)(cidx - 1), y - 2.0f * (float)(cidx - 1), angle + 0.4f * (float)(cidx - 1));
        float* d = cars + cidx * 8;
        d[0] = rc.px; d[1] = rc.py; d[2] = rc.ca; d[3] = rc.sa; d[4] = rc.fx; d[5] = rc.fy; d[6] = u2f((uint32_t)rc.cell0); d[7] = u2f(rc.h0);
    }
    for (int i = 0; i < 3 * 41; ++i) obs[i] = -1.0f;
    cast_rays_queue<false>(T, cars, 3, (131072u + 2u) / 3u, 3, 0, -1, &ctr, obs + 22, 41, rot, &tests);
    for (int i = 0; i < 16; ++i) if (f2u(a[i]) != f2u(obs[41 + 22 + i])) ++bad;
    for (int cidx = 0; cidx < 3; ++cidx) for (int i = 0; i < 16; ++i) if (!(obs[cidx * 41 + 22 + i] >= 0.0f)) ++bad;      // every job ran
    if (ctr != 49) ++bad;                                  // 48 jobs claimed, one claim past the end
    return bad;
}
// chord-nearest segment with the per-cell candidate mask (masked = 1) or over all segments (masked = 0: a view with an
// empty grid sends nearest_segment down its all-segments path)
void hc_nearest_segment(const float* blob, float x, float y, int masked, float* out2) {
    Track T = track_view(blob, blob);
    if (!masked) { T.gnx = 0; T.gny = 0; }
    nearest_segment(T, x, y, out2, out2 + 1);
}
void hc_sincos_heading(float a, float* out2) { sincos_heading(a, out2, out2 + 1); }
int hc_on_track(const float* blob, float x, float y) { Track T = track_view(blob, blob); return on_track(T, x, y) ? 1 : 0; }
// sweep_face_bound (the exact TOI early-out's lower bound of the car-wall distance over a whole sweep); in: c0x c0y a0 c1x c1y a1,
// wall: px py angle hx hy
float hc_sweep_face_bound(const float* sw6, const float* wall5) {
    Sweep s; s.c0 = mk(sw6[0], sw6[1]); s.a0 = sw6[2]; s.c = mk(sw6[3], sw6[4]); s.a = sw6[5]; s.alpha0 = 0.0f;
    Xf xfB; xfB.p = mk(wall5[0], wall5[1]); xfB.q = rot(wall5[2]);
    Box bB; bB.hx = wall5[3]; bB.hy = wall5[4];
    return sweep_face_bound(s, rot(s.a), xfB, bB, 1e30f);
}
void hc_toi_counters(unsigned long long* out3) { out3[0] = g_toi_full; out3[1] = g_toi_skip_reach; out3[2] = g_toi_skip_face; }
void hc_synthetic_action(unsigned long long seed, unsigned car, unsigned step, int mode, int discrete, float* out3) {
    action_synthetic(seed, car, step, mode, discrete != 0, out3, out3 + 1, out3 + 2);
}
}
