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
        reward[i] = car_step(R, T, act3[i * 3], act3[i * 3 + 1], act3[i * 3 + 2], contacts != 0, obs + i * 38, &xf[i], &cnt);
        unsigned tests = 0;
        for (int k = 0; k < 16; ++k) obs[i * 38 + 22 + k] = sensor_obs(cast_ray(T, R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], k, &tests));
        cnt.ray_tests += tests;
    }
    bool te, tr; int why;
    env_decide(xf, C, reset_on_lap != 0, f2u(records[NCG_R_STEP]), &te, &tr, &why);
    for (int i = 0; i < C; ++i) car_finish(records + i * NCG_RECORD_WORDS, reward[i]);
    *terminated = te; *truncated = tr; *reason = why;
    if (counters) { counters[0] = cnt.ray_tests; counters[1] = cnt.contact_steps; counters[2] = cnt.toi_events; counters[3] = cnt.overflow; counters[4] = cnt.laps; }
}

void hc_env_reset(const float* blob, float* records, int C, int fresh, int track_id, float* obs) {
    Track T = track_view(blob, blob);
    for (int i = 0; i < C; ++i) {
        float* R = records + i * NCG_RECORD_WORDS;
        reset_record(R, T, fresh != 0, (uint32_t)track_id);
        if (obs) {
            observe_state(R, obs + i * 38);
            unsigned tests = 0;
            for (int k = 0; k < 16; ++k) obs[i * 38 + 22 + k] = sensor_obs(cast_ray(T, R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], k, &tests));
        }
    }
}

// brute-force sensor (all walls) for checking the grid traversal
void hc_sensors_brute(const float* blob, float x, float y, float angle, float* out16) {
    Track T = track_view(blob, blob);
    for (int i = 0; i < 16; ++i) {
        Track B = T; B.gnx = 0; B.gny = 0;   // an empty grid sends cast_ray down its all-walls path
        unsigned tests = 0;
        out16[i] = cast_ray(B, x, y, angle, i, &tests);
    }
}
void hc_sensors_grid(const float* blob, float x, float y, float angle, float* out16, unsigned* tests) {
    Track T = track_view(blob, blob);
    for (int i = 0; i < 16; ++i) out16[i] = cast_ray(T, x, y, angle, i, tests);
}
// the generic b2PolygonShape::RayCast loop, kept here to prove the branch-free box version bit-identical
static float ray_box_generic(const float* w, V2 P1, V2 P2, float maxFraction) {
    Rot q; q.c = w[2]; q.s = w[3]; V2 pos = mk(w[0], w[1]); Box b; b.hx = w[4]; b.hy = w[5];
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
// returns the number of walls on which the two implementations disagree (bitwise) for one ray
int hc_ray_box_mismatches(const float* blob, float x1, float y1, float x2, float y2, float maxFraction) {
    Track T = track_view(blob, blob); int bad = 0;
    for (int wi = 0; wi < T.n_walls; ++wi) {
        float a = ray_box_fraction(T.walls + wi * WALL_STRIDE, mk(x1, y1), mk(x2, y2), maxFraction);
        float b = ray_box_generic(T.walls + wi * WALL_STRIDE, mk(x1, y1), mk(x2, y2), maxFraction);
        if (f2u(a) != f2u(b)) ++bad;
    }
    return bad;
}
int hc_on_track(const float* blob, float x, float y) { Track T = track_view(blob, blob); return on_track(T, x, y) ? 1 : 0; }
void hc_synthetic_action(unsigned long long seed, unsigned car, unsigned step, int mode, int discrete, float* out3) {
    action_synthetic(seed, car, step, mode, discrete != 0, out3, out3 + 1, out3 + 2);
}
}
