// ncg_b200.cu -- kernels and the C ABI (include/ncg_b200.h) of the batched CarEnv stepping engine.
//
// Work decomposition (DESIGN.md "Kernels"): a CTA owns up to 32 car slots (whole envs) and is warp-specialised.
//   warp 0  ("physics warp")  one car per lane: action -> forces -> tyres -> b2World.Step -> lap timer -> disable
//                             rules -> obs[0..21] -> reward -> env termination -> same-step auto-reset.
//   warps 1..RW ("ray warps") the CTA's 32 x 16 sensor rays, RPL rays per lane, then the coalesced store of the
//                             finished 38-float observation rows to HBM.
// The two halves are a producer/consumer pair over a double-buffered pose + observation block in shared memory,
// handed over with named barriers (bar.sync / bar.arrive), so in a multi-step rollout the physics of step t+1
// overlaps the rays of step t.  Records stay in shared memory for the whole launch (row stride 129 words: lane l
// reading word k of its own record hits bank (l+k) mod 32, conflict-free) and move to/from HBM as coalesced
// 16-byte accesses; the CTA's track table is staged into shared memory once per launch with a TMA bulk copy
// (cp.async.bulk + mbarrier).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "ncg_car.cuh"

using namespace ncg;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(NCG_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

struct DevStats { unsigned long long car_steps, episodes, laps, ray_tests, contact_steps, toi_events, overflow; double return_sum; };

struct KParams {
    float* records; const float* blob; const long long* track_off;
    const int2* cta_tab;                                   // per group: {first slot, number of envs}; all of one track, <= 32 cars
    const int* slot_env;                                   // slot -> env, envs ordered by track; NULL = identity (the map is already sorted)
    const int2* pair_tab;                                  // two-physics-warp shape: per CTA the two groups it serves {g0, g1 or -1}
    const float* reset_obs;                                // [n_tracks][NCG_OBS_DIM]: the observation every reset_car yields on a track
    int E, C, discrete, reset_on_lap, auto_reset, contacts, stage, track_info, debug_skip, queue;
    const void* actions; float* obs; float* reward; uint8_t* term; uint8_t* trunc; float* final_obs;
    int T; unsigned long long seed; int mode; unsigned step_base; unsigned car_base;
    float* obs_roll; float* rew_roll; uint8_t* done_roll;
    StartPose start;                                       // CarEnv(start_position, start_angle)
    int car_contacts; float grid_dx, grid_dy;              // optional shared world: the cars of an env collide (default off)
    float* cc_pairs; World* cc_worlds;                     //   [E][NCG_CC_STRIDE] pair tables; [N] per-car Worlds (scratch of the joint step)
    float2* vel_hist;                                      // optional [N][NCG_VEL_HISTORY]: Car.velocity_history ring (info only)
    float* ep_return; int* ep_length; int* any_done;      // optional: episode return per car / length per env of finished envs
    int redraw, n_tracks; unsigned redraw_step; unsigned long long redraw_seed;   // track_file=None: a finished env re-draws its track
    int* env_track;                                        // [E] the env -> track map as the device sees it (redraw writes it)
    int* redrawn;                                          // mapped host word: set when this launch moved an env
    DevStats* stats;
};

#define CPB 32                    /* car slots per CTA = lanes of the physics warp */
#define REC_STRIDE 129            /* shared-memory row stride of a record (odd: conflict-free column access) */
#define OBS_STRIDE 41             /* shared-memory row stride of an observation row */
/* named barriers, NB step buffers: 1..NB the new poses of buffer b are published (rays can start); NB+1..2NB buffer b
   complete (obs[0..21], flags, reset poses); 2NB+1..3NB buffer b drained by the ray warps */

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS).  Issue and wait are split
// so the record load overlaps the copy.
__device__ __forceinline__ void tma_issue(float* dst, const float* src, unsigned bytes, unsigned long long* mbar) {
    unsigned mb = smem_u32(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(mb) : "memory");
}
__device__ __forceinline__ void tma_wait(unsigned long long* mbar) {
    unsigned mb = smem_u32(mbar), ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(mb) : "memory");
    }
}

// same-track reset of one record (kept out of line: it runs once per episode).  The observation after a reset_car is
// the same for every car of a track -- start pose, zero velocity, fresh tyres, the 16 rays of the start pose -- so it is
// computed once per track (ncg_reset_obs_kernel) instead of once per reset.
__device__ __noinline__ void reset_in_place(float* R, const Track T, const StartPose sp) {
    reset_record(R, T, false, f2u(R[NCG_R_TRACK]), sp);
}
// the start pose of car k of an env: the env's start pose, or its slot on the start grid of the shared world
__device__ __forceinline__ StartPose start_of(const StartPose sp, int k, int car_contacts, float gdx, float gdy) {
    return car_contacts ? cc_start_pose(sp, k, gdx, gdy) : sp;
}

// CarEnv.reset() in random-track mode (car_env.py:264-303): the finished env moves to another track, drawn uniformly among
// the others, and gets brand-new physics worlds (fresh reset) there.  The other track's table is read from global memory:
// this runs once per episode.  The CTA that made the move does not step the env again (single-step launches only); the host
// regroups the envs by track before the next launch.
__device__ __noinline__ uint32_t redraw_track(uint32_t cur, int n_tracks, unsigned env, unsigned step, unsigned long long seed) {
    if (n_tracks < 2) return cur;
    uint32_t r[4]; philox4x32(env, step, 0x7472636bu, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const uint32_t k = r[0] % (uint32_t)(n_tracks - 1);
    return k >= cur ? k + 1u : k;
}
__device__ __noinline__ void reset_on_track(float* R, const float* blob, const long long* track_off, uint32_t tid, const StartPose sp) {
    const float* g = blob + track_off[tid];
    const Track T = track_view(g, g);
    reset_record(R, T, true, tid, sp);
}

struct SmemLayout {
    int rec, obs, pose, flag, xf, gcar, act, otab, ray, rot, ctr, track, total;      // word offsets
};
// cpb = car slots of the CTA: 32 per physics warp
__host__ __device__ inline SmemLayout smem_layout(unsigned stage_words, int cpb, int nb) {
    SmemLayout L; int o = 0;
    L.rec = o; o += cpb * REC_STRIDE;
    L.obs = o; o += nb * cpb * OBS_STRIDE;         // [nb][cpb][OBS_STRIDE]: the step's observation rows
    o = (o + 3) & ~3;
    L.pose = o; o += nb * cpb * 4;                 // [nb][cpb] float4 {x, y, angle, -}
    L.flag = o; o += nb * cpb;                     // [nb][cpb] u32: bit0 terminated, bit1 truncated
    L.xf = o; o += cpb;
    L.gcar = o; o += cpb;                          // [cpb] global car index of a car slot
    o = (o + 3) & ~3;
    L.act = o; o += nb * cpb * 4;                  // [nb][cpb] float4 {throttle, brake, steer, -}: synthetic actions, made nb steps ahead
    L.otab = o; o += 2 * 40;                       // observation scale[38] (padded to 40) and lower clip bound[38]
    o = (o + 3) & ~3;
    L.ray = o; o += cpb * 8;                       // [cpb] RayCar: what a ray job needs to know about its car (ray queue)
    L.rot = o; o += 32;                            // the 16 ray rotations (cos, sin)
    L.ctr = o; o += 4;                             // [2] next unclaimed ray job
    L.track = o; o += (int)stage_words;            // 16-byte aligned for the TMA copy
    L.total = o;
    return L;
}

// MINB = CTAs per SM the register allocation must allow: 1 lets the physics warp keep its whole working set (track
// view, body, tyres) in registers -- right when the batch is at most one CTA per SM; 2 trades a few spills for
// twice the resident warps when there are waves of CTAs.
// PW = physics warps per CTA.  1: the CTA is one group of the CTA table (<= 32 car slots) with 16/RPL ray warps.
// 2: the CTA serves a pair of groups of one track -- two physics warps, 64 car slots, six ray warps that drain one ray
// queue, one staged track table -- which puts four physics warps on an SM (2 CTAs x 256 threads x 128 registers) where
// PW = 1 fits three (shared memory: three tables): at large batches a step is bound by the physics warps' dependent chains.
// 4: one group again, its car slots spread over FOUR physics warps of eight lanes ("spread" shape, single-car envs, batches of at
// most one CTA per SM): the four dependent chains run on the SM's four schedulers side by side, and a warp only walks the
// union of the code paths of its own eight cars -- what a contact-heavy step (a few thousand divergent instructions per car)
// is bound by; the contact-free chain is no shorter for it.
template <int RPL, int MINB, int PW>
__global__ void __launch_bounds__(PW == 2 ? 256 : 32 * (PW + 16 / RPL), MINB) ncg_step_kernel(KParams p) {
    constexpr int RW = PW == 2 ? 6 : 16 / RPL;      // ray warps
    constexpr int NT = 32 * (PW + RW);
    constexpr int GROUPS = PW == 2 ? 2 : 1;         // groups of the CTA table served by this CTA
    constexpr int SLOTS = 32 * GROUPS;              // car slots: 32 per group
    constexpr int LPW = PW == 4 ? 8 : 32;           // car slots per physics warp: warp w owns slots w*LPW .. w*LPW+LPW-1
    constexpr int LPC = 16 / RPL;                   // fixed ray mapping (PW == 1): lanes per car in a ray warp
    constexpr int CPW = 32 / LPC;                   //                              cars per ray warp
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long s_mbar;
    // step buffers between the physics and the ray warps: three when a CTA has an SM to itself (the physics warp then
    // never waits for a drained buffer; shared memory is not the limit there), two otherwise
    constexpr int NB = MINB == 1 ? 3 : 2;
    constexpr int BAR_POSE = 1, BAR_FULL = 1 + NB, BAR_EMPTY = 1 + 2 * NB;
    const SmemLayout L = smem_layout(0, SLOTS, NB);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* s_rec = smem + L.rec;
    float* s_obs = smem + L.obs;
    float4* s_pose = reinterpret_cast<float4*>(smem + L.pose);
    uint32_t* s_flag = reinterpret_cast<uint32_t*>(smem + L.flag);
    uint32_t* s_xf = reinterpret_cast<uint32_t*>(smem + L.xf);
    int* s_gcar = reinterpret_cast<int*>(smem + L.gcar);
    float4* s_act = reinterpret_cast<float4*>(smem + L.act);
    float* s_otab = smem + L.otab;
    float* s_ray = smem + L.ray;
    float* s_rot = smem + L.rot;
    int* s_ctr = reinterpret_cast<int*>(smem + L.ctr);
    float* s_track = smem + L.track;

    // the groups of envs this CTA serves: n0 cars from global car cb0 in slots 0.., n1 cars from cb1 in slots 32..
    int2 g0, g1 = make_int2(0, 0);
    if (GROUPS == 1) g0 = p.cta_tab[blockIdx.x];
    else { const int2 pr = p.pair_tab[blockIdx.x]; g0 = p.cta_tab[pr.x]; if (pr.y >= 0) g1 = p.cta_tab[pr.y]; }
    const int n0 = g0.y * p.C, n1 = g1.y * p.C, n_all = n0 + n1;
    const int N = p.E * p.C;
#define SLOT_OF(ci) ((ci) < n0 ? (ci) : 32 + (ci) - n0)           /* dense car index of the CTA -> slot */
#define GCAR_OF(ci) (s_gcar[SLOT_OF(ci)])                          /*                          -> global car */
    // a group is g.y consecutive entries of the slot list (envs ordered by track); without a list slot s is env s
    for (int ci = threadIdx.x; ci < n_all; ci += NT) {
        const int k = ci < n0 ? ci : ci - n0, le = k / p.C, sl = (ci < n0 ? g0.x : g1.x) + le;
        NCG_CHECK(sl >= 0 && sl < p.E && SLOT_OF(ci) < SLOTS, "slot list index / car slot");
        s_gcar[SLOT_OF(ci)] = (p.slot_env ? p.slot_env[sl] : sl) * p.C + (k - le * p.C);
        NCG_CHECK(s_gcar[SLOT_OF(ci)] >= 0 && s_gcar[SLOT_OF(ci)] < N, "global car index");
    }
    __syncthreads();
    const int cb0 = s_gcar[0];                                       // the CTA's first car: its record names the CTA's track

    // ---- track table: staged by TMA when the whole CTA shares a track, else read through L1/L2
    const float* staged = nullptr;
    if (p.stage) {
        if (threadIdx.x == 0) {
            const uint32_t tid = f2u(p.records[(size_t)cb0 * NCG_RECORD_WORDS + NCG_R_TRACK]);
            const float* g = p.blob + p.track_off[tid];
            tma_issue(s_track, g, f2u(__ldg(g + TH_STAGE_WORDS)) * 4u, &s_mbar);
        }
        staged = s_track;
    }
    // ---- records HBM -> shared (coalesced float4 reads, scalar shared stores into the padded rows)
    for (int i = threadIdx.x; i < n_all * (NCG_RECORD_WORDS / 4); i += NT) {
        const int ci = i >> 5;
        const float4 v = reinterpret_cast<const float4*>(p.records + (size_t)GCAR_OF(ci) * NCG_RECORD_WORDS)[i & 31];
        float* d = s_rec + SLOT_OF(ci) * REC_STRIDE + (i & 31) * 4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    if (threadIdx.x < NCG_OBS_DIM) { s_otab[threadIdx.x] = obs_scale(threadIdx.x); s_otab[40 + threadIdx.x] = obs_lo(threadIdx.x); }
    if (threadIdx.x < 16) ray_rotation((int)threadIdx.x, &s_rot[2 * threadIdx.x], &s_rot[2 * threadIdx.x + 1]);
    const bool synth = p.actions == nullptr;
    // the first GROUPS ray warps make the synthetic actions (one per group), two steps ahead of the physics warps
    const bool act_maker = synth && warp >= PW && warp < PW + GROUPS && lane < (warp == PW ? n0 : n1);
    const int act_slot = (warp - PW) * 32 + lane; const unsigned act_car = p.car_base + (unsigned)(act_maker ? s_gcar[act_slot] : 0);
    if (act_maker) {
        for (int t = 0; t < NB && t < p.T; ++t) {
            float thr, brk, st;
            action_synthetic(p.seed, act_car, p.step_base + (unsigned)t, p.mode, p.discrete != 0, &thr, &brk, &st);
            s_act[t * SLOTS + act_slot] = make_float4(thr, brk, st, 0.0f);
        }
    }
    __syncthreads();
    if (p.stage) tma_wait(&s_mbar);

    // the car slot this thread serves: its lane (physics warps) or, with the fixed ray mapping, the car its rays belong to
    const int slot = warp < PW ? warp * LPW + lane : (warp - PW) * CPW + lane / LPC;
    const bool active = warp < PW ? (PW == 2 ? lane < (warp == 0 ? n0 : n1) : (lane < LPW && slot < n0)) : (GROUPS == 1 && slot < n0);
    // track views are fixed for the launch (auto-reset keeps an env on its track): build them once
    const uint32_t my_tid = f2u(s_rec[NCG_R_TRACK]);              // slot 0: a CTA serves one track
    NCG_CHECK(my_tid < (uint32_t)p.n_tracks, "track id of the CTA");
    NCG_CHECK(!active || warp >= PW || f2u(s_rec[slot * REC_STRIDE + NCG_R_TRACK]) == my_tid, "a CTA serves ONE track");
    const float* gblob = p.blob + p.track_off[my_tid];
    const Track T = track_view(staged ? staged : gblob, gblob);
    const bool do_reset = p.auto_reset != 0;
    unsigned long long ray_tests = 0;

    if (warp < PW) {
        // =============================================================== physics warps: one car per lane
        const int gc = active ? s_gcar[slot] : 0;                // this lane's global car; its env: gc / C
        const int ge = p.C == 1 ? gc : gc / p.C;
        float* R = s_rec + (active ? slot : 0) * REC_STRIDE;
        Counters cnt = {0, 0, 0, 0, 0};
        unsigned long long episodes = 0; double ret_sum = 0.0;
        for (int t = 0, b = 0; t < p.T; ++t, b = b + 1 == NB ? 0 : b + 1) {
            if (t >= NB) bar_sync(BAR_EMPTY + b, NT);           // the ray warps have drained buffer b (step t-NB)
            float* rew_out = p.rew_roll ? p.rew_roll + (size_t)t * N : p.reward;
            float rew = 0.0f;
            StepCtx ctx;
            if (active) {
                float thr, brk, st;
                if (!synth) {
                    if (p.discrete) action_discrete(((const int*)p.actions)[gc], &thr, &brk, &st);
                    else { float2 a = ((const float2*)p.actions)[gc]; action_continuous(a.x, a.y, &thr, &brk, &st); }
                } else { const float4 a = s_act[b * SLOTS + slot]; thr = a.x; brk = a.y; st = a.z; }
                // Car.velocity_history (car.py:384-386): the speed update_physics saw, i.e. before b2World.Step (info only)
                if (p.vel_hist) p.vel_hist[(size_t)gc * NCG_VEL_HISTORY + f2u(R[NCG_R_STEP]) % NCG_VEL_HISTORY] = make_float2(R[NCG_R_VX], R[NCG_R_VY]);
                if (!p.car_contacts) { if (!(p.debug_skip & 2)) car_step_dynamics(R, T, thr, brk, st, p.contacts, &ctx, &cnt); }
                else {
                    // shared world: an env that has car-car contacts is stepped by its first car's lane over the per-car
                    // Worlds (global scratch), the others hand their bodies over and take them back; envs without any
                    // keep the per-lane step.  Every lane of an env takes the same branch (the pair count is the env's).
                    const int k = gc - ge * p.C;                                   // car index inside the env
                    const unsigned envmask = ((1u << p.C) - 1u) << (lane - k);
                    float* PT = p.cc_pairs + (size_t)ge * NCG_CC_STRIDE;
                    const bool joint = f2u(PT[NCG_CC_COUNT]) != 0u;
                    Body W; DynPre pre;
                    car_dyn_pre(R, T, thr, brk, st, &W, &pre);
                    if (!joint) body_step(W, R, T, NCG_DT, p.contacts, &cnt);
                    else {
                        World* Wk = p.cc_worlds + gc;
                        Wk->b = W; Wk->v230 = p.contacts == 2;
                        w_load_contacts(*Wk, R);
                        __syncwarp(envmask);
                        if (k == 0) shared_world_step(p.cc_worlds + (size_t)ge * p.C, p.C, PT, T, NCG_DT, &cnt);
                        __syncwarp(envmask);
                        w_store_contacts(*Wk, R);
                        W = Wk->b; W.inv_dt0 = 1.0f / NCG_DT; W.force = mk(0.0f, 0.0f); W.torque = 0.0f;
                    }
                    car_dyn_post(R, W, pre, &ctx, &cnt);
                    __syncwarp(envmask);
                    if (!joint && k == 0) {                                        // FindNewContacts among the env's cars
                        AABB fat[NCG_MAX_CARS];
                        for (int c = 0; c < p.C; ++c) { const float* Rc = R + c * REC_STRIDE; fat[c].lx = Rc[NCG_R_FAT_LX]; fat[c].ly = Rc[NCG_R_FAT_LY]; fat[c].ux = Rc[NCG_R_FAT_UX]; fat[c].uy = Rc[NCG_R_FAT_UY]; }
                        cc_find_new_pairs(fat, p.C, PT);
                    }
                }
                s_pose[b * SLOTS + slot] = make_float4(R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], 0.0f);
            }
            if (threadIdx.x == 0) s_ctr[b] = 32 * RW;           // ray queue: every ray lane starts on job = its index
            // the pose exists: let the ray warps start while this warp does the rest of the step
            __syncwarp();
            __threadfence_block();
            bar_arrive(BAR_POSE + b, NT);
            // single-car envs (C == 1, uniform) decide from the car's own result word: no exchange through shared memory,
            // no division by C, and the multi-car loops of env_decide unroll away
            const bool solo = p.C == 1;
            uint32_t xf = 0;
            if (active) {
                if (!(p.debug_skip & 2)) rew = car_step_rules<true>(R, T, &ctx, s_obs + (b * SLOTS + slot) * OBS_STRIDE, &xf, &cnt);
                if (!solo) s_xf[slot] = xf;
                if (p.track_info) {
                    uint32_t fl = f2u(R[NCG_R_FLAGS]) & ~(uint32_t)NCG_F_ON_TRACK;
                    if (on_track(T, R[NCG_R_X], R[NCG_R_Y])) fl |= NCG_F_ON_TRACK;
                    R[NCG_R_FLAGS] = u2f(fl);
                }
            }
            if (!solo) __syncwarp();
            // ---- env phase (every car of an env computes the same decision from the env's xf words)
            if (active) {
                const int le = solo ? lane : lane / p.C;
                bool te, tr; int why;
                if (solo) env_decide(&xf, 1, p.reset_on_lap != 0, f2u(R[NCG_R_STEP]), &te, &tr, &why);
                else env_decide(s_xf + warp * LPW + le * p.C, p.C, p.reset_on_lap != 0, f2u(R[NCG_R_STEP]), &te, &tr, &why);
                car_finish(R, rew);
                if (rew_out) rew_out[gc] = rew;
                const bool done = te || tr;
                if (solo || lane == le * p.C) {
                    if (p.done_roll) p.done_roll[(size_t)t * p.E + ge] = (uint8_t)((te ? 1 : 0) | (tr ? 2 : 0));
                    else { if (p.term) p.term[ge] = te ? 1 : 0; if (p.trunc) p.trunc[ge] = tr ? 1 : 0; }
                    if (done) ++episodes;
                }
                if (__builtin_expect(done, 0)) {
                    ret_sum += (double)R[NCG_R_CUM_REWARD];
                    if (p.ep_return) p.ep_return[gc] = R[NCG_R_CUM_REWARD];
                    if (solo || lane == le * p.C) { if (p.ep_length) p.ep_length[ge] = (int)f2u(R[NCG_R_STEP]); if (p.any_done) *p.any_done = 1; }
                }
                // ---- same-step auto-reset: CarPhysics.reset_car on the same track, or (random-track mode, single-step
                // launches) a fresh world on a newly drawn track; bits 8.. of the flag word name the track whose reset
                // observation the ray warps hand out
                uint32_t rtid = my_tid;
                if (__builtin_expect(done && do_reset, 0)) {
                    if (p.redraw) {
                        rtid = redraw_track(my_tid, p.n_tracks, (unsigned)ge, p.redraw_step, p.redraw_seed);
                        reset_on_track(R, p.blob, p.track_off, rtid, start_of(p.start, gc - ge * p.C, p.car_contacts, p.grid_dx, p.grid_dy));
                        if (solo || lane == le * p.C) {
                            p.env_track[ge] = (int)rtid; *p.redrawn = 1;
                            if (p.car_contacts) for (int w_ = 0; w_ < NCG_CC_STRIDE; ++w_) p.cc_pairs[(size_t)ge * NCG_CC_STRIDE + w_] = 0.0f;   // a new world
                        }
                    } else reset_in_place(R, T, start_of(p.start, gc - ge * p.C, p.car_contacts, p.grid_dx, p.grid_dy));
                }
                // (bits 8..: the row of reset_obs to hand out: one per track, or per (track, car of the env) on a start grid)
                s_flag[b * SLOTS + slot] = (te ? 1u : 0u) | (tr ? 2u : 0u) | ((p.car_contacts ? rtid * (uint32_t)p.C + (uint32_t)(gc - ge * p.C) : rtid) << 8);
            }
            __syncwarp();
            __threadfence_block();
            bar_arrive(BAR_FULL + b, NT);
        }
        // ---- counters
        unsigned long long v[7] = {active ? (unsigned long long)p.T : 0ull, episodes, cnt.laps, 0ull, cnt.contact_steps, cnt.toi_events, cnt.overflow};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            unsigned long long x = v[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0 && x) atomicAdd(((unsigned long long*)p.stats) + k, x);
        }
        for (int o = 16; o > 0; o >>= 1) ret_sum += __shfl_down_sync(0xffffffffu, ret_sum, o);
        if (lane == 0 && ret_sum != 0.0) atomicAdd(&p.stats->return_sum, ret_sum);
    } else {
        // =============================================================== ray warps
        const int rt = (warp - PW) * 32 + lane;                  // index among the CTA's ray lanes
        const int q = lane % LPC;
        const int q0 = RPL == 2 ? (q < 4 ? q : q + 4) : q;       // fixed mapping: a lane's rays are q0, q0+4, ... (90 deg apart)
        const int wslot0 = (warp - PW) * CPW;                    //                first car slot of this warp
        const RaySet<RPL> rs = ray_set<RPL>(q0);
        // which (car, word pair) this lane stores in each pass of the row write-out: fixed for the launch.  PW == 1: a warp
        // writes the rows of its own CPW cars; PW == 2: the dense cars of the CTA are dealt over all ray lanes.
        constexpr int NIT = GROUPS == 2 ? (SLOTS * (NCG_OBS_DIM / 2) + 32 * RW - 1) / (32 * RW) : (CPW * (NCG_OBS_DIM / 2) + 31) / 32;
        uint32_t pair_sk[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            if (GROUPS == 2) {
                const int m = rt + 32 * RW * it;
                pair_sk[it] = m < n_all * (NCG_OBS_DIM / 2) ? (uint32_t)(((m / (NCG_OBS_DIM / 2)) << 8) | ((m % (NCG_OBS_DIM / 2)) * 2)) : 0xFFFFFFFFu;
            } else {
                const int m = lane + 32 * it;
                pair_sk[it] = m < CPW * (NCG_OBS_DIM / 2) ? (uint32_t)(((m / (NCG_OBS_DIM / 2)) << 8) | ((m % (NCG_OBS_DIM / 2)) * 2)) : 0xFFFFFFFFu;
            }
        }
        unsigned tests = 0;
        const unsigned magic = (131072u + (unsigned)n_all - 1u) / (unsigned)n_all;
        for (int t = 0, b = 0; t < p.T; ++t, b = b + 1 == NB ? 0 : b + 1) {
            float* obs_out = p.obs_roll ? p.obs_roll + (size_t)t * N * NCG_OBS_DIM : p.obs;
            bar_sync(BAR_POSE + b, NT);
            if ((GROUPS == 2 || p.queue) && !(p.debug_skip & 1)) {
                // every ray warp derives the cars' ray origins itself (same values to the same words: no barrier between
                // the ray warps; nobody still reads last step's, every ray warp has passed that step's FULL barrier), then
                // all ray lanes of the CTA drain one queue of 16 x n_all rays
                for (int ci = lane; ci < n_all; ci += 32) {
                    const float4 ps = s_pose[b * SLOTS + SLOT_OF(ci)];
                    const RayCar c = ray_car(T, ps.x, ps.y, ps.z);
                    reinterpret_cast<float4*>(s_ray)[2 * ci] = make_float4(c.px, c.py, c.ca, c.sa);
                    reinterpret_cast<float4*>(s_ray)[2 * ci + 1] = make_float4(c.fx, c.fy, u2f((uint32_t)c.cell0), u2f(c.h0));
                }
                __syncwarp();
                float* o22 = s_obs + b * SLOTS * OBS_STRIDE + 22;
                if (staged) cast_rays_queue<true>(T, s_ray, n_all, magic, n0, 32 - n0, rt, s_ctr + b, o22, OBS_STRIDE, s_rot, &tests);
                else cast_rays_queue<false>(T, s_ray, n_all, magic, n0, 32 - n0, rt, s_ctr + b, o22, OBS_STRIDE, s_rot, &tests);
            } else if (GROUPS == 1 && active && !(p.debug_skip & 1)) {
                const float4 ps = s_pose[b * SLOTS + slot];
                float* dst = s_obs + (b * SLOTS + slot) * OBS_STRIDE + 22;
                if (staged) cast_rays<RPL, true>(T, ps.x, ps.y, ps.z, rs, dst, &tests);
                else cast_rays<RPL, false>(T, ps.x, ps.y, ps.z, rs, dst, &tests);
            }
            if (act_maker && t + NB < p.T) {                     // actions of step t+NB (this buffer's next use)
                float thr, brk, st;
                action_synthetic(p.seed, act_car, p.step_base + (unsigned)(t + NB), p.mode, p.discrete != 0, &thr, &brk, &st);
                s_act[b * SLOTS + act_slot] = make_float4(thr, brk, st, 0.0f);
            }
            bar_sync(BAR_FULL + b, NT);
            __syncwarp();
            // ---- observation rows shared -> HBM: 38 consecutive floats per car, written as float2 (a row is 19 float2,
            // so a pair never straddles two cars and every store is 8-byte aligned).  Words 0..21 arrive raw from the
            // physics warp and are scaled and clipped here; the ray words are already in [0,1] (scale 1, lower bound 0
            // leave them unchanged).
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int ci = (GROUPS == 2 ? 0 : wslot0) + (int)(pair_sk[it] >> 8), k = (int)(pair_sk[it] & 255u);
                if (pair_sk[it] != 0xFFFFFFFFu && ci < n_all) {
                    const int sl = SLOT_OF(ci);
                    const float* row = s_obs + (b * SLOTS + sl) * OBS_STRIDE + k;
                    float2 v;
                    v.x = obs_word(row[0], s_otab[k], s_otab[40 + k]);
                    v.y = obs_word(row[1], s_otab[k + 1], s_otab[41 + k]);
                    const size_t o = ((size_t)GCAR_OF(ci) * NCG_OBS_DIM + k) >> 1;
                    const uint32_t fw = s_flag[b * SLOTS + sl];
                    if (do_reset && (fw & 3u) != 0u) {
                        if (p.final_obs) reinterpret_cast<float2*>(p.final_obs)[o] = v;
                        if (obs_out) reinterpret_cast<float2*>(obs_out)[o] = __ldg(reinterpret_cast<const float2*>(p.reset_obs + (size_t)(fw >> 8) * NCG_OBS_DIM + k));   // the (new) track's reset observation
                    } else if (obs_out) reinterpret_cast<float2*>(obs_out)[o] = v;
                }
            }
            if (t + NB < p.T) { __threadfence_block(); bar_arrive(BAR_EMPTY + b, NT); }
        }
        ray_tests = tests;
        for (int o = 16; o > 0; o >>= 1) ray_tests += __shfl_down_sync(0xffffffffu, ray_tests, o);
        if (lane == 0 && ray_tests) atomicAdd(((unsigned long long*)p.stats) + 3, ray_tests);
    }
    // ---- records shared -> HBM
    __syncthreads();
    for (int i = threadIdx.x; i < n_all * (NCG_RECORD_WORDS / 4); i += NT) {
        const int ci = i >> 5;
        const float* d = s_rec + SLOT_OF(ci) * REC_STRIDE + (i & 31) * 4;
        reinterpret_cast<float4*>(p.records + (size_t)GCAR_OF(ci) * NCG_RECORD_WORDS)[i & 31] = make_float4(d[0], d[1], d[2], d[3]);
    }
#undef SLOT_OF
#undef GCAR_OF
}

// the observation of the reset state of each track; one warp per track
__global__ void __launch_bounds__(32) ncg_reset_obs_kernel(const float* blob, const long long* track_off, float* reset_obs, const StartPose sp0,
                                                            int rc, float gdx, float gdy) {
    __shared__ float s_rec[NCG_RECORD_WORDS];
    __shared__ float s_o[40];
    // one block per row: (track, car k of the env) with rc = cars per env on a start grid, else rc = 1
    const int lane = threadIdx.x, tid = blockIdx.x / rc;
    const StartPose sp = rc > 1 ? cc_start_pose(sp0, blockIdx.x % rc, gdx, gdy) : sp0;
    const float* g = blob + track_off[tid];
    Track T = track_view(g, g);
    if (lane == 0) { reset_record(s_rec, T, true, (uint32_t)tid, sp); observe_state(s_rec, s_o); }
    __syncwarp();
    unsigned tests = 0;
    if (lane < 16) cast_rays<1, false>(T, s_rec[NCG_R_X], s_rec[NCG_R_Y], s_rec[NCG_R_ANGLE], lane, s_o + 22, &tests);
    __syncwarp();
    for (int k = lane; k < NCG_OBS_DIM; k += 32) reset_obs[(size_t)blockIdx.x * NCG_OBS_DIM + k] = s_o[k];
}

// reset of masked envs + their initial observation; one warp per car (rays over lanes)
__global__ void __launch_bounds__(256) ncg_reset_kernel(float* records, const float* blob, const long long* track_off, int E, int C,
                                                         const uint8_t* mask, const int* track_id, int fresh, float* obs, const StartPose sp0,
                                                         int car_contacts, float gdx, float gdy, float* cc_pairs) {
    const int lane = threadIdx.x & 31;
    const int car = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (car >= E * C) return;
    const int env = car / C;
    if (mask && !mask[env]) return;
    float* R = records + (size_t)car * NCG_RECORD_WORDS;
    __shared__ float s_obs[8][40];
    float* so = s_obs[threadIdx.x >> 5];
    uint32_t tid = track_id ? (uint32_t)track_id[env] : f2u(R[NCG_R_TRACK]);
    const float* g = blob + track_off[tid];
    Track T = track_view(g, g);
    const StartPose sp = car_contacts ? cc_start_pose(sp0, car - env * C, gdx, gdy) : sp0;
    if (lane == 0) { reset_record(R, T, fresh != 0, tid, sp); observe_state(R, so); }
    if (car_contacts && fresh && car == env * C) for (int w_ = lane; w_ < NCG_CC_STRIDE; w_ += 32) cc_pairs[(size_t)env * NCG_CC_STRIDE + w_] = 0.0f;
    __syncwarp();
    if (obs) {
        unsigned tests = 0;
        if (lane < 16) cast_rays<1, false>(T, R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], lane, so + 22, &tests);
        __syncwarp();
        for (int k = lane; k < NCG_OBS_DIM; k += 32) obs[(size_t)car * NCG_OBS_DIM + k] = so[k];
    }
}

}  // namespace

struct NcgHandle {
    NcgConfig cfg; int N;
    float* d_records = nullptr; float* d_blob = nullptr; long long* d_track_off = nullptr; int n_tracks = 0;
    float* d_reset_obs = nullptr;
    float* d_cc_pairs = nullptr; void* d_cc_worlds = nullptr;   // car_contacts only: pair tables, per-car World scratch
    float2* d_vel_hist = nullptr;                    // track_info only: the last NCG_VEL_HISTORY pre-step velocities of every car
    std::vector<long long> h_track_off; std::vector<unsigned> h_stage_words;
    std::vector<int> h_env_track;
    int2* d_cta_tab = nullptr; int n_ctas = 0; int cap_ctas = 0; bool cta_dirty = true;   // groups of <= 32 car slots (one CTA each, or two per CTA)
    int2* d_pair_tab = nullptr; int n_pairs = 0; int cap_pairs = 0;       // groups paired by track for the two-physics-warp shape
    int* d_slot_env = nullptr; int cap_slots = 0; bool identity = true;   // slot -> env list (envs ordered by track); unused while the map is sorted
    int* d_env_track = nullptr;                      // [E] device copy of the env -> track map; the redraw writes it
    int redraw = 0; unsigned long long redraw_seed = 0; unsigned steps_taken = 0;
    int* p_redrawn = nullptr;                        // page-locked, device-mapped: set by a step that moved an env to another track
    std::vector<int> h_tmp_track;
    DevStats* d_stats = nullptr;
    bool was_reset = false;
    unsigned step_base = 0, car_base = 0;              // Philox counter offsets of ncg_rollout (ncg_set_rollout_base)
    float* d_ep_return = nullptr; int* d_ep_length = nullptr; int* d_ep_any = nullptr;   // ncg_set_episode_outputs
    int rays_per_lane = 0; int num_sms = 0; int max_smem = 0;
    long long launches = 0;
    // host-buffer path
    cudaStream_t stream = nullptr;
    void* d_actions = nullptr; float* d_obs = nullptr; float* d_final = nullptr; float* d_reward = nullptr; uint8_t* d_term = nullptr; uint8_t* d_trunc = nullptr;
    uint8_t* d_mask = nullptr; int* d_tid = nullptr;
    void* p_actions = nullptr; float* p_obs = nullptr; float* p_final = nullptr; float* p_reward = nullptr; uint8_t* p_flags = nullptr;
    void* d_pack = nullptr; void* p_pack = nullptr; size_t pack_bytes = 0;
    int* p_any_done = nullptr;                       // page-locked, device-mapped flag word of ncg_step_mapped
};

namespace {

StartPose start_pose(const NcgHandle* h) { StartPose sp; sp.x = h->cfg.start_x; sp.y = h->cfg.start_y; sp.a = h->cfg.start_angle; return sp; }

// Whole envs per CTA: at most CPB car slots, fewer when that spreads a small batch over all SMs (4096 single-car
// envs: 28 per CTA = 147 CTAs on 148 SMs instead of 128 CTAs of 32).
int envs_per_cta(int E, int C, int sms) {
    const int epb_max = CPB / C;                               // C <= 10 < CPB
    const int min_ctas = (E + epb_max - 1) / epb_max;
    if (sms <= 0) sms = 148;
    const int target = ((min_ctas + sms - 1) / sms) * sms;     // whole waves of one CTA per SM
    int epb = (E + target - 1) / target;
    if (epb < 1) epb = 1;
    if (epb > epb_max) epb = epb_max;
    return epb;
}
// CTA table: consecutive envs are cut into CTAs of at most envs_per_cta, and never across a track boundary, so every CTA
// stages exactly one track table.
static int count_ctas(const int* env_track, int E, int epb) {
    int n = 0, e = 0;
    while (e < E) { int k = 1; while (k < epb && e + k < E && env_track[e + k] == env_track[e]) ++k; ++n; e += k; }
    return n;
}
void plan_ctas(const int* env_track, int E, int C, int sms, std::vector<int2>& tab) {
    int epb = envs_per_cta(E, C, sms);
    // track boundaries add CTAs (a CTA never spans two tracks); grow the CTAs a little if that spills the plan into one
    // more wave of SMs than the batch needs (4096 envs over 8 tracks: 29 per CTA = 144 CTAs, not 28 = 152 on 148 SMs)
    const int epb_max = CPB / C;
    if (sms <= 0) sms = 148;
    const int waves = (count_ctas(env_track, E, epb_max) + sms - 1) / sms;
    while (epb < epb_max && count_ctas(env_track, E, epb) > waves * sms) ++epb;
    tab.clear();
    int e = 0;
    while (e < E) {
        int n = 1;
        while (n < epb && e + n < E && env_track[e + n] == env_track[e]) ++n;
        tab.push_back(make_int2(e, n));
        e += n;
    }
}
// Rebuilt (host side, two small uploads) whenever the env -> track map changes.  Envs are ordered by track (a stable
// counting sort: the slot list) and the groups are cut from that order, so every CTA serves one track whatever the map
// looks like; the records themselves never move.  A map that is already sorted needs no list.
int build_cta_table(NcgHandle* h) {
    const int E = h->cfg.num_envs;
    CUDA_TRY(cudaDeviceSynchronize());                    // no launch may still be reading the tables that are replaced here
    const std::vector<int>& et = h->h_env_track;
    bool sorted = true;
    for (int e = 1; e < E && sorted; ++e) sorted = et[e] >= et[e - 1];
    std::vector<int> slot_env, slot_track;
    const int* seq = et.data();
    if (!sorted) {
        std::vector<int> start(h->n_tracks + 1, 0);
        for (int e = 0; e < E; ++e) ++start[et[e] + 1];
        for (int t = 0; t < h->n_tracks; ++t) start[t + 1] += start[t];
        slot_env.resize(E); slot_track.resize(E);
        for (int e = 0; e < E; ++e) { const int s = start[et[e]]++; slot_env[s] = e; slot_track[s] = et[e]; }
        seq = slot_track.data();
        if (E > h->cap_slots) {
            cudaFree(h->d_slot_env); h->d_slot_env = nullptr; h->cap_slots = 0;
            CUDA_TRY(cudaMalloc(&h->d_slot_env, (size_t)E * sizeof(int)));
            h->cap_slots = E;
        }
        CUDA_TRY(cudaMemcpyAsync(h->d_slot_env, slot_env.data(), (size_t)E * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    h->identity = sorted;
    std::vector<int2> tab;
    plan_ctas(seq, E, h->cfg.cars_per_env, h->num_sms, tab);
    // (each table has its own capacity: the allocation only ever grows, and growing one never touches the other)
    if ((int)tab.size() > h->cap_ctas) {
        cudaFree(h->d_cta_tab); h->d_cta_tab = nullptr; h->cap_ctas = 0;
        CUDA_TRY(cudaMalloc(&h->d_cta_tab, tab.size() * sizeof(int2)));
        h->cap_ctas = (int)tab.size();
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_cta_tab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    h->n_ctas = (int)tab.size();
    // neighbouring groups of one track, two by two (a group without such a neighbour stays alone)
    std::vector<int2> pairs;
    for (size_t g = 0; g < tab.size();) {
        const bool two = g + 1 < tab.size() && seq[tab[g + 1].x] == seq[tab[g].x];
        pairs.push_back(make_int2((int)g, two ? (int)g + 1 : -1));
        g += two ? 2 : 1;
    }
    if ((int)pairs.size() > h->cap_pairs) {
        cudaFree(h->d_pair_tab); h->d_pair_tab = nullptr; h->cap_pairs = 0;
        CUDA_TRY(cudaMalloc(&h->d_pair_tab, pairs.size() * sizeof(int2)));
        h->cap_pairs = (int)pairs.size();
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_pair_tab, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_env_track, et.data(), (size_t)E * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    // (pageable sources: the copies are staged before the calls return; the plan must be in place before the caller's
    // stream runs the step)
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->n_pairs = (int)pairs.size();
    h->cta_dirty = false;
    return NCG_OK;
}

// Random-track mode: a step may have moved finished envs to other tracks (the kernel wrote the new ids to d_env_track and
// raised the mapped flag).  Before the next launch the host map and the launch plan follow.  The flag is only meaningful
// once the previous step has completed, hence the synchronisation of the caller's stream (the host-buffer paths have
// already waited for it).
int follow_redraws(NcgHandle* h, cudaStream_t s) {
    if (!h->redraw) return NCG_OK;
    CUDA_TRY(cudaStreamSynchronize(s));
    if (!*(volatile int*)h->p_redrawn) return NCG_OK;
    *h->p_redrawn = 0;
    h->h_tmp_track.resize(h->cfg.num_envs);
    CUDA_TRY(cudaMemcpy(h->h_tmp_track.data(), h->d_env_track, (size_t)h->cfg.num_envs * sizeof(int), cudaMemcpyDeviceToHost));
    for (int e = 0; e < h->cfg.num_envs; ++e) {
        const int t = h->h_tmp_track[e];
        if (t >= 0 && t < h->n_tracks) h->h_env_track[e] = t;
    }
    h->cta_dirty = true;
    return NCG_OK;
}

int launch_step(NcgHandle* h, KParams& p, cudaStream_t s) {
    { int rc = follow_redraws(h, s); if (rc) return rc; }
    if (h->cta_dirty) { int rc = build_cta_table(h); if (rc) return rc; }
    p.redraw = (h->redraw && p.T == 1 && p.auto_reset && h->n_tracks > 1) ? 1 : 0;
    p.redraw_step = h->steps_taken; h->steps_taken += (unsigned)p.T;
    const int sms = h->num_sms > 0 ? h->num_sms : 148;
    // rays per lane: 2 (8 ray warps per CTA) while the batch is at most one CTA per SM and latency-bound, 4 (4 ray warps,
    // better lane balance and fewer instructions per car-step) beyond that; measured in profiles/.  Two things that were
    // measured and are worse: cutting the batch into smaller CTAs so that they fill whole waves of resident slots evenly
    // (a CTA's step time is set by the physics warp's dependent chain, so fewer cars per CTA only lowers the work per
    // chain), and a fourth resident CTA per SM (28 car slots, wall AABBs left in L2, 96 registers with spills).
    const int RPL = h->rays_per_lane ? h->rays_per_lane : (h->n_ctas > sms ? 4 : 2);
    // physics warps per CTA: one CTA per group (three resident per SM), or one per pair of groups (two resident per SM =
    // four physics warps).  Measured per resident wave the pair shape is ~1.2x slower (12 ray warps per SM serve 128 cars
    // instead of 96), so it is chosen when it saves enough waves: 16384 envs run as 256 CTAs in one wave instead of 512 in
    // two (+23 %), 8192 ten-car envs in 5 waves instead of 7 (+30 %), the 65536-env track mix in 4 instead of 5 (+7 %).
    int PW = 1;
    if (h->n_ctas > 2 * sms) {
        const int waves1 = (h->n_ctas + 3 * sms - 1) / (3 * sms), waves2 = (h->n_pairs + 2 * sms - 1) / (2 * sms);
        if (12 * waves2 < 10 * waves1) PW = 2;
    }
    { const char* pw = getenv("NCG_PHYS_WARPS"); if (pw && (atoi(pw) == 1 || atoi(pw) == 2 || atoi(pw) == 4)) PW = atoi(pw); }
    if (PW == 4 && (h->cfg.cars_per_env != 1 || h->n_ctas > sms)) PW = 1;       // the spread shape: single-car envs, one CTA per SM
    p.cta_tab = h->d_cta_tab; p.pair_tab = h->d_pair_tab; p.slot_env = h->identity ? nullptr : h->d_slot_env;
    { const char* ns = getenv("NCG_NO_STAGE"); p.stage = (ns && atoi(ns)) ? 0 : 1; }
    unsigned mx = 0;
    if (p.stage) for (unsigned w : h->h_stage_words) mx = w > mx ? w : mx;
    size_t smem = 0;        // (set below, once the shape is known: the number of step buffers depends on it)
    // rays handed out from a per-CTA queue (longest first) instead of a fixed lane -> rays map: pays once the SM is
    // issue-bound, i.e. with three resident CTAs per SM (measured on B200, daytona: +22 % at 65536 envs, +9 % at 16384,
    // -5 % at 8192 and -7 % at 4096, where a step is bound by latency and the queue's claims and job set-up only add to it)
    p.queue = h->n_ctas > 2 * sms ? 1 : 0;
    { const char* q = getenv("NCG_RAY_QUEUE"); if (q) p.queue = atoi(q) ? 1 : 0; }
    // resident CTAs per SM the register allocation allows: as many as the batch has use for, up to what shared memory
    // (~72 KB per CTA) admits; the 4-rays-per-lane shape (160 threads) fits three
    int minb = h->n_ctas <= sms ? 1 : (h->n_ctas <= 2 * sms || RPL != 4 ? 2 : 3);
    { const char* mb = getenv("NCG_MIN_BLOCKS"); if (mb && atoi(mb) >= 1 && atoi(mb) <= 3) minb = atoi(mb); }
    if (minb == 3 && RPL != 4) minb = 2;
    if (PW == 2) minb = 2;
    if (PW == 4) minb = 1;
    const int nb = minb == 1 ? 3 : 2;                 // step buffers: NB of the kernel
    smem = (size_t)smem_layout(mx, PW == 2 ? 64 : 32, nb).total * 4;
    if (p.stage && h->max_smem > 0 && smem > (size_t)h->max_smem) {      // a user track too large to stage: read it through L1/L2
        p.stage = 0;
        smem = (size_t)smem_layout(0, PW == 2 ? 64 : 32, nb).total * 4;
    }
    void (*k)(KParams) = PW == 4 ? ncg_step_kernel<2, 1, 4> : PW == 2 ? ncg_step_kernel<4, 2, 2>
                       : minb == 1 ? (RPL == 4 ? ncg_step_kernel<4, 1, 1> : ncg_step_kernel<2, 1, 1>)
                       : minb == 2 ? (RPL == 4 ? ncg_step_kernel<4, 2, 1> : ncg_step_kernel<2, 2, 1>)
                                   : ncg_step_kernel<4, 3, 1>;
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (PW == 4) k<<<h->n_ctas, 32 * (4 + 8), smem, s>>>(p);
    else if (PW == 2) k<<<h->n_pairs, 256, smem, s>>>(p);
    else k<<<h->n_ctas, 32 * (1 + 16 / RPL), smem, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    ++h->launches;
    return NCG_OK;
}

KParams base_params(NcgHandle* h) {
    KParams p; memset(&p, 0, sizeof(p));
    p.start = start_pose(h); p.vel_hist = h->d_vel_hist;
    p.car_contacts = h->cfg.car_contacts; p.grid_dx = h->cfg.grid_dx; p.grid_dy = h->cfg.grid_dy; p.cc_pairs = h->d_cc_pairs; p.cc_worlds = (World*)h->d_cc_worlds;
    p.n_tracks = h->n_tracks; p.env_track = h->d_env_track; p.redraw_seed = h->redraw_seed; p.redrawn = h->p_redrawn;
    p.records = h->d_records; p.blob = h->d_blob; p.track_off = h->d_track_off; p.reset_obs = h->d_reset_obs;
    p.E = h->cfg.num_envs; p.C = h->cfg.cars_per_env; p.discrete = h->cfg.discrete; p.reset_on_lap = h->cfg.reset_on_lap;
    p.auto_reset = h->cfg.auto_reset; p.contacts = h->cfg.contacts; p.track_info = h->cfg.track_info; p.stats = h->d_stats; p.T = 1;
    { const char* d = getenv("NCG_DEBUG_SKIP"); p.debug_skip = d ? atoi(d) : 0; }   // profiling only: 1 = no rays, 2 = no physics
    return p;
}

}  // namespace

extern "C" {

const char* ncg_last_error(void) { return g_err.c_str(); }
int ncg_version(void) { return 1; }

int ncg_create(const NcgConfig* cfg, NcgHandle** out) {
    if (!cfg || !out) return fail(NCG_E_INVALID, "null argument");
    if (cfg->cars_per_env < 1 || cfg->cars_per_env > NCG_MAX_CARS) return fail(NCG_E_INVALID, "Number of cars must be between 1 and 10");
    if (cfg->num_envs < 1) return fail(NCG_E_INVALID, "num_envs must be >= 1");
    if (cfg->car_contacts && !(cfg->grid_dx > 0.0f && cfg->grid_dy > 0.0f)) return fail(NCG_E_INVALID, "car_contacts needs a start grid: grid_dx, grid_dy > 0");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(NCG_E_INVALID, "no such CUDA device");
    CUDA_TRY(cudaSetDevice(cfg->device));
    NcgHandle* h = new NcgHandle();
    CUDA_TRY(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
    CUDA_TRY(cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
    h->cfg = *cfg; h->N = cfg->num_envs * cfg->cars_per_env;
    const char* g = getenv("NCG_RAYS_PER_LANE");
    int rpl = g ? atoi(g) : 0;
    if (rpl != 2 && rpl != 4) rpl = 0;                            // 0 = chosen per launch from the batch size
    h->rays_per_lane = rpl;
    size_t N = (size_t)h->N, E = (size_t)cfg->num_envs;
    CUDA_TRY(cudaMalloc(&h->d_records, N * NCG_RECORD_WORDS * 4));
    CUDA_TRY(cudaMemset(h->d_records, 0, N * NCG_RECORD_WORDS * 4));
    if (cfg->track_info) CUDA_TRY(cudaMalloc(&h->d_vel_hist, N * NCG_VEL_HISTORY * sizeof(float2)));
    if (cfg->car_contacts) {
        CUDA_TRY(cudaMalloc(&h->d_cc_pairs, E * NCG_CC_STRIDE * 4)); CUDA_TRY(cudaMemset(h->d_cc_pairs, 0, E * NCG_CC_STRIDE * 4));
        CUDA_TRY(cudaMalloc(&h->d_cc_worlds, N * sizeof(World)));
    }
    CUDA_TRY(cudaMalloc(&h->d_stats, sizeof(DevStats)));
    CUDA_TRY(cudaMemset(h->d_stats, 0, sizeof(DevStats)));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    // host path: results packed as obs | reward | terminated | truncated so one D2H copy brings a whole step back
    h->pack_bytes = N * NCG_OBS_DIM * 4 + N * 4 + 2 * E;
    CUDA_TRY(cudaMalloc(&h->d_actions, N * 8)); CUDA_TRY(cudaMalloc(&h->d_pack, h->pack_bytes)); CUDA_TRY(cudaMalloc(&h->d_final, N * NCG_OBS_DIM * 4));
    h->d_obs = (float*)h->d_pack; h->d_reward = h->d_obs + N * NCG_OBS_DIM; h->d_term = (uint8_t*)(h->d_reward + N); h->d_trunc = h->d_term + E;
    CUDA_TRY(cudaMalloc(&h->d_mask, E)); CUDA_TRY(cudaMalloc(&h->d_tid, E * 4));
    CUDA_TRY(cudaMallocHost(&h->p_actions, N * 8)); CUDA_TRY(cudaMallocHost(&h->p_pack, h->pack_bytes)); CUDA_TRY(cudaMallocHost(&h->p_final, N * NCG_OBS_DIM * 4));
    h->p_obs = (float*)h->p_pack; h->p_reward = h->p_obs + N * NCG_OBS_DIM; h->p_flags = (uint8_t*)(h->p_reward + N);
    CUDA_TRY(cudaHostAlloc((void**)&h->p_any_done, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    CUDA_TRY(cudaHostAlloc((void**)&h->p_redrawn, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    *h->p_redrawn = 0;
    CUDA_TRY(cudaMalloc(&h->d_env_track, E * sizeof(int)));
    CUDA_TRY(cudaMemset(h->d_env_track, 0, E * sizeof(int)));
    h->h_env_track.assign(E, 0);
    *out = h;
    return NCG_OK;
}

int ncg_destroy(NcgHandle* h) {
    if (!h) return NCG_OK;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->d_cc_pairs); cudaFree(h->d_cc_worlds); cudaFree(h->d_vel_hist); cudaFree(h->d_records); cudaFree(h->d_blob); cudaFree(h->d_track_off); cudaFree(h->d_stats); cudaFree(h->d_reset_obs);
    cudaFree(h->d_actions); cudaFree(h->d_pack); cudaFree(h->d_final);
    cudaFree(h->d_mask); cudaFree(h->d_tid); cudaFree(h->d_cta_tab); cudaFree(h->d_pair_tab); cudaFree(h->d_slot_env); cudaFree(h->d_env_track);
    cudaFreeHost(h->p_redrawn);
    cudaFreeHost(h->p_actions); cudaFreeHost(h->p_pack); cudaFreeHost(h->p_final); cudaFreeHost(h->p_any_done);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return NCG_OK;
}

int ncg_upload_tracks(NcgHandle* h, const float* h_blob, const int64_t* h_offsets, int32_t n_tracks) {
    if (!h || !h_blob || !h_offsets || n_tracks < 1) return fail(NCG_E_INVALID, "bad track upload");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    for (int i = 0; i <= n_tracks; ++i) if (h_offsets[i] % 4) return fail(NCG_E_INVALID, "track offsets must be multiples of 4 words");
    cudaFree(h->d_blob); cudaFree(h->d_track_off); cudaFree(h->d_reset_obs); h->d_blob = nullptr; h->d_track_off = nullptr; h->d_reset_obs = nullptr;
    size_t words = (size_t)h_offsets[n_tracks];
    CUDA_TRY(cudaMalloc(&h->d_blob, words * 4));
    CUDA_TRY(cudaMemcpy(h->d_blob, h_blob, words * 4, cudaMemcpyHostToDevice));
    h->h_track_off.assign(h_offsets, h_offsets + n_tracks + 1);
    CUDA_TRY(cudaMalloc(&h->d_track_off, (n_tracks + 1) * sizeof(long long)));
    CUDA_TRY(cudaMemcpy(h->d_track_off, h->h_track_off.data(), (n_tracks + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    h->h_stage_words.clear();
    for (int i = 0; i < n_tracks; ++i) { uint32_t w; memcpy(&w, h_blob + h_offsets[i] + TH_STAGE_WORDS, 4); h->h_stage_words.push_back(w); }
    h->n_tracks = n_tracks;
    const int rc = h->cfg.car_contacts ? h->cfg.cars_per_env : 1;             // reset rows per track (a start grid: one per car of an env)
    CUDA_TRY(cudaMalloc(&h->d_reset_obs, (size_t)n_tracks * rc * NCG_OBS_DIM * 4));
    ncg_reset_obs_kernel<<<n_tracks * rc, 32>>>(h->d_blob, h->d_track_off, h->d_reset_obs, start_pose(h), rc, h->cfg.grid_dx, h->cfg.grid_dy);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    ++h->launches;
    return NCG_OK;
}

int ncg_reset(NcgHandle* h, const uint8_t* d_env_mask, const int32_t* d_track_id, int32_t fresh, float* d_obs, void* stream) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    if (!h->d_blob) return fail(NCG_E_STATE, "ncg_upload_tracks must be called before ncg_reset");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (!h->was_reset && (d_env_mask || !fresh)) return fail(NCG_E_STATE, "the first reset must be a full fresh reset");
    const int E = h->cfg.num_envs;
    std::vector<int> ids; std::vector<uint8_t> mk;
    if (d_track_id) {
        // device-side ids: brought back and range-checked BEFORE the kernel follows them into the track blob (one
        // synchronisation of the caller's stream; the env -> track map on the host needs them anyway for the CTA table)
        ids.resize(E); mk.resize(d_env_mask ? E : 0);
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        CUDA_TRY(cudaMemcpy(ids.data(), d_track_id, (size_t)E * 4, cudaMemcpyDeviceToHost));
        if (d_env_mask) CUDA_TRY(cudaMemcpy(mk.data(), d_env_mask, E, cudaMemcpyDeviceToHost));
        for (int e = 0; e < E; ++e) if ((!d_env_mask || mk[e]) && (ids[e] < 0 || ids[e] >= h->n_tracks)) return fail(NCG_E_INVALID, "track id out of range");
    }
    const int threads = 256, cars_per_block = threads / 32;
    const int grid = (h->N + cars_per_block - 1) / cars_per_block;
    ncg_reset_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(h->d_records, h->d_blob, h->d_track_off, h->cfg.num_envs, h->cfg.cars_per_env,
                                                                d_env_mask, d_track_id, fresh, d_obs, start_pose(h),
                                                                h->cfg.car_contacts, h->cfg.grid_dx, h->cfg.grid_dy, h->d_cc_pairs);
    CUDA_TRY(cudaGetLastError());
    ++h->launches;
    h->was_reset = true;
    if (d_track_id) {                            // only now, with the launch accepted, does the host map follow
        for (int e = 0; e < E; ++e) if (!d_env_mask || mk[e]) h->h_env_track[e] = ids[e];
        h->cta_dirty = true;
    }
    return NCG_OK;
}

int ncg_step(NcgHandle* h, const void* d_actions, float* d_obs, float* d_reward, uint8_t* d_terminated, uint8_t* d_truncated,
             float* d_final_obs, void* stream) {
    if (!h || !d_actions || !d_obs || !d_reward) return fail(NCG_E_INVALID, "null argument");
    if (!h->was_reset) return fail(NCG_E_STATE, "Environment not properly initialized. Call reset() first.");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    KParams p = base_params(h);
    p.actions = d_actions; p.obs = d_obs; p.reward = d_reward; p.term = d_terminated; p.trunc = d_truncated; p.final_obs = d_final_obs;
    p.ep_return = h->d_ep_return; p.ep_length = h->d_ep_length; p.any_done = h->d_ep_any;
    return launch_step(h, p, (cudaStream_t)stream);
}

int ncg_rollout(NcgHandle* h, int32_t steps, uint64_t seed, int32_t mode, float* d_obs_rollout, float* d_reward_rollout,
                uint8_t* d_done_rollout, float* d_obs_last, void* stream) {
    if (!h || steps < 1) return fail(NCG_E_INVALID, "bad rollout arguments");
    if (!h->was_reset) return fail(NCG_E_STATE, "Environment not properly initialized. Call reset() first.");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    KParams p = base_params(h);
    p.T = steps; p.seed = seed; p.mode = mode; p.step_base = h->step_base; p.car_base = h->car_base; p.auto_reset = 1;
    p.obs_roll = d_obs_rollout; p.rew_roll = d_reward_rollout; p.done_roll = d_done_rollout; p.obs = d_obs_last; p.reward = nullptr;
    if (!d_reward_rollout) { p.reward = h->d_reward; }
    p.term = h->d_term; p.trunc = h->d_trunc;
    h->step_base += (unsigned)steps;
    return launch_step(h, p, (cudaStream_t)stream);
}

int ncg_reset_host(NcgHandle* h, const uint8_t* h_env_mask, const int32_t* h_track_id, int32_t fresh, float* h_obs) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t E = h->cfg.num_envs;
    if (h_track_id) for (size_t e = 0; e < E; ++e) if ((!h_env_mask || h_env_mask[e]) && (h_track_id[e] < 0 || h_track_id[e] >= h->n_tracks)) return fail(NCG_E_INVALID, "track id out of range");
    if (h_env_mask) CUDA_TRY(cudaMemcpyAsync(h->d_mask, h_env_mask, E, cudaMemcpyHostToDevice, h->stream));
    if (h_track_id) CUDA_TRY(cudaMemcpyAsync(h->d_tid, h_track_id, E * 4, cudaMemcpyHostToDevice, h->stream));
    int rc = ncg_reset(h, h_env_mask ? h->d_mask : nullptr, h_track_id ? h->d_tid : nullptr, fresh, h->d_obs, h->stream);
    if (rc) return rc;
    if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs, h->d_obs, (size_t)h->N * NCG_OBS_DIM * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return NCG_OK;
}

// One step through the pinned staging buffers: actions are read from p_actions, results land in p_pack (and p_final
// when an env finished).  *any_done tells the caller whether p_final holds terminal observations.
static int step_pinned(NcgHandle* h, bool want_final, int* any_done) {
    const size_t N = h->N, E = h->cfg.num_envs, abytes = h->cfg.discrete ? N * 4 : N * 8;
    CUDA_TRY(cudaMemcpyAsync(h->d_actions, h->p_actions, abytes, cudaMemcpyHostToDevice, h->stream));
    int rc = ncg_step(h, h->d_actions, h->d_obs, h->d_reward, h->d_term, h->d_trunc, want_final ? h->d_final : nullptr, h->stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->p_pack, h->d_pack, h->pack_bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int done = 0;
    for (size_t e = 0; e < 2 * E; ++e) done |= h->p_flags[e];
    if (want_final && done && h->cfg.auto_reset) {
        CUDA_TRY(cudaMemcpyAsync(h->p_final, h->d_final, N * NCG_OBS_DIM * 4, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    if (any_done) *any_done = done ? 1 : 0;
    return NCG_OK;
}

int ncg_step_host(NcgHandle* h, const void* h_actions, float* h_obs, float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated,
                  float* h_final_obs) {
    if (!h || !h_actions || !h_obs || !h_reward || !h_terminated || !h_truncated) return fail(NCG_E_INVALID, "null argument");
    if (!h->was_reset) return fail(NCG_E_STATE, "Environment not properly initialized. Call reset() first.");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t N = h->N, E = h->cfg.num_envs, abytes = h->cfg.discrete ? N * 4 : N * 8;
    memcpy(h->p_actions, h_actions, abytes);
    int done = 0;
    int rc = step_pinned(h, h_final_obs != nullptr, &done);
    if (rc) return rc;
    memcpy(h_obs, h->p_obs, N * NCG_OBS_DIM * 4); memcpy(h_reward, h->p_reward, N * 4);
    memcpy(h_terminated, h->p_flags, E); memcpy(h_truncated, h->p_flags + E, E);
    if (h_final_obs && done && h->cfg.auto_reset) memcpy(h_final_obs, h->p_final, N * NCG_OBS_DIM * 4);
    return NCG_OK;
}

int ncg_host_buffers(NcgHandle* h, void** actions, float** obs, float** reward, uint8_t** terminated, uint8_t** truncated, float** final_obs) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    if (actions) *actions = h->p_actions;
    if (obs) *obs = h->p_obs;
    if (reward) *reward = h->p_reward;
    if (terminated) *terminated = h->p_flags;
    if (truncated) *truncated = h->p_flags + h->cfg.num_envs;
    if (final_obs) *final_obs = h->p_final;
    return NCG_OK;
}

int ncg_step_pinned(NcgHandle* h, int32_t want_final, int32_t* any_done) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    if (!h->was_reset) return fail(NCG_E_STATE, "Environment not properly initialized. Call reset() first.");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    return step_pinned(h, want_final != 0, any_done);
}

int32_t ncg_plan_ctas(const int32_t* h_env_track, int32_t num_envs, int32_t cars_per_env, int32_t num_sms, int32_t* h_first_env,
                      int32_t* h_num_envs, int32_t capacity) {
    if (!h_env_track || num_envs < 1 || cars_per_env < 1 || cars_per_env > NCG_MAX_CARS) return fail(NCG_E_INVALID, "bad ncg_plan_ctas arguments");
    std::vector<int2> tab;
    plan_ctas(h_env_track, num_envs, cars_per_env, num_sms, tab);
    for (int i = 0; i < (int)tab.size() && i < capacity; ++i) { if (h_first_env) h_first_env[i] = tab[i].x; if (h_num_envs) h_num_envs[i] = tab[i].y; }
    return (int32_t)tab.size();
}

int ncg_host_alloc(size_t bytes, void** out) {
    if (!out || !bytes) return fail(NCG_E_INVALID, "bad ncg_host_alloc arguments");
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    return NCG_OK;
}
int ncg_host_free(void* p) {
    if (p) CUDA_TRY(cudaFreeHost(p));
    return NCG_OK;
}

// One step with every buffer in page-locked, device-mapped host memory: the kernel reads the actions and writes
// observations / rewards / flags over PCIe itself, so a step is one launch and one stream synchronise.
int ncg_step_mapped(NcgHandle* h, const void* h_actions, float* h_obs, float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated,
                    float* h_final_obs, float* h_ep_return, int32_t* h_ep_length, int32_t* any_done) {
    if (!h || !h_actions || !h_obs || !h_reward || !h_terminated || !h_truncated) return fail(NCG_E_INVALID, "null argument");
    if (!h->was_reset) return fail(NCG_E_STATE, "Environment not properly initialized. Call reset() first.");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    KParams p = base_params(h);
    p.actions = h_actions; p.obs = h_obs; p.reward = h_reward; p.term = h_terminated; p.trunc = h_truncated; p.final_obs = h_final_obs;
    p.ep_return = h_ep_return; p.ep_length = h_ep_length; p.any_done = h->p_any_done;
    *h->p_any_done = 0;
    int rc = launch_step(h, p, h->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (any_done) *any_done = *h->p_any_done;
    return NCG_OK;
}

int ncg_get_state(NcgHandle* h, float* d_records, void* stream) {
    if (!h || !d_records) return fail(NCG_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(d_records, h->d_records, (size_t)h->N * NCG_RECORD_WORDS * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return NCG_OK;
}
int ncg_set_state(NcgHandle* h, const float* d_records, void* stream) {
    if (!h || !d_records) return fail(NCG_E_INVALID, "null argument");
    if (!h->d_blob) return fail(NCG_E_STATE, "ncg_upload_tracks must be called first");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    // the records carry the env -> track map (word NCG_R_TRACK of each env's first car): read it from the caller's copy,
    // range-check it and let the CTA table follow, as ncg_set_state_host does -- a CTA stages ONE track table and would
    // otherwise step foreign records against the wrong walls.  One synchronisation of the caller's stream.
    const int E = h->cfg.num_envs;
    std::vector<uint32_t> tid(E);
    const size_t pitch = (size_t)h->cfg.cars_per_env * NCG_RECORD_WORDS * 4;
    CUDA_TRY(cudaMemcpy2DAsync(tid.data(), 4, d_records + NCG_R_TRACK, pitch, 4, (size_t)E, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    for (int e = 0; e < E; ++e) if ((int)tid[e] < 0 || (int)tid[e] >= h->n_tracks) return fail(NCG_E_INVALID, "record names a track id that was not uploaded");
    CUDA_TRY(cudaMemcpyAsync(h->d_records, d_records, (size_t)h->N * NCG_RECORD_WORDS * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    bool changed = false;
    for (int e = 0; e < E; ++e) if (h->h_env_track[e] != (int)tid[e]) { h->h_env_track[e] = (int)tid[e]; changed = true; }
    if (changed) h->cta_dirty = true;
    h->was_reset = true;
    return NCG_OK;
}
int ncg_get_state_host(NcgHandle* h, float* h_records) {
    if (!h || !h_records) return fail(NCG_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h_records, h->d_records, (size_t)h->N * NCG_RECORD_WORDS * 4, cudaMemcpyDeviceToHost));
    return NCG_OK;
}
int ncg_set_state_host(NcgHandle* h, const float* h_records) {
    if (!h || !h_records) return fail(NCG_E_INVALID, "null argument");
    if (!h->d_blob) return fail(NCG_E_STATE, "ncg_upload_tracks must be called first");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    for (int e = 0; e < h->cfg.num_envs; ++e) {
        uint32_t t; memcpy(&t, h_records + (size_t)e * h->cfg.cars_per_env * NCG_RECORD_WORDS + NCG_R_TRACK, 4);
        if ((int)t < 0 || (int)t >= h->n_tracks) return fail(NCG_E_INVALID, "record names a track id that was not uploaded");
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h->d_records, h_records, (size_t)h->N * NCG_RECORD_WORDS * 4, cudaMemcpyHostToDevice));
    for (int e = 0; e < h->cfg.num_envs; ++e) {
        uint32_t t; memcpy(&t, h_records + (size_t)e * h->cfg.cars_per_env * NCG_RECORD_WORDS + NCG_R_TRACK, 4);
        h->h_env_track[e] = (int)t;
    }
    h->cta_dirty = true;
    h->was_reset = true;
    return NCG_OK;
}

int ncg_get_velocity_history_host(NcgHandle* h, float* h_out) {
    if (!h || !h_out) return fail(NCG_E_INVALID, "null argument");
    if (!h->d_vel_hist) return fail(NCG_E_STATE, "the velocity history is kept only with NcgConfig.track_info = 1");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h_out, h->d_vel_hist, (size_t)h->N * NCG_VEL_HISTORY * sizeof(float2), cudaMemcpyDeviceToHost));
    return NCG_OK;
}

int ncg_read_stats(NcgHandle* h, NcgStats* out, int32_t reset) {
    if (!h || !out) return fail(NCG_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    DevStats s;
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(&s, h->d_stats, sizeof(s), cudaMemcpyDeviceToHost));
    out->car_steps = s.car_steps; out->episodes = s.episodes; out->laps = s.laps; out->ray_tests = s.ray_tests;
    out->contact_steps = s.contact_steps; out->toi_events = s.toi_events; out->overflow = s.overflow; out->return_sum = s.return_sum;
    if (reset) CUDA_TRY(cudaMemset(h->d_stats, 0, sizeof(DevStats)));
    return NCG_OK;
}

int64_t ncg_launch_count(NcgHandle* h) { return h ? h->launches : 0; }

int ncg_set_track_redraw(NcgHandle* h, int32_t enable, uint64_t seed) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    h->redraw = enable ? 1 : 0; h->redraw_seed = seed;
    return NCG_OK;
}

int ncg_get_env_tracks(NcgHandle* h, int32_t* h_out) {
    if (!h || !h_out) return fail(NCG_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (h->redraw) { CUDA_TRY(cudaDeviceSynchronize()); int rc = follow_redraws(h, h->stream); if (rc) return rc; }
    for (int e = 0; e < h->cfg.num_envs; ++e) h_out[e] = h->h_env_track[e];
    return NCG_OK;
}

int ncg_set_rollout_base(NcgHandle* h, uint32_t car_base, uint32_t step_base) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    h->car_base = car_base; h->step_base = step_base;
    return NCG_OK;
}

int ncg_set_episode_outputs(NcgHandle* h, float* d_ep_return, int32_t* d_ep_length, int32_t* d_any_done) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    h->d_ep_return = d_ep_return; h->d_ep_length = d_ep_length; h->d_ep_any = d_any_done;
    return NCG_OK;
}

}  // extern "C"
