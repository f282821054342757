// ncg_step.cuh -- the step kernel (ncg_step_kernel) and what it is launched with (KParams).  A header because the kernel is
// compiled in two translation units: ncg_b200.cu instantiates the default shapes, ncg_b200_cc.cu the shared-world variant
// (CC = true).  Instantiating the variant next to the default kernels moved their out-of-line contact code and cost the
// contact-heavy ("driving") benchmark 10 % (125 -> 112 M car-steps/s at 4096 envs), hence its own unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ncg_car.cuh"

using namespace ncg;

namespace {

// NCG_TIMELINE (variant build, measurement only): clock64 stamps of one CTA's warps per step and phase
#ifdef NCG_TIMELINE
__device__ long long g_timeline[12][128][6];
__device__ __forceinline__ long long tl_clock() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory"); return c; }
#define TL(ev) do { const long long c_ = tl_clock(); if (blockIdx.x == 5 && lane == 0 && t < 128 && warp < 12) g_timeline[warp][t][ev] = c_; } while (0)
// a stamp taken after a barrier: BAR.SYNC.DEFER_BLOCKING lets the warp run on until its next memory access, so the clock is
// read under a predicate that depends on a shared-memory load issued after the barrier
__device__ __forceinline__ long long tl_clock_after(const float* sm) {
    long long c = 0; const float x = *(const volatile float*)sm;
    asm volatile("{ .reg .pred p; setp.neu.f32 p, %1, 0f7FC01234; @p mov.u64 %0, %%clock64; }" : "+l"(c) : "f"(x) : "memory");
    return c;
}
#define TLB(ev) do { const long long c_ = tl_clock_after(smem); if (blockIdx.x == 5 && lane == 0 && t < 128 && warp < 12) g_timeline[warp][t][ev] = c_; } while (0)
__device__ long long g_cta_cycles[4096][5];
__device__ unsigned long long g_launch_ns[4096][2];   // per launch (index = step_base & 4095): first CTA start, last CTA end (globaltimer, ns)
__device__ __forceinline__ unsigned long long tl_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); return t; }     // per CTA: {kernel cycles, physics-warp wait cycles, ray warp 1 wait cycles, -}
#define TLWAIT(stmt) do { const long long w0_ = tl_clock(); stmt; tl_wait += tl_clock_after(smem) - w0_; } while (0)
#else
#define TL(ev) do { } while (0)
#define TLB(ev) do { } while (0)
#define TLWAIT(stmt) do { stmt; } while (0)
#endif

// NCG_RES_TIMELINE (variant build, measurement only): clock64 sums of CTA 0's phases of a resident step, in res_done_ctr[4 + k]
#ifdef NCG_RES_TIMELINE
#ifndef NCG_RES_TL_CTA
#define NCG_RES_TL_CTA 0
#endif
#define RTL(k) do { if (RES && blockIdx.x == NCG_RES_TL_CTA && lane == 0) { long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); atomicAdd(p.res_done_ctr + 4 + (k), (unsigned long long)(c_ - *(volatile long long*)&s_res_t0)); } } while (0)
#else
#define RTL(k) do { } while (0)
#endif
// NCG_RES_TIMELINE2 (variant build): the same from the ray warps only -- %globaltimer since CTA 0 saw the command, stamped by
// lane 0 of one CTA's first ray warp, so that the physics warp runs exactly the code of the default build
#ifdef NCG_RES_TIMELINE2
#ifndef NCG_RES_TL_CTA
#define NCG_RES_TL_CTA 5
#endif
#define RTG(k) do { if (RES && blockIdx.x == NCG_RES_TL_CTA && warp == PW && lane == 0) atomicAdd(p.res_done_ctr + 4 + (k), res_ns() - *(volatile unsigned long long*)(p.res_done_ctr + 1)); } while (0)
#else
#define RTG(k) do { } while (0)
#endif
struct DevStats { unsigned long long car_steps, episodes, laps, ray_tests, contact_steps, toi_events, overflow; double return_sum; };

struct KParams {
    float* records; const float* blob; const long long* track_off;
    const int2* cta_tab;                                   // per group: {first slot, number of envs}; all of one track, <= 32 cars
    const int* slot_env;                                   // slot -> env, envs ordered by track; NULL = identity (the map is already sorted)
    const int2* pair_tab;                                  // two-physics-warp shape: per CTA the two groups it serves {g0, g1 or -1}
    const int2* cta_stage; const int2* pair_stage;         // per group / per pair: {word offset of the track's table in blob, bytes to stage}
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
    // resident mode (ncg_b200_res.cu, RES = true): the kernel stays on the SMs between the steps of a host-driven loop and is
    // fed through a mailbox in page-locked host memory
    const unsigned long long* res_host_cmd;                // mapped host word: seq << 32 | table generation << 16 | op << 8 | result slot
    unsigned long long* res_dev_cmd;                       // the same word, relayed to device memory by CTA 0
    const unsigned long long* res_host_tab;                // mapped host [NCG_RES_SLOTS][4]: obs, reward, terminated, truncated of a result slot
    unsigned long long* res_dev_tab;                       // device copy of the rows CTA 0 has seen
    unsigned long long* res_done_ctr;                      // CTAs that have finished a step, summed over the launch
    volatile unsigned* res_host_done;                      // mapped host: [0] last completed seq, [1] the kernel left by itself (idle)
    unsigned res_seq0; unsigned long long res_idle_ns;     // seq of the last step before this launch; idle time after which the kernel leaves
    int res_fence_gpu;                                     // experiment (NCG_RESIDENT_FENCE=gpu): no system-scope fence before the done word
    int pdl;                                               // launched with programmatic stream serialization (single-step launches)
};
#define NCG_RES_SLOTS 16
#define NCG_RES_OP_EXIT 1u

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

// ---- resident mode: the mailbox.  The host posts one 64-bit command word per step into page-locked memory; lane 0 of CTA 0's
// physics warp (the "dispatcher") polls it across PCIe and relays it through a device word every CTA polls in L2, so that the
// decision "step seq" / "leave" is taken once for the whole grid (a CTA that timed out on its own while another one saw the
// next command would leave the batch half stepped).
__device__ __forceinline__ unsigned long long res_ld_sys(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long res_ld_gpu(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void res_st_rel(unsigned long long* p, unsigned long long v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
// (acquire-release fences: __threadfence_system() / __threadfence() are the sequentially consistent MEMBAR.SC.*, which this
// release -> count -> acquire -> release chain does not need)
__device__ __forceinline__ void res_fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void res_fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned long long res_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); return t; }
// returns the command of step `want` or an exit word (op != 0); gen / valid: which rows of the result-slot table the device copy holds
__device__ __noinline__ unsigned long long res_dispatch(const KParams& p, unsigned want, unsigned& gen, unsigned& valid) {
    const unsigned long long t0 = res_ns();
    unsigned long long c, seen = 0;
    for (;;) {
        c = res_ld_sys(p.res_host_cmd);
        if (((c >> 8) & 0xffu) != 0u) break;                                   // the host asks the kernel to leave
        if ((unsigned)(c >> 32) == want) {
            seen = res_ns();
            // (no system-scope fence on the way to the actions: the host ordered them before the command (sfence + one store), and
            // every read of them is an uncached PCIe read issued after this one has returned)
            const unsigned g = (unsigned)(c >> 16) & 0xffffu, slot = (unsigned)c & (NCG_RES_SLOTS - 1);
            if (g != gen) { gen = g; valid = 0u; }
            if (!((valid >> slot) & 1u)) {
                res_fence_sys();
                for (int k = 0; k < 4; ++k) p.res_dev_tab[slot * 4 + k] = res_ld_sys(p.res_host_tab + slot * 4 + k);
                valid |= 1u << slot;
            }
            break;
        }
        if (res_ns() - t0 > p.res_idle_ns) {                                   // nobody stepped for a while: give the SMs back
            c = ((unsigned long long)(want - 1u) << 32) | (NCG_RES_OP_EXIT << 8);
            p.res_host_done[1] = 1u;
            break;
        }
    }
    res_st_rel(p.res_dev_cmd, c);
    if (seen) p.res_done_ctr[1] = seen;                                        // (diagnostic: when the grid learnt of the step; after the relay, nothing waits for it)
    return c;
}
__device__ __noinline__ unsigned long long res_wait(const KParams& p, unsigned want) {
    // (relaxed polls of the relay word in L2, one acquire fence when the word has changed; the clock is looked at every 4096 polls)
    for (unsigned long long t0 = 0;;) {
        for (int i = 0; i < 4096; ++i) {
            const unsigned long long c = res_ld_gpu(p.res_dev_cmd);
            if (((c >> 8) & 0xffu) != 0u || (unsigned)(c >> 32) == want) { res_fence_gpu(); return c; }
        }
        // (never expected: CTA 0 relays every command and its own exit; a bound so that no CTA can spin for ever)
        const unsigned long long now = res_ns();
        if (!t0) t0 = now;
        if (now - t0 > p.res_idle_ns + 10000000000ull) return ((unsigned long long)(want - 1u) << 32) | (NCG_RES_OP_EXIT << 8);
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
// CC = the optional shared world (NcgConfig.car_contacts): a separate instantiation, so that the default kernels carry none of
// its code (measured: as a run-time branch it cost the default path 2 % contact-free and 12 % with the driving distribution).
// RES = resident mode (its own translation unit, ncg_b200_res.cu; PW = 1 shapes): the step loop has no end, every step waits for
// a command of the host's mailbox, reads the caller's actions from mapped host memory, writes its results into the result slot the
// command names and reports completion through a grid-wide counter whose last arrival raises the host's done word.
template <int RPL, int MINB, int PW, bool CC = false, bool RES = false>
__global__ void __launch_bounds__(PW == 2 ? (RPL == 2 ? 320 : 256) : 32 * (PW + 16 / RPL), MINB) ncg_step_kernel(KParams p) {
    static_assert(!RES || (PW == 1 && !CC), "resident mode: one physics warp per CTA, no shared world");
    constexpr int RW = PW == 2 ? (RPL == 2 ? 8 : 6) : 16 / RPL;      // ray warps (PW == 2: RPL only picks six or eight of them, the rays come from the queue)
    constexpr int NT = 32 * (PW + RW);
    constexpr int GROUPS = PW == 2 ? 2 : 1;         // groups of the CTA table served by this CTA
    constexpr int SLOTS = 32 * GROUPS;              // car slots: 32 per group
    constexpr int LPW = PW == 4 ? 8 : 32;           // car slots per physics warp: warp w owns slots w*LPW .. w*LPW+LPW-1
    constexpr int LPC = 16 / RPL;                   // fixed ray mapping (PW == 1): lanes per car in a ray warp
    constexpr int CPW = 32 / LPC;                   //                              cars per ray warp
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long s_mbar;
    __shared__ float* s_res_obs[3]; __shared__ int s_res_exit;          // RES only (unreferenced otherwise)
#ifdef NCG_RES_TIMELINE
    __shared__ long long s_res_t0;
#endif
    // step buffers between the physics and the ray warps: three when a CTA has an SM to itself (the physics warp then
    // never waits for a drained buffer; shared memory is not the limit there), two otherwise
    constexpr int NB = MINB == 1 ? 3 : 2;
    constexpr int BAR_POSE = 1, BAR_FULL = 1 + NB, BAR_EMPTY = 1 + 2 * NB, BAR_RES = 1 + 3 * NB;
    const SmemLayout L = smem_layout(0, SLOTS, NB);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef NCG_TIMELINE
    long long tl_wait = 0; const long long tl_k0 = tl_clock();
    if (threadIdx.x == 0) atomicMin(&g_launch_ns[p.step_base & 4095][0], tl_ns());
#endif
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

    // ---- track table: staged by TMA when the whole CTA shares a track (else read through L1/L2).  Issued first: where the
    // table is and how much of it to stage comes with the launch plan (one independent load), not from the first record's
    // track id -> offset table -> table header, a chain of three dependent global loads (1.5 us of a single-step launch).
    const float* staged = nullptr;
    if (p.stage) {
        if (threadIdx.x == 0) {
            const int2 sg = (GROUPS == 1 ? p.cta_stage : p.pair_stage)[blockIdx.x];
            tma_issue(s_track, p.blob + sg.x, (unsigned)sg.y, &s_mbar);
        }
        staged = s_track;
    }
    // Programmatic dependent launch (launch_step sets cudaLaunchAttributeProgrammaticStreamSerialization): this kernel lets the NEXT
    // launch of the stream start its CTAs as SMs free up, and does itself everything that does not depend on the launch before it
    // (the table's bulk copy above, the slot table below) in front of griddepcontrol.wait -- in a loop of single-step launches a
    // CTA's prologue then runs under the stragglers of the step before.  Without a programmatic predecessor both are no-ops.
    if (!RES && p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
    if (!RES && p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");      // records, actions and every output belong to the launch before until here
    __syncthreads();

    // (the track table's bulk copy was issued at the top of the kernel)
    // the car slot this thread serves: its lane (physics warps) or, with the fixed ray mapping, the car its rays belong to
    const int slot = warp < PW ? warp * LPW + lane : (warp - PW) * CPW + lane / LPC;
    const bool active = warp < PW ? (PW == 2 ? lane < (warp == 0 ? n0 : n1) : (lane < LPW && slot < n0)) : (GROUPS == 1 && slot < n0);
    // the caller's action of this lane's car (single-step launches): asked for now, while the records and the table are on
    // their way -- with mapped host memory (ncg_step_mapped) it crosses PCIe, ~2 us that would otherwise sit in front of the dynamics
    const bool synth = p.actions == nullptr;
    float2 act_c = make_float2(0.0f, 0.0f); int act_d = 0;
    if (!RES && !synth && warp < PW && active) {
        if (p.discrete) act_d = ((const int*)p.actions)[s_gcar[slot]]; else act_c = ((const float2*)p.actions)[s_gcar[slot]];
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
    if (RES && threadIdx.x == 0) s_res_exit = 0;
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
        int steps_done = p.T;
        unsigned res_gen = 0xffffffffu, res_valid = 0u;          // RES, the dispatcher's view of the result-slot table
        uint8_t* term_out = p.term; uint8_t* trunc_out = p.trunc;
        for (int t = 0, b = 0; t < p.T; ++t, b = b + 1 == NB ? 0 : b + 1) {
            TL(0);
            TLWAIT(if (t >= NB) bar_sync(BAR_EMPTY + b, NT));   // the ray warps have drained buffer b (step t-NB)
            TL(1);
            float* rew_out = p.rew_roll ? p.rew_roll + (size_t)t * N : p.reward;
            if (RES) {
                // ---- the mailbox: wait for the command of this step (or for the word that ends the launch)
                const unsigned want = p.res_seq0 + (unsigned)t + 1u;
                unsigned long long c = 0;
                if (lane == 0) {
                    c = blockIdx.x == 0 ? res_dispatch(p, want, res_gen, res_valid) : res_wait(p, want);
#ifdef NCG_RES_TIMELINE2
                    if (blockIdx.x == NCG_RES_TL_CTA) atomicAdd(p.res_done_ctr + 4 + 0, res_ns() - *(volatile unsigned long long*)(p.res_done_ctr + 1));
#endif
                }
                c = __shfl_sync(0xffffffffu, c, 0);
                if (((c >> 8) & 0xffu) != 0u) {
                    // leave: take the ray warps' outstanding "drained" arrivals, then wake them on the pose barrier with the exit flag up
                    for (int s = t - NB + 1 > 0 ? t - NB + 1 : 0; s < t; ++s) bar_sync(BAR_EMPTY + s % NB, NT);
                    if (lane == 0) s_res_exit = 1;
                    __syncwarp();
                    __threadfence_block();
                    bar_arrive(BAR_POSE + b, NT);
                    steps_done = t;
                    break;
                }
#ifdef NCG_RES_TIMELINE
                if (blockIdx.x == NCG_RES_TL_CTA && lane == 0) { long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); s_res_t0 = c_; }
                __syncwarp();
#endif
                // the caller's actions (mapped host memory, written before the command) and the result slot's pointers
                // (ld.global.cv: fetched again from host memory on every step, never from a cache line of the step before)
                if (active) {
                    if (p.discrete) act_d = __ldcv((const int*)p.actions + gc); else act_c = __ldcv((const float2*)p.actions + gc);
                }
#ifdef NCG_RES_TIMELINE
                if (active && act_c.x == 12345.0f && act_d == 77) s_res_exit = 2;      // (a use of the action before the stamp)
                RTL(15);
#endif
                const unsigned long long* row = p.res_dev_tab + ((unsigned)c & (NCG_RES_SLOTS - 1)) * 4;
                const unsigned long long q0 = __ldcg(row), q1 = __ldcg(row + 1), q2 = __ldcg(row + 2), q3 = __ldcg(row + 3);
                if (lane == 0) s_res_obs[b] = reinterpret_cast<float*>(q0);
                rew_out = reinterpret_cast<float*>(q1); term_out = reinterpret_cast<uint8_t*>(q2); trunc_out = reinterpret_cast<uint8_t*>(q3);
                __syncwarp();       // the lanes go through the step together (lane 0 has been on its own in the mailbox)
            }
            float rew = 0.0f;
            StepCtx ctx;
            if (active) {
                float thr, brk, st;
                if (RES) { if (p.discrete) action_discrete(act_d, &thr, &brk, &st); else action_continuous(act_c.x, act_c.y, &thr, &brk, &st); }
                else if (!synth) {
                    // (t == 0: asked for at the top of the kernel; launches with caller actions are single steps)
                    if (p.discrete) action_discrete(t == 0 ? act_d : ((const int*)p.actions)[gc], &thr, &brk, &st);
                    else { const float2 a = t == 0 ? act_c : ((const float2*)p.actions)[gc]; action_continuous(a.x, a.y, &thr, &brk, &st); }
                } else { const float4 a = s_act[b * SLOTS + slot]; thr = a.x; brk = a.y; st = a.z; }
#ifdef NCG_RES_TIMELINE
                if (RES && thr == 12345.0f) s_res_exit = 2;      // (a use of the action before the stamp)
                RTL(0);
#endif
                // Car.velocity_history (car.py:384-386): the speed update_physics saw, i.e. before b2World.Step (info only)
                if (p.vel_hist) p.vel_hist[(size_t)gc * NCG_VEL_HISTORY + f2u(R[NCG_R_STEP]) % NCG_VEL_HISTORY] = make_float2(R[NCG_R_VX], R[NCG_R_VY]);
                if (!CC) { if (!(p.debug_skip & 2)) car_step_dynamics(R, T, thr, brk, st, p.contacts, &ctx, &cnt); }
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
#ifdef NCG_RES_TIMELINE
                if (RES && blockIdx.x == NCG_RES_TL_CTA) {      // per lane: when its own dynamics were done (slot 11: sum over lanes, 12: lanes counted, 13: lanes later than 1.5 x lane 0)
                    long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory");
                    const long long d_ = c_ - *(volatile long long*)&s_res_t0;
                    atomicAdd(p.res_done_ctr + 4 + 11, (unsigned long long)d_); atomicAdd(p.res_done_ctr + 4 + 12, 1ull);
                    const long long d0_ = __shfl_sync(__activemask(), d_, 0);
                    if (2 * d_ > 3 * d0_) atomicAdd(p.res_done_ctr + 4 + 13, 1ull);
                    if (f2u(R[NCG_R_NCONTACT]) & 255u) atomicAdd(p.res_done_ctr + 4 + 14, 1ull);
                }
#endif
                s_pose[b * SLOTS + slot] = make_float4(R[NCG_R_X], R[NCG_R_Y], R[NCG_R_ANGLE], 0.0f);
            }
            if (threadIdx.x == 0) s_ctr[b] = 32 * RW;           // ray queue: every ray lane starts on job = its index
            // the pose exists: let the ray warps start while this warp does the rest of the step
            RTL(9);
            __syncwarp();
            __threadfence_block();
            RTL(10);
            bar_arrive(BAR_POSE + b, NT);
            RTL(1);
            TL(2);
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
            TL(3);
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
                    else { if (term_out) term_out[ge] = te ? 1 : 0; if (trunc_out) trunc_out[ge] = tr ? 1 : 0; }
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
                        reset_on_track(R, p.blob, p.track_off, rtid, start_of(p.start, gc - ge * p.C, CC, p.grid_dx, p.grid_dy));
                        if (solo || lane == le * p.C) {
                            p.env_track[ge] = (int)rtid; *p.redrawn = 1;
                            if (CC) for (int w_ = 0; w_ < NCG_CC_STRIDE; ++w_) p.cc_pairs[(size_t)ge * NCG_CC_STRIDE + w_] = 0.0f;   // a new world
                        }
                    } else reset_in_place(R, T, start_of(p.start, gc - ge * p.C, CC, p.grid_dx, p.grid_dy));
                }
                // (bits 8..: the row of reset_obs to hand out: one per track, or per (track, car of the env) on a start grid)
                s_flag[b * SLOTS + slot] = (te ? 1u : 0u) | (tr ? 2u : 0u) | ((CC ? rtid * (uint32_t)p.C + (uint32_t)(gc - ge * p.C) : rtid) << 8);
            }
            __syncwarp();
            RTL(2);
            __threadfence_block();     // (RES: rewards and flags went to mapped host memory; the ray warps' signaller orders them, below)
            RTL(3);
            bar_arrive(BAR_FULL + b, NT);
            TL(4);
        }
        // ---- counters
        unsigned long long v[7] = {active ? (unsigned long long)steps_done : 0ull, episodes, cnt.laps, 0ull, cnt.contact_steps, cnt.toi_events, cnt.overflow};
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
            TL(0);
            TLWAIT(bar_sync(BAR_POSE + b, NT));
            TL(1);
            if (RES) {
                if (*(volatile int*)&s_res_exit) break;
                obs_out = *(float* volatile*)&s_res_obs[b];
                RTG(1);
            }
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
            TL(2);
            if (warp == PW) RTL(4);
            RTG(4);
            TLWAIT(bar_sync(BAR_FULL + b, NT));
            TL(3);
            __syncwarp();
            if (warp == PW) RTL(5);
            RTG(5);
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
            TL(5);
            if (RES) {
                // the step is complete for the host once every CTA's rows are on their way: the ray warps meet, one thread
                // orders the CTA's writes before its count, and the grid's last arrival raises the host's done word
                if (warp == PW) RTL(6);
                RTG(6);
                bar_sync(BAR_RES, 32 * RW);
                if (warp == PW) RTL(7);
                if (warp == PW && lane == 0) {
                    // (release at GPU scope, cumulative over the barrier above: the CTA's writes are ordered before its count; the
                    // grid's last arrival, which has observed every count, is the one thread that pays for the system-scope fence)
#ifdef NCG_RES_TIMELINE2
                    const unsigned long long g0_ = res_ns();
#endif
                    res_fence_gpu();
                    RTL(8);
#ifdef NCG_RES_TIMELINE2
                    const unsigned long long g1_ = res_ns();
#endif
                    const unsigned long long old = atomicAdd(p.res_done_ctr, 1ull);
                    if (old + 1ull == (unsigned long long)gridDim.x * (unsigned long long)(t + 1)) {
#ifdef NCG_RES_TIMELINE2
                        const unsigned long long g2_ = res_ns();
#endif
                        if (p.res_fence_gpu) res_fence_gpu(); else res_fence_sys();
                        p.res_host_done[0] = p.res_seq0 + (unsigned)t + 1u;
                        p.res_done_ctr[2] += res_ns() - *(volatile unsigned long long*)(p.res_done_ctr + 1); p.res_done_ctr[3] += 1ull;   // (diagnostic: command seen -> done raised)
#ifdef NCG_RES_TIMELINE2
                        // the grid's last arrival: when its ray warps had met, its GPU-scope fence was done, it had counted itself
                        { const unsigned long long ts_ = *(volatile unsigned long long*)(p.res_done_ctr + 1);
                          p.res_done_ctr[4 + 8] += g0_ - ts_; p.res_done_ctr[4 + 9] += g1_ - ts_; p.res_done_ctr[4 + 10] += g2_ - ts_; }
#endif
                    }
                }
            }
            if (t + NB < p.T) { __threadfence_block(); bar_arrive(BAR_EMPTY + b, NT); }
            TL(4);
        }
        ray_tests = tests;
        for (int o = 16; o > 0; o >>= 1) ray_tests += __shfl_down_sync(0xffffffffu, ray_tests, o);
        if (lane == 0 && ray_tests) atomicAdd(((unsigned long long*)p.stats) + 3, ray_tests);
    }
#ifdef NCG_TIMELINE
    if (lane == 0 && warp < 3 && blockIdx.x < 4096) g_cta_cycles[blockIdx.x][1 + warp] = tl_wait;
    if (threadIdx.x == 0 && blockIdx.x < 4096) { g_cta_cycles[blockIdx.x][0] = tl_clock() - tl_k0; g_cta_cycles[blockIdx.x][4] = tl_k0; }
#endif
    // ---- records shared -> HBM
    __syncthreads();
#ifdef NCG_TIMELINE
    if (threadIdx.x == 0 && blockIdx.x == 5) g_timeline[11][127][0] = tl_clock();
#endif
    for (int i = threadIdx.x; i < n_all * (NCG_RECORD_WORDS / 4); i += NT) {
        const int ci = i >> 5;
        const float* d = s_rec + SLOT_OF(ci) * REC_STRIDE + (i & 31) * 4;
        reinterpret_cast<float4*>(p.records + (size_t)GCAR_OF(ci) * NCG_RECORD_WORDS)[i & 31] = make_float4(d[0], d[1], d[2], d[3]);
    }
#ifdef NCG_TIMELINE
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(&g_launch_ns[p.step_base & 4095][1], tl_ns());
#endif
#undef SLOT_OF
#undef GCAR_OF
}

}  // namespace
