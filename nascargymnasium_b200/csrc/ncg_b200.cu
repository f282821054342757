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
#include <emmintrin.h>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "ncg_car.cuh"

using namespace ncg;

// ncg_b200_cc.cu: launches ncg_step_kernel<4, 2, 1, true>; returns NULL or the CUDA error string
extern "C" __attribute__((visibility("hidden"))) const char* ncg_cc_launch(const void* kparams, size_t kparams_bytes, int n_ctas, int smem_bytes, cudaStream_t stream);

// ncg_b200_res.cu: launches the resident variant ncg_step_kernel<RPL, MINB, 1, false, true>; NULL, or why it was not launched
extern "C" __attribute__((visibility("hidden"))) const char* ncg_res_launch(const void* kparams, size_t kparams_bytes, int rpl, int minb, int n_ctas,
                                                                            int num_sms, int smem_bytes, cudaStream_t stream);

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
// NVTX ranges around the launches of the C ABI (tracing), opt-in with NCG_NVTX=1 so that the default path pays nothing:
// `nsys`/`ncu --nvtx` then show ncg_reset / ncg_step / ncg_rollout next to the caller's own ranges.
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* name) { static const bool en = [] { const char* v = getenv("NCG_NVTX"); return v && atoi(v) != 0; }(); on = en; if (on) nvtxRangePushA(name); }
    ~NvtxRange() { if (on) nvtxRangePop(); }
};
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(NCG_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

}  // namespace
#include "ncg_step.cuh"
namespace {

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
    int2* d_cta_stage = nullptr; int2* d_pair_stage = nullptr;            // per group / pair: where its track's table is, bytes to stage
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
    // diagnostic overrides, read once at ncg_create (-1 = not set): a launch does not scan the environment
    int ov_phys_warps = -1, ov_no_stage = -1, ov_ray_queue = -1, ov_min_blocks = -1, ov_pair_rw8 = -1;
    long long launches = 0;
    // host-buffer path
    cudaStream_t stream = nullptr;
    void* d_actions = nullptr; float* d_obs = nullptr; float* d_final = nullptr; float* d_reward = nullptr; uint8_t* d_term = nullptr; uint8_t* d_trunc = nullptr;
    uint8_t* d_mask = nullptr; int* d_tid = nullptr;
    void* p_actions = nullptr; float* p_obs = nullptr; float* p_final = nullptr; float* p_reward = nullptr; uint8_t* p_flags = nullptr;
    void* d_pack = nullptr; void* p_pack = nullptr; size_t pack_bytes = 0;
    int* p_any_done = nullptr;                       // page-locked, device-mapped flag word of ncg_step_mapped
    // resident mode of ncg_step_mapped (ncg_b200_res.cu): the mailbox.  p_res: [0] command word, [16] {done seq, left-by-itself flag}
    // (words of 4 bytes, its own cache line), [32] the word the device relay starts from, [64..] the result-slot table
    int res_enabled = 1; bool res_running = false; unsigned res_seq = 0, res_gen = 0; int res_slots = 0;
    unsigned long long res_idle_ns = 1000000ull;
    int res_idle_exits = 0, res_backoff = 0, res_backoff_len = 256;     // a caller that is slower than the idle time gets per-step launches
    unsigned long long* p_res = nullptr; unsigned long long* d_res = nullptr;
    const void* res_fixed[4] = {nullptr, nullptr, nullptr, nullptr};   // what the running kernel was launched with: actions, final_obs, ep_return, ep_length
    const void* res_slot_ptrs[NCG_RES_SLOTS][4];
    long long res_launches = 0;
    unsigned long long res_wait_ns = 0, res_steps = 0, res_dev_ns = 0, res_dev_steps = 0;   // diagnostics (ncg_debug_resident)
    // a step posted by ncg_step_mapped_post and not yet waited for: 0 none, 1 in the resident kernel's mailbox, 2 a launched kernel
    int pend_kind = 0; unsigned pend_seq = 0; unsigned long long pend_cmd = 0; NcgMappedBuffers pend_buf;
    std::chrono::steady_clock::time_point pend_t0;
};

namespace {

StartPose start_pose(const NcgHandle* h) { StartPose sp; sp.x = h->cfg.start_x; sp.y = h->cfg.start_y; sp.a = h->cfg.start_angle; return sp; }

// Whole envs per CTA: at most CPB car slots, fewer when that spreads a small batch over all SMs (4096 single-car
// envs: 28 per CTA = 147 CTAs on 148 SMs instead of 128 CTAs of 32).  Beyond two CTAs per SM the groups are full: a CTA's
// step costs about the same with 24 cars as with 32 (measured in the dispersed steady state, tools/ab_group.sh: 20480 envs
// 1.15 G car-steps/s with groups of 32 vs 1.03 G with 28, 40960 envs 1.42 vs 1.26 G; at 8192 and fewer, spreading still wins
// by 2 %), so fewer CTAs win.
int envs_per_cta(int E, int C, int sms) {
    const int epb_max = CPB / C;                               // C <= 10 < CPB
    const int min_ctas = (E + epb_max - 1) / epb_max;
    if (sms <= 0) sms = 148;
    const int target = ((min_ctas + sms - 1) / sms) * sms;     // whole waves of one CTA per SM
    int epb = min_ctas > 2 * sms ? epb_max : (E + target - 1) / target;
    { const char* v = getenv("NCG_GROUP_ENVS"); if (v && atoi(v) > 0) epb = atoi(v); }       // A/B runs
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
    if (!getenv("NCG_GROUP_ENVS")) while (epb < epb_max && count_ctas(env_track, E, epb) > waves * sms) ++epb;
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
        cudaFree(h->d_cta_tab); h->d_cta_tab = nullptr; h->cap_ctas = 0; cudaFree(h->d_cta_stage); h->d_cta_stage = nullptr;
        CUDA_TRY(cudaMalloc(&h->d_cta_tab, tab.size() * sizeof(int2)));
        CUDA_TRY(cudaMalloc(&h->d_cta_stage, tab.size() * sizeof(int2)));
        h->cap_ctas = (int)tab.size();
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_cta_tab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    h->n_ctas = (int)tab.size();
    auto stage_of = [&](int g) { const int t = seq[tab[g].x]; return make_int2((int)h->h_track_off[t], (int)(h->h_stage_words[t] * 4u)); };
    std::vector<int2> stage(tab.size());
    for (size_t g = 0; g < tab.size(); ++g) stage[g] = stage_of((int)g);
    CUDA_TRY(cudaMemcpyAsync(h->d_cta_stage, stage.data(), stage.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    // neighbouring groups of one track, two by two (a group without such a neighbour stays alone)
    std::vector<int2> pairs;
    for (size_t g = 0; g < tab.size();) {
        const bool two = g + 1 < tab.size() && seq[tab[g + 1].x] == seq[tab[g].x];
        pairs.push_back(make_int2((int)g, two ? (int)g + 1 : -1));
        g += two ? 2 : 1;
    }
    if ((int)pairs.size() > h->cap_pairs) {
        cudaFree(h->d_pair_tab); h->d_pair_tab = nullptr; h->cap_pairs = 0; cudaFree(h->d_pair_stage); h->d_pair_stage = nullptr;
        CUDA_TRY(cudaMalloc(&h->d_pair_tab, pairs.size() * sizeof(int2)));
        CUDA_TRY(cudaMalloc(&h->d_pair_stage, pairs.size() * sizeof(int2)));
        h->cap_pairs = (int)pairs.size();
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_pair_tab, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    std::vector<int2> pstage(pairs.size());
    for (size_t q = 0; q < pairs.size(); ++q) pstage[q] = stage_of(pairs[q].x);
    CUDA_TRY(cudaMemcpyAsync(h->d_pair_stage, pstage.data(), pstage.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
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

#define NCG_RES_UNSUPPORTED 1000      /* launch_step(resident): this batch has no resident kernel; the caller launches per step */
#define NCG_RES_SKIP 1001             /* res_step: not this time (the caller steps slower than the kernel's idle time) */
int launch_step(NcgHandle* h, KParams& p, cudaStream_t s, bool resident = false) {
    NvtxRange nvtx_(resident ? "ncg_step_resident" : p.T == 1 ? "ncg_step" : "ncg_rollout");
    { int rc = follow_redraws(h, s); if (rc) return rc; }
    if (h->cta_dirty) { int rc = build_cta_table(h); if (rc) return rc; }
    p.redraw = (!resident && h->redraw && p.T == 1 && p.auto_reset && h->n_tracks > 1) ? 1 : 0;
    p.redraw_step = h->steps_taken; if (!resident) h->steps_taken += (unsigned)p.T;
    const int sms = h->num_sms > 0 ? h->num_sms : 148;
    // rays per lane: 2 (8 ray warps per CTA) while the batch is at most one CTA per SM and latency-bound, 4 (4 ray warps,
    // better lane balance and fewer instructions per car-step) beyond that; measured in profiles/.  Two things that were
    // measured and are worse: cutting the batch into smaller CTAs so that they fill whole waves of resident slots evenly
    // (a CTA's step time is set by the physics warp's dependent chain, so fewer cars per CTA only lowers the work per
    // chain), and a fourth resident CTA per SM (28 car slots, wall AABBs left in L2, 96 registers with spills).
    const int RPL = h->cfg.car_contacts ? 4 : h->rays_per_lane ? h->rays_per_lane : (h->n_ctas > sms ? 4 : 2);
    // physics warps per CTA: one CTA per group (three resident per SM), or one per pair of groups (two resident per SM =
    // four physics warps).  Measured per resident wave the pair shape is ~1.2x slower (12 ray warps per SM serve 128 cars
    // instead of 96), so it is chosen when it saves enough waves: 16384 envs run as 256 CTAs in one wave instead of 512 in
    // two (+23 %), 8192 ten-car envs in 5 waves instead of 7 (+30 %), the 65536-env track mix in 4 instead of 5 (+7 %).
    int PW = 1;
    if (h->n_ctas > 2 * sms) {
        const int waves1 = (h->n_ctas + 3 * sms - 1) / (3 * sms), waves2 = (h->n_pairs + 2 * sms - 1) / (2 * sms);
        // (single-car envs: the pair shape with eight ray warps; fitted over 12288..131072 envs, tools/ab_pw.sh)
        if ((h->cfg.cars_per_env == 1 ? 13 : 12) * waves2 < 10 * waves1) PW = 2;
    }
    if (h->ov_phys_warps == 1 || h->ov_phys_warps == 2 || h->ov_phys_warps == 4) PW = h->ov_phys_warps;
    const bool cc = h->cfg.car_contacts != 0;                // shared world: one shape (an env's cars sit in one physics warp)
    if (cc) PW = 1;
    if (PW == 4 && (h->cfg.cars_per_env != 1 || h->n_ctas > sms)) PW = 1;       // the spread shape: single-car envs, one CTA per SM
    p.cta_tab = h->d_cta_tab; p.pair_tab = h->d_pair_tab; p.cta_stage = h->d_cta_stage; p.pair_stage = h->d_pair_stage; p.slot_env = h->identity ? nullptr : h->d_slot_env;
    p.stage = h->ov_no_stage > 0 ? 0 : 1;
    unsigned mx = 0;
    if (p.stage) for (unsigned w : h->h_stage_words) mx = w > mx ? w : mx;
    size_t smem = 0;        // (set below, once the shape is known: the number of step buffers depends on it)
    // rays handed out from a per-CTA queue (longest first) instead of a fixed lane -> rays map: pays once the SM is
    // issue-bound, i.e. with three resident CTAs per SM (measured on B200, daytona: +22 % at 65536 envs, +9 % at 16384,
    // -5 % at 8192 and -7 % at 4096, where a step is bound by latency and the queue's claims and job set-up only add to it)
    p.queue = h->n_ctas > 2 * sms ? 1 : 0;
    if (h->ov_ray_queue >= 0) p.queue = h->ov_ray_queue ? 1 : 0;
    // resident CTAs per SM the register allocation allows: as many as the batch has use for, up to what shared memory
    // (~72 KB per CTA) admits; the 4-rays-per-lane shape (160 threads) fits three
    int minb = h->n_ctas <= sms ? 1 : (h->n_ctas <= 2 * sms || RPL != 4 ? 2 : 3);
    if (h->ov_min_blocks >= 1 && h->ov_min_blocks <= 3) minb = h->ov_min_blocks;
    if (minb == 3 && RPL != 4) minb = 2;
    if (PW == 2) minb = 2;
    if (PW == 4) minb = 1;
    if (cc) minb = 2;
    const int nb = minb == 1 ? 3 : 2;                 // step buffers: NB of the kernel
    smem = (size_t)smem_layout(mx, PW == 2 ? 64 : 32, nb).total * 4;
    if (p.stage && h->max_smem > 0 && smem > (size_t)h->max_smem) {      // a user track too large to stage: read it through L1/L2
        p.stage = 0;
        smem = (size_t)smem_layout(0, PW == 2 ? 64 : 32, nb).total * 4;
    }
    // the pair shape's ray warps: eight (320 threads x 96 registers) for single-car envs, six (256 x 128) otherwise.  At large
    // batches a step waits for the ray warps while the physics warps idle half of the time (tools/timeline_probe.py), so two more
    // ray warps per CTA pay for the physics warps' spills: +10 % at 16384 and 65536 single-car envs; ten ray warps at 80
    // registers lose it again (1.32 vs 1.41 G at 65536), and ten-car envs on talladega, whose physics warps do the env phase and
    // frequent resets on top, lose 18 % with eight (1.07 vs 1.30 G).  NCG_PAIR_RW8 = 0 | 1 overrides.
    int pair_rw8 = h->cfg.cars_per_env == 1 ? 1 : 0;
    if (h->ov_pair_rw8 >= 0) pair_rw8 = h->ov_pair_rw8 ? 1 : 0;
    void (*k)(KParams) = PW == 4 ? ncg_step_kernel<2, 1, 4> : PW == 2 ? (pair_rw8 ? ncg_step_kernel<2, 2, 2> : ncg_step_kernel<4, 2, 2>)
                       : minb == 1 ? (RPL == 4 ? ncg_step_kernel<4, 1, 1> : ncg_step_kernel<2, 1, 1>)
                       : minb == 2 ? (RPL == 4 ? ncg_step_kernel<4, 2, 1> : ncg_step_kernel<2, 2, 1>)
                                   : ncg_step_kernel<4, 3, 1>;
    if (resident) {                                   // the resident variant lives in its own translation unit (ncg_b200_res.cu)
        if (cc || PW != 1) return NCG_RES_UNSUPPORTED;
        const char* err = ncg_res_launch(&p, sizeof(p), RPL, minb, h->n_ctas, sms, (int)smem, s);
        if (err) { g_err = err; return NCG_RES_UNSUPPORTED; }
        ++h->launches; ++h->res_launches;
        return NCG_OK;
    }
    if (cc) {                                         // the shared-world variant lives in its own translation unit (ncg_b200_cc.cu)
        const char* err = ncg_cc_launch(&p, sizeof(p), h->n_ctas, (int)smem, s);
        if (err) return fail(NCG_E_CUDA, std::string("shared-world step kernel: ") + err);
        ++h->launches;
        return NCG_OK;
    }
    {   // the kernel's dynamic shared memory limit: a driver call, made only when a launch needs more than any before it
        // (the attribute belongs to the function, not to the handle: raised, never lowered; per device)
        static std::mutex mu; static std::map<std::pair<const void*, int>, int> limit;
        std::lock_guard<std::mutex> lock(mu);
        int& have = limit[std::make_pair((const void*)k, h->cfg.device)];
        if ((int)smem > have) { CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); have = (int)smem; }
    }
    {
        // single-step launches chain with programmatic dependent launch: the next launch's CTAs start as this one's leave their SMs
        // and run their prologue in front of griddepcontrol.wait (NCG_PDL=0 switches it off)
        static const int pdl_on = [] { const char* v = getenv("NCG_PDL"); return v ? atoi(v) != 0 : 1; }();
        p.pdl = (pdl_on && p.T == 1) ? 1 : 0;
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(PW == 2 ? h->n_pairs : h->n_ctas); cfg.blockDim = dim3(PW == 4 ? 32 * (4 + 8) : PW == 2 ? (pair_rw8 ? 320 : 256) : 32 * (1 + 16 / RPL));
        cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = p.pdl ? 1 : 0;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, k, p));
    }
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

// ---- resident mode of ncg_step_mapped ------------------------------------------------------------------------------------
// Every other entry point that looks at or changes device state ends the resident launch first (RES_STOP): the records live in
// the kernel's shared memory while it runs and return to HBM when it leaves.
int res_wait(NcgHandle* h, int32_t* any_done);
int res_stop(NcgHandle* h) {
    if (h->pend_kind == 1) { int rc = res_wait(h, nullptr); if (rc) return rc; }      // (a posted step completes before anything else touches the state)
    if (!h->res_running) return NCG_OK;
    volatile unsigned* done = reinterpret_cast<volatile unsigned*>(h->p_res + 16);
    const unsigned long long idle_word = ((unsigned long long)h->res_seq << 32) | ((unsigned long long)(h->res_gen & 0xffffu) << 16);
    if (!done[1]) *reinterpret_cast<volatile unsigned long long*>(h->p_res) = idle_word | (NCG_RES_OP_EXIT << 8);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    h->res_running = false; done[1] = 0;
    *reinterpret_cast<volatile unsigned long long*>(h->p_res) = idle_word;
    { unsigned long long d[2] = {0, 0}; if (e == cudaSuccess && cudaMemcpy(d, h->d_res + 18, 16, cudaMemcpyDeviceToHost) == cudaSuccess) { h->res_dev_ns += d[0]; h->res_dev_steps += d[1]; } }
    if (e != cudaSuccess) return fail(NCG_E_CUDA, std::string("resident step kernel: ") + cudaGetErrorString(e));
    return NCG_OK;
}
#define RES_STOP(h) do { if ((h)->res_running) { int rc_ = res_stop(h); if (rc_) return rc_; } } while (0)

KParams base_params(NcgHandle* h);
// launch the resident kernel for the caller's fixed buffers; NCG_RES_UNSUPPORTED = this batch has none
// first_cmd: the command the kernel finds in the mailbox when it starts (posted BEFORE the launch: under a tool that makes launches
// synchronous -- ncu, a sanitizer -- the launch only returns when the kernel has left, i.e. after it has taken that step and idled out)
int res_start(NcgHandle* h, const void* h_actions, float* h_final_obs, float* h_ep_return, int32_t* h_ep_length, unsigned long long first_cmd) {
    KParams p = base_params(h);
    p.actions = h_actions; p.final_obs = h_final_obs; p.ep_return = h_ep_return; p.ep_length = h_ep_length; p.any_done = h->p_any_done;
    p.T = 0x7fffffff;
    p.res_host_cmd = h->p_res; p.res_host_done = reinterpret_cast<volatile unsigned*>(h->p_res + 16); p.res_host_tab = h->p_res + 64;
    p.res_dev_cmd = h->d_res; p.res_done_ctr = h->d_res + 16; p.res_dev_tab = h->d_res + 64;
    p.res_seq0 = h->res_seq; p.res_idle_ns = h->res_idle_ns;
    { static const int fence_gpu = [] { const char* v = getenv("NCG_RESIDENT_FENCE"); return v && !strcmp(v, "gpu") ? 1 : 0; }(); p.res_fence_gpu = fence_gpu; }
    const unsigned long long idle_word = ((unsigned long long)h->res_seq << 32) | ((unsigned long long)(h->res_gen & 0xffffu) << 16);
    volatile unsigned* done = reinterpret_cast<volatile unsigned*>(h->p_res + 16);
    done[0] = h->res_seq; done[1] = 0;
    h->p_res[32] = idle_word;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);                            // actions and table rows before the command
    *reinterpret_cast<volatile unsigned long long*>(h->p_res) = first_cmd;
    CUDA_TRY(cudaMemcpyAsync(h->d_res, h->p_res + 32, 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_res + 16, 0, 32, h->stream));         // (d_res + 20..: NCG_RES_TIMELINE sums, kept across launches)
    int rc = launch_step(h, p, h->stream, true);
    if (rc) return rc;
    h->res_running = true;
    h->res_fixed[0] = h_actions; h->res_fixed[1] = h_final_obs; h->res_fixed[2] = h_ep_return; h->res_fixed[3] = h_ep_length;
    return NCG_OK;
}
// post one step into the mailbox (the kernel is started if it is not running); NCG_RES_UNSUPPORTED / NCG_RES_SKIP = not taken
// (the caller launches the step kernel instead).  res_wait spins on the done word.
int res_post(NcgHandle* h, const NcgMappedBuffers& B) {
    if (h->redraw || h->cfg.car_contacts) return NCG_RES_UNSUPPORTED;
    if (h->res_running && (h->res_fixed[0] != B.actions || h->res_fixed[1] != B.final_obs || h->res_fixed[2] != B.ep_return || h->res_fixed[3] != B.ep_length)) RES_STOP(h);
    volatile unsigned* done = reinterpret_cast<volatile unsigned*>(h->p_res + 16);
    if (h->res_running && done[1]) {
        // it left by itself: nobody stepped for res_idle_ns.  A caller that keeps doing that (a policy that takes longer than
        // the idle time per step) gains nothing from a resident kernel and its own GPU work would wait for the SMs: after three
        // idle exits in a row the next res_backoff_len steps are launched one by one, twice as many every time it happens again
        RES_STOP(h);
        if (++h->res_idle_exits >= 3) {
            h->res_idle_exits = 0; h->res_backoff = h->res_backoff_len;
            if (h->res_backoff_len < 65536) h->res_backoff_len *= 2;
            return NCG_RES_SKIP;
        }
    } else if (h->res_running) h->res_idle_exits = 0;
    // the result slot of this set of buffers (callers rotate a few blocks so that returned arrays stay valid)
    const void* want[4] = {B.obs, B.reward, B.terminated, B.truncated};
    int slot = -1;
    for (int i = 0; i < h->res_slots && slot < 0; ++i) if (!memcmp(h->res_slot_ptrs[i], want, sizeof(want))) slot = i;
    if (slot < 0) {
        if (h->res_slots == NCG_RES_SLOTS) { h->res_slots = 0; ++h->res_gen; }      // table full: a new generation, the device drops its copy
        slot = h->res_slots++;
        memcpy(h->res_slot_ptrs[slot], want, sizeof(want));
        for (int k = 0; k < 4; ++k) reinterpret_cast<volatile unsigned long long*>(h->p_res + 64)[slot * 4 + k] = (unsigned long long)(uintptr_t)want[k];
    }
    *(volatile int*)h->p_any_done = 0;
    h->pend_seq = h->res_seq + 1u;
    h->pend_cmd = ((unsigned long long)h->pend_seq << 32) | ((unsigned long long)(h->res_gen & 0xffffu) << 16) | (unsigned)slot;
    h->pend_buf = B; h->pend_t0 = std::chrono::steady_clock::now();
    if (!h->res_running) { int rc = res_start(h, B.actions, B.final_obs, B.ep_return, B.ep_length, h->pend_cmd); if (rc) return rc; }
    else {
        __atomic_thread_fence(__ATOMIC_SEQ_CST);                        // actions and table rows before the command
        *reinterpret_cast<volatile unsigned long long*>(h->p_res) = h->pend_cmd;
    }
    h->pend_kind = 1;
    return NCG_OK;
}
int res_wait(NcgHandle* h, int32_t* any_done) {
    volatile unsigned* done = reinterpret_cast<volatile unsigned*>(h->p_res + 16);
    const unsigned seq = h->pend_seq;
    h->pend_kind = 0;
    for (;;) {
        bool left = false;
        for (unsigned spins = 1;; ++spins) {
            if (done[0] == seq) break;
            if (done[1]) { left = true; break; }
            __builtin_ia32_pause();
            if ((spins & 0xfffu) == 0u) {
                const cudaError_t q = cudaStreamQuery(h->stream);
                if (q != cudaErrorNotReady) {
                    if (done[0] == seq) break;
                    if (q != cudaSuccess) { h->res_running = false; return fail(NCG_E_CUDA, std::string("resident step kernel: ") + cudaGetErrorString(q)); }
                    left = true; break;                                 // the kernel has ended without answering: start it again
                }
                if (std::chrono::steady_clock::now() - h->pend_t0 > std::chrono::seconds(20)) {
                    *reinterpret_cast<volatile unsigned long long*>(h->p_res) = h->pend_cmd | (NCG_RES_OP_EXIT << 8);
                    h->res_enabled = 0;
                    return fail(NCG_E_CUDA, "resident step kernel did not answer within 20 s");
                }
            }
        }
        if (!left) break;
        // it left by itself between two looks at the mailbox (idle time-out): the command is still there for a new launch
        { cudaError_t e = cudaStreamSynchronize(h->stream); h->res_running = false; done[1] = 0;
          { unsigned long long d[2] = {0, 0}; if (e == cudaSuccess && cudaMemcpy(d, h->d_res + 18, 16, cudaMemcpyDeviceToHost) == cudaSuccess) { h->res_dev_ns += d[0]; h->res_dev_steps += d[1]; } }
          if (e != cudaSuccess) return fail(NCG_E_CUDA, std::string("resident step kernel: ") + cudaGetErrorString(e)); }
        if (done[0] == seq) break;
        { const NcgMappedBuffers& B = h->pend_buf; int rc = res_start(h, B.actions, B.final_obs, B.ep_return, B.ep_length, h->pend_cmd); if (rc) return rc == NCG_RES_UNSUPPORTED ? fail(NCG_E_CUDA, "resident step kernel could not be started again") : rc; }
    }
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    h->res_wait_ns += (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - h->pend_t0).count(); ++h->res_steps;
    h->res_seq = seq;
    if (any_done) *any_done = *(volatile int*)h->p_any_done;
    return NCG_OK;
}

}  // namespace

extern "C" {

const char* ncg_last_error(void) { return g_err.c_str(); }
#ifdef NCG_TIMELINE
int ncg_debug_timeline(long long* out) { return cudaMemcpyFromSymbol(out, g_timeline, sizeof(g_timeline)) == cudaSuccess ? 0 : 1; }
int ncg_debug_launch_ns(unsigned long long* out, int reset) { if (reset) { static unsigned long long init[4096][2]; for (int i = 0; i < 4096; ++i) { init[i][0] = ~0ull; init[i][1] = 0; } return cudaMemcpyToSymbol(g_launch_ns, init, sizeof(init)) == cudaSuccess ? 0 : 1; } return cudaMemcpyFromSymbol(out, g_launch_ns, sizeof(g_launch_ns)) == cudaSuccess ? 0 : 1; }
int ncg_debug_cta_cycles(long long* out) { return cudaMemcpyFromSymbol(out, g_cta_cycles, sizeof(g_cta_cycles)) == cudaSuccess ? 0 : 1; }
#endif
int ncg_version(void) { return 1; }
// diagnostics of the resident mode: {ns ncg_step_mapped spent from entry to the done word, steps taken through the mailbox,
// ns on the device from "command seen by CTA 0" to "done word raised", steps counted there (updated when a resident launch ends)}
#if defined(NCG_RES_TIMELINE) || defined(NCG_RES_TIMELINE2)
int ncg_debug_resident_timeline(NcgHandle* h, unsigned long long* out16) {
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    CUDA_TRY(cudaMemcpy(out16, h->d_res + 20, 16 * 8, cudaMemcpyDeviceToHost));
    return NCG_OK;
}
#endif
int ncg_resident_pause(NcgHandle* h) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    return NCG_OK;
}
int ncg_debug_resident(NcgHandle* h, unsigned long long* out4) {
    if (!h || !out4) return fail(NCG_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    out4[0] = h->res_wait_ns; out4[1] = h->res_steps; out4[2] = h->res_dev_ns; out4[3] = h->res_dev_steps;
    return NCG_OK;
}

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
    { auto ov = [](const char* name) { const char* v = getenv(name); return v ? atoi(v) : -1; };
      h->ov_phys_warps = ov("NCG_PHYS_WARPS"); h->ov_no_stage = ov("NCG_NO_STAGE"); h->ov_ray_queue = ov("NCG_RAY_QUEUE");
      h->ov_min_blocks = ov("NCG_MIN_BLOCKS"); h->ov_pair_rw8 = ov("NCG_PAIR_RW8"); }
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
    CUDA_TRY(cudaHostAlloc((void**)&h->p_res, 1024, cudaHostAllocMapped | cudaHostAllocPortable)); memset(h->p_res, 0, 1024);
    CUDA_TRY(cudaMalloc(&h->d_res, 1024)); CUDA_TRY(cudaMemset(h->d_res, 0, 1024));
    { const char* v = getenv("NCG_RESIDENT"); if (v) h->res_enabled = atoi(v) != 0;
      const char* u = getenv("NCG_RESIDENT_IDLE_US"); if (u && atoll(u) > 0) h->res_idle_ns = (unsigned long long)atoll(u) * 1000ull;
      // a launch shape forced for an A/B run is a request for the per-step launches
      if (h->ov_phys_warps >= 0 || h->ov_ray_queue >= 0 || h->ov_min_blocks >= 0 || h->ov_no_stage >= 0 || h->rays_per_lane) h->res_enabled = v && atoi(v) != 0; }
    CUDA_TRY(cudaMalloc(&h->d_env_track, E * sizeof(int)));
    CUDA_TRY(cudaMemset(h->d_env_track, 0, E * sizeof(int)));
    h->h_env_track.assign(E, 0);
    *out = h;
    return NCG_OK;
}

int ncg_destroy(NcgHandle* h) {
    if (!h) return NCG_OK;
    cudaSetDevice(h->cfg.device);
    if (h->res_running) res_stop(h);
    cudaFreeHost(h->p_res); cudaFree(h->d_res);
    cudaFree(h->d_cc_pairs); cudaFree(h->d_cc_worlds); cudaFree(h->d_vel_hist); cudaFree(h->d_records); cudaFree(h->d_blob); cudaFree(h->d_track_off); cudaFree(h->d_stats); cudaFree(h->d_reset_obs);
    cudaFree(h->d_actions); cudaFree(h->d_pack); cudaFree(h->d_final);
    cudaFree(h->d_mask); cudaFree(h->d_tid); cudaFree(h->d_cta_tab); cudaFree(h->d_pair_tab); cudaFree(h->d_cta_stage); cudaFree(h->d_pair_stage); cudaFree(h->d_slot_env); cudaFree(h->d_env_track);
    cudaFreeHost(h->p_redrawn);
    cudaFreeHost(h->p_actions); cudaFreeHost(h->p_pack); cudaFreeHost(h->p_final); cudaFreeHost(h->p_any_done);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return NCG_OK;
}

int ncg_upload_tracks(NcgHandle* h, const float* h_blob, const int64_t* h_offsets, int32_t n_tracks) {
    if (!h || !h_blob || !h_offsets || n_tracks < 1) return fail(NCG_E_INVALID, "bad track upload");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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
    h->cta_dirty = true;                                                      // the plan carries table offsets
    const int rc = h->cfg.car_contacts ? h->cfg.cars_per_env : 1;             // reset rows per track (a start grid: one per car of an env)
    CUDA_TRY(cudaMalloc(&h->d_reset_obs, (size_t)n_tracks * rc * NCG_OBS_DIM * 4));
    ncg_reset_obs_kernel<<<n_tracks * rc, 32>>>(h->d_blob, h->d_track_off, h->d_reset_obs, start_pose(h), rc, h->cfg.grid_dx, h->cfg.grid_dy);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    ++h->launches;
    return NCG_OK;
}

int ncg_reset(NcgHandle* h, const uint8_t* d_env_mask, const int32_t* d_track_id, int32_t fresh, float* d_obs, void* stream) {
    NvtxRange nvtx_("ncg_reset");
    if (!h) return fail(NCG_E_INVALID, "null handle");
    if (!h->d_blob) return fail(NCG_E_STATE, "ncg_upload_tracks must be called before ncg_reset");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    KParams p = base_params(h);
    p.actions = d_actions; p.obs = d_obs; p.reward = d_reward; p.term = d_terminated; p.trunc = d_truncated; p.final_obs = d_final_obs;
    p.ep_return = h->d_ep_return; p.ep_length = h->d_ep_length; p.any_done = h->d_ep_any;
    return launch_step(h, p, (cudaStream_t)stream);
}

int ncg_rollout(NcgHandle* h, int32_t steps, uint64_t seed, int32_t mode, float* d_obs_rollout, float* d_reward_rollout,
                uint8_t* d_done_rollout, float* d_obs_last, void* stream) {
    if (!h || steps < 1) return fail(NCG_E_INVALID, "bad rollout arguments");
    if (!h->was_reset) return fail(NCG_E_STATE, "Environment not properly initialized. Call reset() first.");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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

// One step with every buffer in page-locked, device-mapped host memory: the kernel reads the actions and writes observations /
// rewards / flags over PCIe itself.  Posting and waiting are separate calls so that a binding can do its own per-step
// bookkeeping while the GPU works; ncg_step_mapped / ncg_step_mapped_from are post + wait.
int ncg_step_mapped_post(NcgHandle* h, const void* src_actions, int32_t validate, const NcgMappedBuffers* b) {
    if (!h || !b || !b->actions || !b->obs || !b->reward || !b->terminated || !b->truncated) return fail(NCG_E_INVALID, "null argument");
    if (!h->was_reset) return fail(NCG_E_STATE, "Environment not properly initialized. Call reset() first.");
    if (h->pend_kind) return fail(NCG_E_STATE, "a posted step has not been waited for");
    if (src_actions) {
        // the caller's own action array, staged into the mapped buffer and range-checked in the same pass (CarEnv.step asserts
        // action_space.contains(action), /root/reference/src/car_env.py:694): nothing is stepped when the check fails
        // Streaming (non-temporal) stores: the GPU reads these lines across PCIe a microsecond later, and lines left dirty in this
        // core's cache have to be snooped out of it one by one (measured: the physics warps' action fetch 5.9 us instead of 1.5).
        const size_t N = (size_t)h->N;
        const bool aligned = (reinterpret_cast<uintptr_t>(b->actions) & 15u) == 0;
        if (h->cfg.discrete) {
            const int32_t* s = static_cast<const int32_t*>(src_actions); int32_t* d = static_cast<int32_t*>(b->actions);
            uint32_t bad = 0; size_t i = 0;
            if (aligned) {
                __m128i acc = _mm_setzero_si128(); const __m128i four = _mm_set1_epi32(4), zero = _mm_setzero_si128();
                for (; i + 4 <= N; i += 4) {
                    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
                    acc = _mm_or_si128(acc, _mm_or_si128(_mm_cmpgt_epi32(v, four), _mm_cmplt_epi32(v, zero)));
                    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), v);
                }
                bad |= (uint32_t)_mm_movemask_epi8(acc);
            }
            for (; i < N; ++i) { const int32_t v = s[i]; bad |= (uint32_t)v > 4u; d[i] = v; }
            _mm_sfence();
            if (validate && bad) return fail(NCG_E_INVALID, "Invalid action");
        } else {
            const float* s = static_cast<const float*>(src_actions); float* d = static_cast<float*>(b->actions);
            int ok = 1; size_t i = 0;
            if (aligned) {
                const __m128 lo = _mm_set1_ps(-1.0f), hi = _mm_set1_ps(1.0f);
                __m128 acc = _mm_cmpeq_ps(lo, lo);
                for (; i + 4 <= 2 * N; i += 4) {
                    const __m128 v = _mm_loadu_ps(s + i);
                    acc = _mm_and_ps(acc, _mm_and_ps(_mm_cmpge_ps(v, lo), _mm_cmple_ps(v, hi)));      // (a NaN fails both comparisons)
                    _mm_stream_ps(d + i, v);
                }
                ok &= _mm_movemask_ps(acc) == 0xF;
            }
            for (; i < 2 * N; ++i) { const float v = s[i]; ok &= (v >= -1.0f) & (v <= 1.0f); d[i] = v; }
            _mm_sfence();
            if (validate && !ok) return fail(NCG_E_INVALID, "Invalid action");
        }
    }
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (h->res_enabled && h->res_backoff > 0) --h->res_backoff;
    else if (h->res_enabled) {
        // host-driven loop: the resident kernel takes the step through its mailbox (no launch, no staging, no record traffic)
        const int rc = res_post(h, *b);
        if (rc != NCG_RES_UNSUPPORTED && rc != NCG_RES_SKIP) return rc;
        if (rc == NCG_RES_UNSUPPORTED) h->res_enabled = 0;     // this batch has no resident kernel: per-step launches from here on
    }
    RES_STOP(h);
    KParams p = base_params(h);
    p.actions = b->actions; p.obs = b->obs; p.reward = b->reward; p.term = b->terminated; p.trunc = b->truncated; p.final_obs = b->final_obs;
    p.ep_return = b->ep_return; p.ep_length = b->ep_length; p.any_done = h->p_any_done;
    *h->p_any_done = 0;
    int rc = launch_step(h, p, h->stream);
    if (rc) return rc;
    h->pend_kind = 2;
    return NCG_OK;
}
int ncg_step_mapped_wait(NcgHandle* h, int32_t* any_done) {
    if (!h) return fail(NCG_E_INVALID, "null handle");
    if (!h->pend_kind) return fail(NCG_E_STATE, "no step has been posted");
    if (h->pend_kind == 1) return res_wait(h, any_done);
    h->pend_kind = 0;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (any_done) *any_done = *h->p_any_done;
    return NCG_OK;
}
int ncg_step_mapped(NcgHandle* h, const void* h_actions, float* h_obs, float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated,
                    float* h_final_obs, float* h_ep_return, int32_t* h_ep_length, int32_t* any_done) {
    const NcgMappedBuffers b = {const_cast<void*>(h_actions), h_obs, h_reward, h_terminated, h_truncated, h_final_obs, h_ep_return, h_ep_length};
    const int rc = ncg_step_mapped_post(h, nullptr, 0, &b);
    return rc ? rc : ncg_step_mapped_wait(h, any_done);
}
int ncg_step_mapped_from(NcgHandle* h, const void* src_actions, int32_t validate, void* h_actions, float* h_obs, float* h_reward, uint8_t* h_terminated,
                         uint8_t* h_truncated, float* h_final_obs, float* h_ep_return, int32_t* h_ep_length, int32_t* any_done) {
    if (!src_actions) return fail(NCG_E_INVALID, "null argument");
    const NcgMappedBuffers b = {h_actions, h_obs, h_reward, h_terminated, h_truncated, h_final_obs, h_ep_return, h_ep_length};
    const int rc = ncg_step_mapped_post(h, src_actions, validate, &b);
    return rc ? rc : ncg_step_mapped_wait(h, any_done);
}

int ncg_get_state(NcgHandle* h, float* d_records, void* stream) {
    if (!h || !d_records) return fail(NCG_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    CUDA_TRY(cudaMemcpyAsync(d_records, h->d_records, (size_t)h->N * NCG_RECORD_WORDS * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return NCG_OK;
}
int ncg_set_state(NcgHandle* h, const float* d_records, void* stream) {
    if (!h || !d_records) return fail(NCG_E_INVALID, "null argument");
    if (!h->d_blob) return fail(NCG_E_STATE, "ncg_upload_tracks must be called first");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h_records, h->d_records, (size_t)h->N * NCG_RECORD_WORDS * 4, cudaMemcpyDeviceToHost));
    return NCG_OK;
}
int ncg_set_state_host(NcgHandle* h, const float* h_records) {
    if (!h || !h_records) return fail(NCG_E_INVALID, "null argument");
    if (!h->d_blob) return fail(NCG_E_STATE, "ncg_upload_tracks must be called first");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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

static_assert(NCG_CAR_PAIR_WORDS == NCG_CC_STRIDE, "include/ncg_b200.h and csrc/ncg_car.cuh disagree on the pair table");
int ncg_get_car_pairs_host(NcgHandle* h, float* h_pairs) {
    if (!h || !h_pairs) return fail(NCG_E_INVALID, "null argument");
    if (!h->d_cc_pairs) return fail(NCG_E_STATE, "the car-car contact table exists only with NcgConfig.car_contacts = 1");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h_pairs, h->d_cc_pairs, (size_t)h->cfg.num_envs * NCG_CC_STRIDE * 4, cudaMemcpyDeviceToHost));
    return NCG_OK;
}
int ncg_set_car_pairs_host(NcgHandle* h, const float* h_pairs) {
    if (!h || !h_pairs) return fail(NCG_E_INVALID, "null argument");
    if (!h->d_cc_pairs) return fail(NCG_E_STATE, "the car-car contact table exists only with NcgConfig.car_contacts = 1");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h->d_cc_pairs, h_pairs, (size_t)h->cfg.num_envs * NCG_CC_STRIDE * 4, cudaMemcpyHostToDevice));
    return NCG_OK;
}

int ncg_get_velocity_history_host(NcgHandle* h, float* h_out) {
    if (!h || !h_out) return fail(NCG_E_INVALID, "null argument");
    if (!h->d_vel_hist) return fail(NCG_E_STATE, "the velocity history is kept only with NcgConfig.track_info = 1");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h_out, h->d_vel_hist, (size_t)h->N * NCG_VEL_HISTORY * sizeof(float2), cudaMemcpyDeviceToHost));
    return NCG_OK;
}

int ncg_read_stats(NcgHandle* h, NcgStats* out, int32_t reset) {
    if (!h || !out) return fail(NCG_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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
    CUDA_TRY(cudaSetDevice(h->cfg.device)); RES_STOP(h);
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
