// ncg_defs.cuh -- build-mode macros and the numeric constants of the CarEnv stepping path.
// Constants are the reference's (/root/reference/src/constants/*.py, file:line per group); the host-side
// mirror is nascargymnasium_b200/constants.py and tests/test_host.py checks the two agree.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../include/ncg_b200.h"

#if defined(__CUDACC__)
#define NCG_HD __host__ __device__ __forceinline__
#define NCG_HDN __host__ __device__ __noinline__
#else
#define NCG_HD static inline
#define NCG_HDN static
#endif

// Checked build (-DNCG_CHECKED, `NCG_CHECKED=1 python -c "from nascargymnasium_b200 import engine; engine.build_library()"` ->
// libncg_b200_checked.so): every table / record / slot index of the step kernel is range-checked on the device and a violation
// prints its site and traps.  tests/test_gpu_checked.py runs the launch shapes through it.  (compute-sanitizer is closed on
// the measurement pool -- profiles/README.md -- so this is the memory-safety evidence that can be produced there.)
#if defined(NCG_CHECKED) && defined(__CUDA_ARCH__)
#include <stdio.h>
#define NCG_CHECK(cond, what) do { if (!(cond)) { printf("NCG_CHECK failed: %s (%s:%d) block %d thread %d\n", what, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define NCG_CHECK(cond, what) ((void)0)
#endif

namespace ncg {

// 16-byte vector load type: CUDA's float4 on the device build, a plain struct in the host test build
#if defined(__CUDACC__)
typedef float4 F4;
#else
struct F4 { float x, y, z, w; };
#endif

// bit casts between the float32 record words and u32 payloads
NCG_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
NCG_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// index of the lowest set bit (v != 0)
NCG_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// approximate division (MUFU.RCP + FMUL, <= 2 ulp) for the sensor rays only: their tolerance is 1e-3 normalised
NCG_HD float fdiv_fast(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
}

// ---- Box2D 2.3 settings (b2Settings.h) -------------------------------------------------------------
#define NCG_B2_PI 3.14159265359f
#define NCG_B2_EPS 1.1920928955078125e-7f
#define NCG_B2_MAXFLOAT 3.402823466e+38f
#define NCG_B2_LINEAR_SLOP 0.005f
#define NCG_B2_POLY_RADIUS (2.0f * NCG_B2_LINEAR_SLOP)
#define NCG_B2_AABB_EXT 0.1f
#define NCG_B2_AABB_MULT 2.0f
#define NCG_B2_VEL_THRESHOLD 1.0f
#define NCG_B2_BAUMGARTE 0.2f
#define NCG_B2_TOI_BAUMGARTE 0.75f
#define NCG_B2_MAX_LIN_CORR 0.2f
#define NCG_B2_MAX_TRANSLATION 2.0f
#define NCG_B2_MAX_ROTATION (0.5f * NCG_B2_PI)
#define NCG_B2_MAX_SUBSTEPS 8
#define NCG_B2_TIME_TO_SLEEP 0.5f
#define NCG_B2_LIN_SLEEP_TOL 0.01f
#define NCG_B2_ANG_SLEEP_TOL (2.0f / 180.0f * NCG_B2_PI)

// ---- car_specs.py:6-100 -----------------------------------------------------------------------------
#define NCG_CAR_MASS 1500.0f
#define NCG_CAR_HALF_LENGTH 2.521f          /* CAR_LENGTH 5.042 / 2 */
#define NCG_CAR_HALF_WIDTH 0.998f           /* CAR_WIDTH 1.996 / 2 */
#define NCG_CAR_WHEELBASE 2.794f
#define NCG_CAR_MOI 1837.86125f             /* CAR_MASS*(L^2+W^2)*0.5/12 */
#define NCG_CAR_MAX_TORQUE 820.0f
#define NCG_CAR_MAX_POWER 499619.0f         /* 670 hp * 745.7 */
#define NCG_CAR_MAX_SPEED 89.408f
#define NCG_DRAG_CONSTANT 0.581875f
#define NCG_WEIGHT 14715.0f                 /* CAR_MASS * 9.81 */
#define NCG_STATIC_TYRE_LOAD 3678.75f
#define NCG_CAR_FRICTION 0.7f
#define NCG_CAR_RESTITUTION 0.1f
// ---- physics.py:6-33 --------------------------------------------------------------------------------
#define NCG_WALL_FRICTION 0.333f
#define NCG_WALL_RESTITUTION 0.25f
#define NCG_DT (1.0f / 60.0f)
#define NCG_VEL_ITERS 6
#define NCG_POS_ITERS 4
// ---- rewards.py / environment.py / collision.py ---------------------------------------------------
#define NCG_STUCK_STEPS 600                 /* float64 sum of 600 x (1/60) first exceeds 10.0 */
#define NCG_STUCK_EXT_STEPS 900
#define NCG_TERMINATE_STEPS 3601            /* simulation_time > 60 */
#define NCG_TRUNCATE_STEPS 10800            /* simulation_time > 180 */
#define NCG_MIN_LAP_STEPS 600               /* lap time < 10 s is rejected */

// per-car result bits handed from the car phase to the env phase of one step
enum {
    NCG_X_DIS_PRE = 1, NCG_X_DIS_MID = 2, NCG_X_DIS_POST = 4, NCG_X_LAP_PRE = 8, NCG_X_LAP_MID = 16, NCG_X_COMPLETED = 32,
    NCG_X_JUST_DISABLED = 64, NCG_X_LOW_REWARD = 128
};

struct Counters { unsigned long long ray_tests, contact_steps, toi_events, overflow, laps; };

}  // namespace ncg
