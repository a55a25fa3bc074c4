// Shared device helpers for the generated robot headers: rigid frames and the reference's trig.
#pragma once
#include <cuda_runtime.h>

namespace b200mp {

struct Frame {
    float r[9];  // row-major rotation
    float t[3];
};

// One step of a robot's state check (generated table `tasks`, run by check_state_grid in validate.cuh):
//   kTaskFk    : bring the streamed frame to `frame`, park the bounding-sphere centres of links [a, b)
//   kTaskEnv   : link a against the environment (bounding sphere; its fine spheres when that touches something)
//   kTaskSelfA : self-collision pair (a, b): bounding spheres; when they touch, the frame of link a is saved ...
//   kTaskSelfB : ... and, with link b's frame, a's fine spheres (outer loop) are tested against b's
//   kTaskGroup : link a against the sphere of pair group b (every member's bounding sphere lies inside it); not touched by any lane:
//                the group's pair tasks, which follow, are skipped together
// (one 32-bit word: a single constant load per task)
struct RobotTask {
    unsigned kind : 3, frame : 5, a : 12, b : 12;
};
enum { kTaskFk = 0, kTaskEnv = 1, kTaskSelfA = 2, kTaskSelfB = 3, kTaskGroup = 4 };

__device__ __forceinline__ void frame_identity(Frame &F) {
    F.r[0] = 1.f; F.r[1] = 0.f; F.r[2] = 0.f;
    F.r[3] = 0.f; F.r[4] = 1.f; F.r[5] = 0.f;
    F.r[6] = 0.f; F.r[7] = 0.f; F.r[8] = 1.f;
    F.t[0] = 0.f; F.t[1] = 0.f; F.t[2] = 0.f;
}

// world position of a frame-local point
__device__ __forceinline__ void frame_apply(const Frame &F, float lx, float ly, float lz, float &x, float &y, float &z) {
    x = fmaf(F.r[0], lx, fmaf(F.r[1], ly, fmaf(F.r[2], lz, F.t[0])));
    y = fmaf(F.r[3], lx, fmaf(F.r[4], ly, fmaf(F.r[5], lz, F.t[1])));
    z = fmaf(F.r[6], lx, fmaf(F.r[7], ly, fmaf(F.r[8], lz, F.t[2])));
}

// The reference's vector sin: cephes polynomials with the octant index ROUNDED to nearest
// (external/vamp/src/impl/vamp/vector/avx.hh:455-549, _mm256_cvtps_epi32 at :203-215), which is what gives
// it its 4.2e-6 error profile. sinf()/__sinf() would yield different robot geometry.
__device__ __forceinline__ float vamp_sin(float x) {
    const unsigned sign_in = __float_as_uint(x) & 0x80000000u;
    x = fabsf(x);
    float y = x * 1.27323954473516f;
    int j = __float2int_rn(y);
    j = (j + 1) & ~1;
    y = __int2float_rn(j);
    const unsigned sign = sign_in ^ ((static_cast<unsigned>(j) & 4u) << 29);
    x = fmaf(y, -0.78515625f, x);
    x = fmaf(y, -2.4187564849853515625e-4f, x);
    x = fmaf(y, -3.77489497744594108e-8f, x);
    const float z = x * x;
    float yc = 2.443315711809948E-005f;
    yc = fmaf(yc, z, -1.388731625493765E-003f);
    yc = fmaf(yc, z, 4.166664568298827E-002f);
    yc = yc * z;
    yc = yc * z;
    yc = fmaf(z, -0.5f, yc);
    yc = yc + 1.0f;
    float ys = -1.9515295891E-4f;
    ys = fmaf(ys, z, 8.3321608736E-3f);
    ys = fmaf(ys, z, -1.6666654611E-1f);
    ys = ys * z;
    ys = fmaf(ys, x, x);
    const float r = (j & 2) == 0 ? ys : yc;
    return __uint_as_float(__float_as_uint(r) ^ sign);
}

// cos(x) = sin(wrap(x + pi/2)) with the reference's float constants (vector/interface.hh:451-459)
__device__ __forceinline__ float vamp_cos(float x) {
    constexpr float PI = 3.14159265359;
    const float v = x + static_cast<float>(PI / 2.);
    const float w = v - ((v >= PI) ? static_cast<float>(2 * PI) : 0.0f);
    return vamp_sin(w);
}

}  // namespace b200mp
