// hrl_math.cuh - small vector helpers, model constants and the counter RNG for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HRL_FULL_MASK 0xffffffffu

// -DHRL_BOUNDS=1 (libhrl_b200_chk.so, tests/test_gpu_bounds_build.py): device-side asserts on every computed index into the
// shared-memory row / impulse / candidate / staging buffers - compute-sanitizer is not available on the GPU pool
#ifdef HRL_BOUNDS
#include <assert.h>
#define HRL_CHECK(c) assert(c)
#else
#define HRL_CHECK(c) ((void)0)
#endif

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// MUFU approximations without the IEEE fix-up / denormal rescue sequences (1-2 ulp; the physics
// tolerances are 1e-3 m / 1e-2 rad/s and the measured CUDA-vs-oracle error stays ~1e-5)
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float norm(V3 a) {
  const float d = dot(a, a);
  return d * rsqrt_ftz(fmaxf(d, 1e-30f));  // |a| = d / sqrt(d); exact 0 for a = 0
}
// I x for an inertia tensor that is axisymmetric about unit axis z: Ix*1 + (Iz-Ix) z z^T
__device__ __forceinline__ V3 axisym(float ix, float iz, V3 z, V3 x) { return ix * x + ((iz - ix) * dot(z, x)) * z; }

// sum / broadcast inside the 4-lane group that owns one env (lanes 4e..4e+3 of a warp)
__device__ __forceinline__ float gsum(float v) {
  v += __shfl_xor_sync(HRL_FULL_MASK, v, 1);
  v += __shfl_xor_sync(HRL_FULL_MASK, v, 2);
  return v;
}

// ---- Ant model constants (reference: assets/ant.xml; derivation SURVEY.md App. C.1) --------
// masses = 1000 kg/m^3 x volume, inertias = Bullet compound-shape AABB rule [3P-MEM]
namespace ant {
constexpr float R_TORSO = 0.25f;                    // ant.xml:13
constexpr float R_CAPS = 0.08f;                     // ant.xml:16
constexpr float M_TORSO = 65.44984694978736f;
constexpr float I_TORSO = 2.7270769562411402f;
constexpr float M_SHORT = 7.831583314284915f;       // 0.2*sqrt(2) capsules (leg + aux links)
constexpr float IX_SHORT = 0.10128847753141823f;
constexpr float IZ_SHORT = 0.16916219958855416f;
constexpr float M_LONG = 13.51850726010076f;        // 0.4*sqrt(2) capsule (foot link)
constexpr float IX_LONG = 0.3821231385521815f;
constexpr float IZ_LONG = 0.706567312794599f;
// torso + its 4 rigidly attached leg capsules, about the torso origin (= composite COM)
constexpr float M_COMP = M_TORSO + 4.0f * M_SHORT;
constexpr float IX_COMP = I_TORSO + 4.0f * (IX_SHORT + M_SHORT * 0.01f);
constexpr float IZ_COMP = I_TORSO + 4.0f * (IZ_SHORT + M_SHORT * 0.02f);
constexpr float HIP_LO = -0.6981317007977318f, HIP_HI = 0.6981317007977318f;  // ant.xml:18
constexpr float ANK_LO = 0.5235987755982988f, ANK_HI = 1.7453292519943295f;   // ant.xml:21
constexpr float IS2 = 0.70710678118654752440f;
}  // namespace ant
namespace pointbot {
constexpr float MASS = 10.0f;  // assets/player_cube.xml:8
constexpr float HALF = 0.35f;
}  // namespace pointbot

// ---- Philox4x32-10, same addressing as oracle/hrl_oracle.c -----------------------------------
enum { STREAM_JOINT = 0, STREAM_ITEM = 1, STREAM_GOAL = 2, STREAM_FLAG = 3, STREAM_ITEM_RESET = 4, STREAM_FLAG_CLOSE = 5 };
#define HRL_MAX_PLACE_ATTEMPTS 16
#define HRL_MAX_CLOSE_ATTEMPTS 64  // create_close_target near a corner rejects ~3 of 4 draws
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void rng_u4(uint64_t seed, uint32_t env, uint32_t stream, uint32_t draw, uint32_t sub,
                                        float u[4]) {
  uint32_t o[4];
  philox4x32_10(draw, env, stream, sub, (uint32_t)seed, (uint32_t)(seed >> 32), o);
#pragma unroll
  for (int i = 0; i < 4; i++) u[i] = (float)(o[i] >> 8) * (1.0f / 16777216.0f);
}
