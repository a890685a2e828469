// gmpe_device.cuh -- small device-side building blocks of the GraphMPE step kernels:
// Real traits (fp32 production / fp64 strict parity), Philox4x32-10, exact fast
// division by launch-time constants, and the scalar math of the step path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gmpe.h"

namespace gmpe {

template <typename Real> struct RT;
template <> struct RT<float> {
  using V2 = float2;
  using V4 = float4;
  static __device__ __forceinline__ V2 mk2(float x, float y) { return make_float2(x, y); }
  static __device__ __forceinline__ V4 mk4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
};
template <> struct RT<double> {
  using V2 = double2;
  using V4 = double4;
  static __device__ __forceinline__ V2 mk2(double x, double y) { return make_double2(x, y); }
  static __device__ __forceinline__ V4 mk4(double x, double y, double z, double w) { return make_double4(x, y, z, w); }
};

// ---- exact division by a launch-time constant d (d <= 2^12, n < 2^20) -------------
struct FastDiv {
  uint32_t mul, d;
  __host__ static FastDiv make(uint32_t d) {
    FastDiv f;
    f.d = d ? d : 1;
    f.mul = (uint32_t)((0x100000000ull + f.d - 1) / f.d);  // ceil(2^32 / d)
    return f;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : __umulhi(n, mul); }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = div(n);
    r = n - q * d;
  }
};

// Division by a divisor that is either a compile-time constant C (> 0: the compiler emits the
// multiply-shift itself and folds it into the surrounding address arithmetic) or, for C == 0,
// a launch-time constant handled by FastDiv.
template <int C>
struct CDiv {
  FastDiv f;
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return C > 0 ? n / (uint32_t)(C > 0 ? C : 1) : f.div(n); }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    if (C > 0) { q = n / (uint32_t)(C > 0 ? C : 1); r = n - q * (uint32_t)C; }
    else f.divmod(n, q, r);
  }
};

// ---- Philox4x32-10, counter = (idx, episode, world_lo, world_hi), key = seed -------
// Same specification as the CPU oracle (oracle/gmpe_oracle.c) so free-running batches
// can be compared; the code is independent.
__device__ __forceinline__ double philox_uniform01(uint64_t seed, int64_t world, uint32_t episode, uint32_t idx) {
  uint32_t c0 = idx, c1 = episode, c2 = (uint32_t)world, c3 = (uint32_t)((uint64_t)world >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return ((double)(c0 >> 5) * 67108864.0 + (double)(c1 >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, int64_t world, uint32_t ep, uint32_t idx, double lo, double hi) {
  return lo + (hi - lo) * philox_uniform01(seed, world, ep, idx);
}

// Source of the uniform draws of one world's reset: Philox, or -- teacher forcing for reset
// parity against the reference's own np.random stream -- the caller's buffer of uniform [0, 1)
// doubles (GmpeStepInject.reset_draws), consumed in the order reset_world calls np.random
// (july:368, :530, :467, core.py:328); NaN entries and indices past the buffer fall back to Philox.
struct DrawSrc {
  uint64_t seed;
  int64_t gw;      // global world id
  uint32_t ep;     // episode counter
  const double* inj;  // this world's injected draws or nullptr
  uint32_t inj_n;
  __device__ __forceinline__ double u01(uint32_t idx) const {
    if (inj && idx < inj_n) {
      const double u = inj[idx];
      if (u == u) return u;
    }
    return philox_uniform01(seed, gw, ep, idx);
  }
  __device__ __forceinline__ double uniform(uint32_t idx, double lo, double hi) const { return lo + (hi - lo) * u01(idx); }  // np.random.uniform
};

// Rejection sampling of spawn positions is bounded: the reference loops until a draw fits
// (july:452-486) and can be interrupted; a kernel cannot.  After this many rejected draws the
// last one is kept and the handle's overflow counter is raised (gmpe_placement_overflows).
#ifndef GMPE_PLACE_MAX_ATTEMPTS
#define GMPE_PLACE_MAX_ATTEMPTS 1024
#endif

// Landmark l of L for GmpeConfig.formation (july:492-497): 'point' = set_landmarks_in_point
// (utils.py:165-193: (0, -ws/3) rotated by the tube angle, from the tube exit); 'line' =
// set_landmarks_in_line (utils.py:77-131) from (-ws/2, -ws/2) to (ws/2, -ws/2) with np.linspace's
// arithmetic for a vector whose y-step is zero ((l / div) * delta + start, last = stop);
// 'circle' = set_landmarks_in_circle (utils.py:231-262), centre (0, exit_y + ws/5), radius ws/3.
__host__ __device__ __forceinline__ void formation_landmark(const GmpeConfig& c, int l, int L, double ext0, double ext1, double sa,
                                                            double ca, double& lx, double& ly) {
  const double ws = c.world_size;
  if (c.formation == GMPE_FORMATION_LINE) {
    const double start = -ws / 2, stop = ws / 2;
    lx = (L > 1 && l == L - 1) ? stop : (L > 1 ? ((double)l / (double)(L - 1)) * (stop - start) + start : start);
    ly = -ws / 2;
  } else if (c.formation == GMPE_FORMATION_CIRCLE) {
    const double ang = (double)l * (2 * 3.141592653589793 / (double)L);
    lx = 0.0 + (ws / 3) * cos(ang);
    ly = (ext1 + ws / 5) + (ws / 3) * sin(ang);
  } else {
    lx = ext0 + sa * (-ws / 3);
    ly = ext1 + ca * (-ws / 3);
  }
}

// ---- scalar math --------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
// fp32 production arithmetic.  IEEE sqrtf and the IEEE fp32 quotient cost 10-12 instructions
// each plus a ~35-instruction out-of-line slow path that is taken whenever an operand is
// zero or denormal (coincident landmarks, agents at rest, counters that have not moved) --
// on this path they were a quarter of all issued instructions (profiles/r1g_*).  The fp32
// instantiation therefore uses the SFU forms (sqrt.approx, div.approx: <= 2 ulp), which is
// three orders of magnitude inside the 1e-5 relative tolerance the path is specified to
// (BASELINE.json north_star); the fp64 strict-parity instantiation stays IEEE throughout.
__device__ __forceinline__ float sqrt_(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU.SQRT
  return r;
}
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float abs_(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_(double x) { return fabs(x); }
__device__ __forceinline__ void sincos_(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincos_(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ float asin_(float x) { return asinf(x); }
__device__ __forceinline__ double asin_(double x) { return asin(x); }
__device__ __forceinline__ float exp_(float x) { return expf(x); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
__device__ __forceinline__ double log1p_(double x) { return log1p(x); }
__device__ __forceinline__ bool isnan_(float x) { return isnan(x); }
__device__ __forceinline__ bool isnan_(double x) { return isnan(x); }

// quotient (see sqrt_)
__device__ __forceinline__ float div_(float a, float b) {  // MUFU.RCP + FMUL
  float r;
  asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ double div_(double a, double b) { return a / b; }
__device__ __forceinline__ float qdiv(float a, float b) { return div_(a, b); }
__device__ __forceinline__ double qdiv(double a, double b) { return a / b; }
// Quotient for the info means (july:800-823): the approximate reciprocal plus one residual
// correction (the residual is exact in an fma).  A mean of equal integers then comes out as
// that integer, so the std next to it is exactly 0 -- the reference's mean / (std + 1e-4)
// column turns a 1-ulp error in the mean into 0.2 % otherwise.
__device__ __forceinline__ float mean_div(float a, float b) {
  const float q = div_(a, b);
  return q + div_(fmaf(-q, b, a), b);
}
__device__ __forceinline__ double mean_div(double a, double b) { return a / b; }

__device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
// (x + pi) % (2 pi) - pi of the reference's heading errors (three_phase_graph.py:1058, 1106, 1222)
__device__ __forceinline__ float wrap_pi(float x) { return x - 6.283185307179586f * rintf(x * 0.15915494309189535f); }
__device__ __forceinline__ double wrap_pi(double x) {
  double m = fmod(x + 3.14159265358979323846, 6.283185307179586476925286766559);
  if (m < 0) m += 6.283185307179586476925286766559;
  return m - 3.14159265358979323846;
}
// np.hypot
__device__ __forceinline__ float hypot_(float x, float y) { return sqrt_(x * x + y * y); }
__device__ __forceinline__ double hypot_(double x, double y) { return hypot(x, y); }

// np.linalg.norm of a 2-vector; one definition for every distance so that adjacency,
// edges and collision tests see identical values.
template <typename Real>
__device__ __forceinline__ Real norm2(Real x, Real y) { return sqrt_(x * x + y * y); }

template <typename Real>
__device__ __forceinline__ Real clip(Real v, Real lo, Real hi) { return v < lo ? lo : (v > hi ? hi : v); }

// np.logaddexp(0, x) = max(x, 0) + log1p(exp(-|x|)); fp32: SFU exp / log, absolute error
// <= 1e-7 (the contact force it scales is O(1))
__device__ __forceinline__ double softplus(double x) { return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x)); }
__device__ __forceinline__ float softplus(float x) { return fmaxf(x, 0.0f) + __logf(1.0f + __expf(-fabsf(x))); }

// the reference casts positions to float32 before projecting on the tube frame (july:624)
__device__ __forceinline__ float trunc32(float x) { return x; }
__device__ __forceinline__ double trunc32(double x) { return (double)(float)x; }

// fp32 production keeps headings wrapped to [-pi, pi) (the reference never
// wraps theta, which costs fp32 ulps); fp64 strict parity leaves theta raw.
__device__ __forceinline__ float wrap_heading(float th) {
  const float two_pi = 6.283185307179586f, inv = 0.15915494309189535f;
  return th - two_pi * rintf(th * inv);
}
__device__ __forceinline__ double wrap_heading(double th) { return th; }

// One feature row of 8 Reals with 256-bit stores (sm_100: STG.E.ENL2.256): a lane writes
// whole 32-byte sectors, so per-world rows need no staging to reach the memory system as
// full-sector writes.  p must be 32-byte aligned.
__device__ __forceinline__ void store_row8(float* p, const float (&f)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]),
               "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7])
               : "memory");
}
__device__ __forceinline__ void store_row8(double* p, const double (&f)[8]) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(f[0]), "d"(f[1]), "d"(f[2]), "d"(f[3]) : "memory");
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "d"(f[4]), "d"(f[5]), "d"(f[6]), "d"(f[7]) : "memory");
}

}  // namespace gmpe
