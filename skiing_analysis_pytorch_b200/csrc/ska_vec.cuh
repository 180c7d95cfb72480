// ska_vec.cuh - one arithmetic vocabulary for `float` (one point per thread) and `F2` (two points
// per thread in lockstep).  On sm_100a the F2 overloads are the packed fp32 instructions
// (fma.rn.f32x2 / mul.f32x2 / add.f32x2 -> SASS FFMA2 / FMUL2 / FADD2): one issue slot does the
// work of two scalar FFMAs, each component rounded exactly like a scalar IEEE fma, so the packed
// path is bit-identical to the scalar one.  Camera constants enter as scalar operands (the SASS
// forms take a `.F32` broadcast register / uniform register, no splat instruction is issued).
// On the host (tests/hostemu) F2 is a plain pair and every op is two scalar ops.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define SKA_HD __host__ __device__ __forceinline__
#define SKA_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SKA_HD inline
#define SKA_HD_NOINLINE inline
#endif

namespace ska {

#if defined(__CUDACC__)
using F2 = float2;
#else
struct F2 {
  float x, y;
};
#endif
struct B2 {
  bool x, y;
};

SKA_HD F2 mk2(float x, float y) {
  F2 r;
  r.x = x;
  r.y = y;
  return r;
}

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  using Mask = bool;
  static constexpr int N = 1;
  static SKA_HD float splat(float s) { return s; }
  static SKA_HD bool mask(bool a, bool) { return a; }
};
template <>
struct Vec<F2> {
  using Mask = B2;
  static constexpr int N = 2;
  static SKA_HD F2 splat(float s) { return mk2(s, s); }
  static SKA_HD B2 mask(bool a, bool b) { return B2{a, b}; }
};

// ---- lane access (compile-time index)
template <int I>
SKA_HD float lane(float a) {
  return a;
}
template <int I>
SKA_HD float lane(F2 a) {
  return I == 0 ? a.x : a.y;
}
template <int I>
SKA_HD bool lane(bool a) {
  return a;
}
template <int I>
SKA_HD bool lane(B2 a) {
  return I == 0 ? a.x : a.y;
}

// ---- scalar
SKA_HD float vfma(float a, float b, float c) { return fmaf(a, b, c); }
SKA_HD float vmul(float a, float b) { return a * b; }
SKA_HD float vadd(float a, float b) { return a + b; }
SKA_HD float vsub(float a, float b) { return a - b; }
SKA_HD float vneg(float a) { return -a; }

// ---- packed
SKA_HD F2 vneg(F2 a) { return mk2(-a.x, -a.y); }
#if defined(__CUDA_ARCH__)
SKA_HD F2 vfma(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
SKA_HD F2 vmul(F2 a, F2 b) { return __fmul2_rn(a, b); }
SKA_HD F2 vadd(F2 a, F2 b) { return __fadd2_rn(a, b); }
#else
SKA_HD F2 vfma(F2 a, F2 b, F2 c) { return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
SKA_HD F2 vmul(F2 a, F2 b) { return mk2(a.x * b.x, a.y * b.y); }
SKA_HD F2 vadd(F2 a, F2 b) { return mk2(a.x + b.x, a.y + b.y); }
#endif
SKA_HD F2 vsub(F2 a, F2 b) { return vadd(a, vneg(b)); }
// mixed forms: a scalar operand is a broadcast (folded into the instruction's operand form)
SKA_HD F2 vfma(F2 a, float b, F2 c) { return vfma(a, mk2(b, b), c); }
SKA_HD F2 vfma(float a, F2 b, F2 c) { return vfma(mk2(a, a), b, c); }
SKA_HD F2 vfma(F2 a, F2 b, float c) { return vfma(a, b, mk2(c, c)); }
SKA_HD F2 vfma(F2 a, float b, float c) { return vfma(a, mk2(b, b), mk2(c, c)); }
SKA_HD F2 vfma(float a, F2 b, float c) { return vfma(mk2(a, a), b, mk2(c, c)); }
SKA_HD F2 vmul(F2 a, float b) { return vmul(a, mk2(b, b)); }
SKA_HD F2 vmul(float a, F2 b) { return vmul(mk2(a, a), b); }
SKA_HD F2 vadd(F2 a, float b) { return vadd(a, mk2(b, b)); }
SKA_HD F2 vadd(float a, F2 b) { return vadd(mk2(a, a), b); }
SKA_HD F2 vsub(F2 a, float b) { return vadd(a, mk2(-b, -b)); }
SKA_HD F2 vsub(float a, F2 b) { return vadd(mk2(a, a), vneg(b)); }

// ---- special function unit (scalar MUFU per component)
SKA_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU.RCP, <= 1 ulp
  return r;
#else
  return 1.0f / x;
#endif
}
SKA_HD float sqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU.SQRT (ftz: no denormal fix-up code)
  return r;
#else
  return sqrtf(x);
#endif
}
SKA_HD F2 rcp_fast(F2 a) { return mk2(rcp_fast(a.x), rcp_fast(a.y)); }
SKA_HD F2 sqrt_fast(F2 a) { return mk2(sqrt_fast(a.x), sqrt_fast(a.y)); }
SKA_HD float vabs(float a) { return fabsf(a); }
SKA_HD F2 vabs(F2 a) { return mk2(fabsf(a.x), fabsf(a.y)); }

// ---- comparisons and masks
SKA_HD bool vgt(float a, float b) { return a > b; }
SKA_HD bool vlt(float a, float b) { return a < b; }
SKA_HD bool vle(float a, float b) { return a <= b; }
SKA_HD B2 vgt(F2 a, F2 b) { return B2{a.x > b.x, a.y > b.y}; }
SKA_HD B2 vlt(F2 a, F2 b) { return B2{a.x < b.x, a.y < b.y}; }
SKA_HD B2 vle(F2 a, F2 b) { return B2{a.x <= b.x, a.y <= b.y}; }
SKA_HD B2 vgt(F2 a, float b) { return B2{a.x > b, a.y > b}; }
SKA_HD B2 vlt(F2 a, float b) { return B2{a.x < b, a.y < b}; }
SKA_HD B2 vle(F2 a, float b) { return B2{a.x <= b, a.y <= b}; }
SKA_HD bool mand(bool a, bool b) { return a & b; }
SKA_HD B2 mand(B2 a, B2 b) { return B2{(bool)(a.x & b.x), (bool)(a.y & b.y)}; }
SKA_HD bool mall(bool a) { return a; }
SKA_HD bool mall(B2 a) { return a.x & a.y; }
SKA_HD float vsel(bool m, float a, float b) { return m ? a : b; }
SKA_HD F2 vsel(B2 m, F2 a, F2 b) { return mk2(m.x ? a.x : b.x, m.y ? a.y : b.y); }

}  // namespace ska
