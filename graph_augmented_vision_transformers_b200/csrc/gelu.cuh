// gelu.cuh - the GELU of nn.GELU (vit.py:84) and its derivative, shared by the streaming edges (edges.cu) and the fused
// fc1 GEMM epilogue (mlp_tc.cu).
#pragma once
#include "common.cuh"

namespace gvit {

// ---- GELU (exact erf form, nn.GELU at vit.py:84) fused with the dropout that follows it (vit.py:92) --------
// Exact form (erff) for fp32 storage - the parity path.  For bf16 storage (8-bit mantissa) the normal CDF comes from
// Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, one MUFU.EX2 + one MUFU.RCP + 7 FMA, branch-free): erff's two-branch
// polynomial made these kernels ALU-bound at 2.7x their HBM time.  The exp(-u^2/2) factor is shared with GELU'.
// packed fp32 pairs (FFMA2 / FMUL2 / FADD2 on sm_100): one issue slot per TWO lanes' worth of fp32 math.  These
// kernels are issue-bound (GELU + Philox + pack/unpack ~ 37 instructions per element against ~23 that fit under the
// HBM time), so the polynomial part runs two elements per instruction.
struct f32x2 { unsigned long long v; };
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(f32x2 x, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

template <bool EXACT> struct Gelu;
template <> struct Gelu<true> {
  static __device__ __forceinline__ float fwd(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f)); }
  static __device__ __forceinline__ float grad(float u) {
    return 0.5f * (1.0f + erff(u * 0.70710678118654752f)) + u * 0.3989422804014327f * __expf(-0.5f * u * u);
  }
  // a[t] = mul * gelu(a[t]);   g[t] = g[t] * mul * gelu'(u[t])
  static __device__ __forceinline__ void fwd8(float (&a)[8], float mul) {
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] = fwd(a[t]) * mul;
  }
  static __device__ __forceinline__ void grad8(float (&g)[8], const float (&u)[8], float mul) {
#pragma unroll
    for (int t = 0; t < 8; ++t) g[t] = g[t] * mul * grad(u[t]);
  }
  static __device__ __forceinline__ void fwd_grad8(float (&a)[8], float (&d)[8], float mul) {
#pragma unroll
    for (int t = 0; t < 8; ++t) { d[t] = mul * grad(a[t]); a[t] = fwd(a[t]) * mul; }
  }
};
template <> struct Gelu<false> {
  // returns Phi(u) and e = exp(-u^2/2); MUFU.RCP / MUFU.EX2 approximations (1 ulp-ish, far below bf16 resolution):
  // __frcp_rn compiled to a subroutine call with a divergent slow path
  static __device__ __forceinline__ float cdf(float u, float& e) {
    float t, ee;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(u), 1.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ee) : "f"(u * (u * -0.72134752044448170f)));   // exp(-u^2/2)
    e = ee;
    float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);    // coefficients pre-multiplied by 0.5
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    const float half_q = poly * t * ee;                // 0.5 * erfc(|u|/sqrt2)
    return 0.5f + copysignf(0.5f - half_q, u);         // u >= 0: 1 - half_q, u < 0: half_q
  }
  static __device__ __forceinline__ float fwd(float u) { float e; return u * cdf(u, e); }
  static __device__ __forceinline__ float grad(float u) { float e; const float c = cdf(u, e); return fmaf(u * 0.3989422804014327f, e, c); }

  // (mul * Phi(u0), mul * Phi(u1)) and (e0, e1) = exp(-u^2/2): the same evaluation as cdf(), two elements per instruction,
  // with `mul` (the dropout scale) folded into the polynomial coefficients.  MUFU.RCP / MUFU.EX2, |u| and copysign stay
  // scalar (no packed form).
  static __device__ __forceinline__ f32x2 cdf2(float u0, float u1, float mul, f32x2& e) {
    const float h = 0.5f * mul;
    float t0, t1, e0, e1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(u0), 1.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(u1), 1.0f)));
    const f32x2 u = pk2(u0, u1), t = pk2(t0, t1);
    float w0, w1;
    up2(mul2(mul2(u, pk2(-0.72134752044448170f, -0.72134752044448170f)), u), w0, w1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(w0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(w1));
    e = pk2(e0, e1);
    const float c5 = h * 1.061405429f, c4 = h * -1.453152027f, c3 = h * 1.421413741f, c2 = h * -0.284496736f, c1 = h * 0.254829592f;
    f32x2 poly = fma2(pk2(c5, c5), t, pk2(c4, c4));
    poly = fma2(poly, t, pk2(c3, c3));
    poly = fma2(poly, t, pk2(c2, c2));
    poly = fma2(poly, t, pk2(c1, c1));
    const f32x2 hq = mul2(mul2(poly, t), e);                         // mul * 0.5 * erfc(|u|/sqrt2)
    float r0, r1;
    up2(fma2(hq, pk2(-1.0f, -1.0f), pk2(h, h)), r0, r1);             // mul * (0.5 - half_q) >= 0
    return add2(pk2(h, h), pk2(copysignf(r0, u0), copysignf(r1, u1)));
  }
  static __device__ __forceinline__ void fwd8(float (&a)[8], float mul) {
#pragma unroll
    for (int t = 0; t < 8; t += 2) {
      f32x2 e;
      const f32x2 c = cdf2(a[t], a[t + 1], mul, e);
      up2(mul2(pk2(a[t], a[t + 1]), c), a[t], a[t + 1]);
    }
  }
  static __device__ __forceinline__ void grad8(float (&g)[8], const float (&u)[8], float mul) {
    const float k = 0.3989422804014327f * mul;
#pragma unroll
    for (int t = 0; t < 8; t += 2) {
      f32x2 e;
      const f32x2 c = cdf2(u[t], u[t + 1], mul, e);
      const f32x2 d = fma2(mul2(pk2(u[t], u[t + 1]), pk2(k, k)), e, c);     // mul * (Phi + u phi)
      up2(mul2(pk2(g[t], g[t + 1]), d), g[t], g[t + 1]);
    }
  }
  // a[t] = mul * gelu(a[t]) and d[t] = mul * gelu'(a[t]) from ONE evaluation of Phi and exp(-u^2/2): two packed instructions
  // more than fwd8.  The fused fc1 GEMM stores d (times the keep bit) instead of the pre-activation, so that the backward
  // GEMM's epilogue is a single multiply (mlp_tc.cu, `factor` mode).
  static __device__ __forceinline__ void fwd_grad8(float (&a)[8], float (&d)[8], float mul) {
    const float k = 0.3989422804014327f * mul;
#pragma unroll
    for (int t = 0; t < 8; t += 2) {
      f32x2 e;
      const f32x2 u = pk2(a[t], a[t + 1]);
      const f32x2 c = cdf2(a[t], a[t + 1], mul, e);
      up2(fma2(mul2(u, pk2(k, k)), e, c), d[t], d[t + 1]);
      up2(mul2(u, c), a[t], a[t + 1]);
    }
  }
};
template <typename T> struct GeluFor { using type = Gelu<true>; };
template <> struct GeluFor<__nv_bfloat16> { using type = Gelu<false>; };

}  // namespace gvit
