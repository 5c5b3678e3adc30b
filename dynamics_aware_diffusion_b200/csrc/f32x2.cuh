// Packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2) and the fast Mish used by the bf16 epilogues.
#pragma once
#include <stdint.h>

namespace dad {

// ---- packed fp32x2 helpers (sm_100 FFMA2 / FADD2 / FMUL2) ---------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Mish on a pair: y * (1 - 2 / ((1 + e^y)^2 + 1))
__device__ __forceinline__ f32x2 mish2(f32x2 y) {
  float z0, z1;
  upk2(fmul2(y, pk2(1.4426950408889634f, 1.4426950408889634f)), z0, z1);
  const f32x2 one = pk2(1.f, 1.f);
  const f32x2 u = fadd2(pk2(ex2_approx(z0), ex2_approx(z1)), one);
  float w0, w1;
  upk2(ffma2(u, u, one), w0, w1);
  const f32x2 t = ffma2(pk2(rcp_approx(w0), rcp_approx(w1)), pk2(-2.f, -2.f), one);
  return fmul2(y, t);
}

}  // namespace dad
