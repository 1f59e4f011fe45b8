// Adam + Polyak on one float4 of one member, shared by the optimizer kernel (kernels_simt.cu) and the weight-gradient
// epilogue of the chained backward (bwd_chain.cu).
//   torch.optim.Adam defaults (jsrl_utils.py:263-265), soft_update iql.py:72-74.
#pragma once
#include "common.cuh"
#include "engine.h"

namespace iql {

// bias_hidden scheme (engine.h): magnitude + / - half a TF32 ulp on the bit pattern; exact and reversible
__device__ __forceinline__ float tf32_bias(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ float tf32_unbias(float x) { return __uint_as_float(__float_as_uint(x) - 0x1000u); }
__device__ __forceinline__ bool in_hidden_weights(const StepCtx& ctx, int64_t i) {
  bool h = false;
  if (ctx.bias_hidden)
    for (int r = 0; r < ctx.n_hid; ++r) h |= (i >= ctx.hid_begin[r] && i < ctx.hid_end[r]);
  return h;
}

// One float4 of one member: Adam on p / m / v, TF32 shadow copies, Polyak on the target for the Q range.
__device__ __forceinline__ void adam_quad(const StepCtx& ctx, int m, int64_t i, const float4 g4, float4 p4, float4 m4,
                                          float4 v4, float4 t4, const AdamScalars as, float adam_w1, float adam_beta2,
                                          float adam_one_minus_b2, float adam_eps, float tau, float one_minus_tau,
                                          float* __restrict__ params, float* __restrict__ exp_avg,
                                          float* __restrict__ exp_avg_sq, float* __restrict__ target) {
  const int64_t off = m * ctx.P + i;
  const float g[4] = {g4.x, g4.y, g4.z, g4.w};
  const bool hidden = in_hidden_weights(ctx, i);
  float p[4] = {p4.x, p4.y, p4.z, p4.w};
  if (hidden) {
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = tf32_unbias(p[j]);
  }
  float mm[4] = {m4.x, m4.y, m4.z, m4.w};
  float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mm[j] = fmaf(adam_w1, g[j] - mm[j], mm[j]);                       // exp_avg.lerp_(grad, 1-beta1)
    vv[j] = __fmul_rn(vv[j], adam_beta2);                              // exp_avg_sq.mul_(beta2)
    vv[j] = fmaf(__fmul_rn(adam_one_minus_b2, g[j]), g[j], vv[j]);     // .addcmul_(g, g, 1-beta2)
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vv[j]), as.bc2_sqrt), adam_eps);
    p[j] = fmaf(as.neg_step_size, __fdiv_rn(mm[j], denom), p[j]);      // addcdiv_
  }
  if (hidden && !ctx.unbias_out)
    *reinterpret_cast<float4*>(params + off) = make_float4(tf32_bias(p[0]), tf32_bias(p[1]), tf32_bias(p[2]), tf32_bias(p[3]));
  else
    *reinterpret_cast<float4*>(params + off) = make_float4(p[0], p[1], p[2], p[3]);
  bool first_layer = false;
  if (ctx.tf32 && !hidden) {
    const float4 hi = make_float4(round_tf32(p[0]), round_tf32(p[1]), round_tf32(p[2]), round_tf32(p[3]));
    *reinterpret_cast<float4*>(ctx.w_shadow + off) = hi;
#pragma unroll
    for (int r = 0; r < 5; ++r) first_layer |= (i >= ctx.first_w_begin[r] && i < ctx.first_w_end[r]);
    if (first_layer)
      *reinterpret_cast<float4*>(ctx.w_shadow_lo + off) =
          make_float4(round_tf32(p[0] - hi.x), round_tf32(p[1] - hi.y), round_tf32(p[2] - hi.z), round_tf32(p[3] - hi.w));
  }
  *reinterpret_cast<float4*>(exp_avg + off) = make_float4(mm[0], mm[1], mm[2], mm[3]);
  *reinterpret_cast<float4*>(exp_avg_sq + off) = make_float4(vv[0], vv[1], vv[2], vv[3]);
  if (i < ctx.PQ) {
    // soft_update with the post-Adam Q: (1-tau)*target + tau*source, two products and a sum
    const int64_t toff = m * ctx.PQ + i;
    float t[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = __fadd_rn(__fmul_rn(one_minus_tau, hidden ? tf32_unbias(t[j]) : t[j]), __fmul_rn(tau, p[j]));
    if (hidden && !ctx.unbias_out)
      *reinterpret_cast<float4*>(target + toff) = make_float4(tf32_bias(t[0]), tf32_bias(t[1]), tf32_bias(t[2]), tf32_bias(t[3]));
    else
      *reinterpret_cast<float4*>(target + toff) = make_float4(t[0], t[1], t[2], t[3]);
    if (ctx.tf32 && !hidden) {
      const float4 hi = make_float4(round_tf32(t[0]), round_tf32(t[1]), round_tf32(t[2]), round_tf32(t[3]));
      *reinterpret_cast<float4*>(ctx.t_shadow + toff) = hi;
      if (first_layer)
        *reinterpret_cast<float4*>(ctx.t_shadow_lo + toff) =
            make_float4(round_tf32(t[0] - hi.x), round_tf32(t[1] - hi.y), round_tf32(t[2] - hi.z), round_tf32(t[3] - hi.w));
    }
  }
}


// Same update with approximate square root / reciprocal (MUFU, ~1 ulp) instead of IEEE sqrt + two IEEE divisions: ~20
// instead of ~130 instructions per element.  Used by the optimizer epilogue of the chained backward, i.e. only on the
// TF32 tensor-core path, whose GEMM operands carry 2^-11 rounding -- five orders of magnitude above the 2^-22 this
// changes in one update (the FP32 validation path keeps the exact torch formulas above).  inv_bc2 = 1 / bc2_sqrt.
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void adam_quad_fast(const StepCtx& ctx, int m, int64_t i, const float4 g4, float4 p4, float4 m4,
                                               float4 v4, float4 t4, float neg_step_size, float inv_bc2, float adam_w1,
                                               float adam_beta2, float adam_one_minus_b2, float adam_eps, float tau,
                                               float one_minus_tau, bool first_layer, float* __restrict__ params,
                                               float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                               float* __restrict__ target) {
  const int64_t off = m * ctx.P + i;
  const float g[4] = {g4.x, g4.y, g4.z, g4.w};
  const bool hidden = in_hidden_weights(ctx, i);
  float p[4] = {p4.x, p4.y, p4.z, p4.w};
  if (hidden) {
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = tf32_unbias(p[j]);
  }
  float mm[4] = {m4.x, m4.y, m4.z, m4.w};
  float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mm[j] = fmaf(adam_w1, g[j] - mm[j], mm[j]);
    vv[j] = fmaf(adam_one_minus_b2 * g[j], g[j], vv[j] * adam_beta2);
    const float denom = fmaf(sqrt_approx(vv[j]), inv_bc2, adam_eps);
    p[j] = fmaf(neg_step_size, mm[j] * rcp_approx(denom), p[j]);
  }
  // master weights, moments and the target are next read one step later, after ~1 GB of other traffic: streaming
  // stores (evict-first) keep them from displacing the operands the pair is about to reload from L2; the TF32 shadow
  // copies are what the next forward reads first and keep the default policy
  if (hidden && !ctx.unbias_out)  // read next by this pair's TMA loads (dgrad of the next step, forward): default policy
    *reinterpret_cast<float4*>(params + off) = make_float4(tf32_bias(p[0]), tf32_bias(p[1]), tf32_bias(p[2]), tf32_bias(p[3]));
  else
    __stcs(reinterpret_cast<float4*>(params + off), make_float4(p[0], p[1], p[2], p[3]));
  const float4 hi = make_float4(round_tf32(p[0]), round_tf32(p[1]), round_tf32(p[2]), round_tf32(p[3]));
  if (!hidden) *reinterpret_cast<float4*>(ctx.w_shadow + off) = hi;
  if (first_layer)
    *reinterpret_cast<float4*>(ctx.w_shadow_lo + off) =
        make_float4(round_tf32(p[0] - hi.x), round_tf32(p[1] - hi.y), round_tf32(p[2] - hi.z), round_tf32(p[3] - hi.w));
  __stcs(reinterpret_cast<float4*>(exp_avg + off), make_float4(mm[0], mm[1], mm[2], mm[3]));
  __stcs(reinterpret_cast<float4*>(exp_avg_sq + off), make_float4(vv[0], vv[1], vv[2], vv[3]));
  if (i < ctx.PQ) {
    const int64_t toff = m * ctx.PQ + i;
    float t[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = __fadd_rn(__fmul_rn(one_minus_tau, hidden ? tf32_unbias(t[j]) : t[j]), __fmul_rn(tau, p[j]));
    if (hidden && !ctx.unbias_out)
      *reinterpret_cast<float4*>(target + toff) = make_float4(tf32_bias(t[0]), tf32_bias(t[1]), tf32_bias(t[2]), tf32_bias(t[3]));
    else
      __stcs(reinterpret_cast<float4*>(target + toff), make_float4(t[0], t[1], t[2], t[3]));
    const float4 th = make_float4(round_tf32(t[0]), round_tf32(t[1]), round_tf32(t[2]), round_tf32(t[3]));
    if (!hidden) *reinterpret_cast<float4*>(ctx.t_shadow + toff) = th;
    if (first_layer)
      *reinterpret_cast<float4*>(ctx.t_shadow_lo + toff) =
          make_float4(round_tf32(t[0] - th.x), round_tf32(t[1] - th.y), round_tf32(t[2] - th.z), round_tf32(t[3] - th.w));
  }
}

}  // namespace iql
