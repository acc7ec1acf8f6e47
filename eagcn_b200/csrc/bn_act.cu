// Fused BatchNorm1d (+ ReLU) (+ dropout) over a [B, C] matrix, forward and backward: the element-wise chain between
// the dense layers of the read-out head (reference models.py:112 Graph_BN, :114-116 relu(bn_den1) + dropout,
// :119 relu(bn_den2)).  Stock PyTorch spends 4 kernels per BatchNorm direction plus one per activation on these
// 256-row matrices (launch-bound: ~85 us per step on the Tox21 shape); here one CTA owns 32 channels for ALL rows, so
// statistics (sums centred on a pivot row -- as well conditioned as torch's two-pass batch_norm), normalisation,
// activation and dropout are one launch per direction.  With B <= 256 a thread keeps its 8 rows in registers: one pass
// over memory (the first version looped 32 rows per thread three times and was latency-bound at ~10 us per launch).
#include "common.cuh"
#include <initializer_list>

namespace eagcn {

inline int& bn_act_mode() { static int m = 0; return m; }   // 0: float4 kernels where the layout allows; 1: 32-channel kernels

constexpr int kBnWarps = 32;          // one CTA = 32 channels x 32 row lanes (1024 threads)
constexpr int kBnRegRows = 8;         // rows a thread keeps in registers: B <= 256 is a single pass over memory

// column sums of a value pair over the 32 row lanes, fixed order: warp 0 adds the 32 partials of its column and
// broadcasts through shared memory (every thread summing them itself cost ~1k LSU cycles per reduction)
__device__ __forceinline__ float2 bn_col_reduce2(float a, float b, float2 (*s)[32], float2* tot, int warp, int lane) {
  s[warp][lane] = make_float2(a, b);
  __syncthreads();
  if (warp == 0) {
    float2 t = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int w = 0; w < kBnWarps; ++w) { const float2 v = s[w][lane]; t.x += v.x; t.y += v.y; }
    tot[lane] = t;
  }
  __syncthreads();
  return tot[lane];
}

// REG: the thread's rows (warp, warp+32, ...) live in registers (B <= 32*kBnRegRows); otherwise every pass re-reads
// the slab from L1/L2.
template <bool REG>
__global__ void __launch_bounds__(kBnWarps * 32) bn_act_fwd_kernel(
    const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ run_mean, float* __restrict__ run_var, long long* __restrict__ nbt, float* __restrict__ mean_out,
    float* __restrict__ invstd_out, int B, int C, int training, int relu, float p_drop, const unsigned long long* rng,
    unsigned long long stream, float momentum, float eps) {
  pdl_prologue();
  __shared__ float2 s[kBnWarps][32];
  __shared__ float2 s_tot[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  const bool act = c < C;
  float xv[kBnRegRows];
  if (REG) {
#pragma unroll
    for (int k = 0; k < kBnRegRows; ++k) {
      const int r = warp + k * kBnWarps;
      xv[k] = (act && r < B) ? __ldg(x + (size_t)r * C + c) : 0.0f;
    }
  }
  // per-channel parameters: issued with the row loads so that their latency is not paid again after the reduction
  const float g = act ? gamma[c] : 0.0f, b = act ? beta[c] : 0.0f;
  const float rm0 = act ? run_mean[c] : 0.0f, rv0 = act ? run_var[c] : 1.0f;
  float mean, invstd;
  if (training) {
    float var;
    if (REG) {
      // one reduction: sums of (x - p) and (x - p)^2 around the pivot p = x[0][c] (a sample of the column, so the
      // variance formula below has no catastrophic cancellation: |mean - p| is of the order of the spread)
      const float pv = act ? __ldg(x + c) : 0.0f;
      float a = 0.0f, q = 0.0f;
#pragma unroll
      for (int k = 0; k < kBnRegRows; ++k)
        if (warp + k * kBnWarps < B) { const float d = xv[k] - pv; a += d; q = fmaf(d, d, q); }
      const float2 t = bn_col_reduce2(a, q, s, s_tot, warp, lane);
      const float dm = t.x / (float)B;
      mean = pv + dm;
      var = fmaxf(t.y / (float)B - dm * dm, 0.0f);
    } else {
      float a = 0.0f;
      if (act)
        for (int r = warp; r < B; r += kBnWarps) a += __ldg(x + (size_t)r * C + c);
      mean = bn_col_reduce2(a, 0.0f, s, s_tot, warp, lane).x / (float)B;
      float q = 0.0f;
      if (act)
        for (int r = warp; r < B; r += kBnWarps) { const float d = __ldg(x + (size_t)r * C + c) - mean; q = fmaf(d, d, q); }
      var = bn_col_reduce2(q, 0.0f, s, s_tot, warp, lane).x / (float)B;
    }
    invstd = 1.0f / sqrtf(var + eps);
    if (warp == 0 && act) {                                  // running statistics like nn.BatchNorm1d (unbiased variance)
      run_mean[c] = (1.0f - momentum) * rm0 + momentum * mean;
      run_var[c] = (1.0f - momentum) * rv0 + momentum * var * ((float)B / (float)(B > 1 ? B - 1 : 1));
    }
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
  } else {
    mean = rm0;
    invstd = act ? 1.0f / sqrtf(rv0 + eps) : 0.0f;
  }
  if (!act) return;
  if (warp == 0) { mean_out[c] = mean; invstd_out[c] = invstd; }
  const bool drop = training && p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  unsigned long long seed = 0, off = 0;
  if (drop) { seed = rng[0]; off = rng[1]; }
  const Philox ph(seed);
  auto emit = [&](int r, float xr) {
    const size_t idx = (size_t)r * C + c;
    float z = (xr - mean) * invstd * g + b;
    if (relu) z = fmaxf(z, 0.0f);
    if (drop) z = dropout_keep(ph, off, stream, (unsigned long long)idx, p_drop) ? z * scale : 0.0f;
    y[idx] = z;
  };
  if (REG) {
#pragma unroll
    for (int k = 0; k < kBnRegRows; ++k) { const int r = warp + k * kBnWarps; if (r < B) emit(r, xv[k]); }
  } else {
    for (int r = warp; r < B; r += kBnWarps) emit(r, __ldg(x + (size_t)r * C + c));
  }
}

template <bool REG>
__global__ void __launch_bounds__(kBnWarps * 32) bn_act_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ mean_in, const float* __restrict__ invstd_in,
    float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C, int training, int relu,
    float p_drop, const unsigned long long* rng, unsigned long long stream) {
  pdl_prologue();
  __shared__ float2 s[kBnWarps][32];
  __shared__ float2 s_tot[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  const bool act = c < C;
  const float mean = act ? mean_in[c] : 0.0f, invstd = act ? invstd_in[c] : 0.0f;
  const float g = act ? gamma[c] : 0.0f, b = act ? beta[c] : 0.0f;
  const bool drop = training && p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  unsigned long long seed = 0, off = 0;
  if (drop) { seed = rng[0]; off = rng[1]; }
  const Philox ph(seed);
  // gradient reaching the BatchNorm output of element (r, c): dropout and ReLU replayed from x
  auto grad_at = [&](int r, float& xh) {
    const size_t idx = (size_t)r * C + c;
    xh = (__ldg(x + idx) - mean) * invstd;
    float gr = __ldg(dy + idx);
    if (drop) gr = dropout_keep(ph, off, stream, (unsigned long long)idx, p_drop) ? gr * scale : 0.0f;
    if (relu && !(xh * g + b > 0.0f)) gr = 0.0f;
    return gr;
  };
  float gv[kBnRegRows], hv[kBnRegRows];
  float s1 = 0.0f, s2 = 0.0f;
  if (REG) {
#pragma unroll
    for (int k = 0; k < kBnRegRows; ++k) {
      const int r = warp + k * kBnWarps;
      gv[k] = 0.0f; hv[k] = 0.0f;
      if (act && r < B) gv[k] = grad_at(r, hv[k]);
      s1 += gv[k]; s2 = fmaf(gv[k], hv[k], s2);
    }
  } else if (act) {
    for (int r = warp; r < B; r += kBnWarps) { float xh; const float gr = grad_at(r, xh); s1 += gr; s2 = fmaf(gr, xh, s2); }
  }
  {
    const float2 t = bn_col_reduce2(s1, s2, s, s_tot, warp, lane);
    s1 = t.x; s2 = t.y;
  }
  if (!act) return;
  if (warp == 0) { dbeta[c] = s1; dgamma[c] = s2; }
  const float m1 = s1 / (float)B, m2 = s2 / (float)B, gi = g * invstd;
  if (REG) {
#pragma unroll
    for (int k = 0; k < kBnRegRows; ++k) {
      const int r = warp + k * kBnWarps;
      if (r < B) dx[(size_t)r * C + c] = gi * (training ? (gv[k] - m1 - hv[k] * m2) : gv[k]);
    }
  } else {
    for (int r = warp; r < B; r += kBnWarps) {
      float xh;
      const float gr = grad_at(r, xh);
      dx[(size_t)r * C + c] = gi * (training ? (gr - m1 - xh * m2) : gr);
    }
  }
}

// ---- float4 form: a CTA owns QL float4 channel groups (4*QL channels) for all rows ----------------------------------
// The 32-channel kernels above leave C/32 CTAs (8 for bn_den1's 256 channels) with 1024 threads each and spend one
// Philox call per element; here C/(4*QL) CTAs of 256 threads, a thread holds 4 consecutive channels of up to
// kBnVecRows rows in registers, one Philox call serves its 4 elements, and the column sums go through a shuffle tree +
// one shared-memory exchange between the 8 warps (fixed order: bit-reproducible).
constexpr int kBnVecThreads = 256;
constexpr int kBnVecRows = 4;

struct F8 { float4 a, b; };

template <int QL>
__device__ __forceinline__ F8 bn_vec_reduce(F8 v, F8 (*s)[QL], int tid) {
  float* f = reinterpret_cast<float*>(&v);
#pragma unroll
  for (int o = 16; o >= QL; o >>= 1)
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += __shfl_xor_sync(0xffffffffu, f[i], o);
  const int warp = tid >> 5, lane = tid & 31, q = tid % QL;
  if (lane < QL) s[warp][lane] = v;
  __syncthreads();
  F8 t = s[0][q];
  float* g = reinterpret_cast<float*>(&t);
#pragma unroll
  for (int w = 1; w < kBnVecThreads / 32; ++w) {
    const F8 u = s[w][q];
    const float* h = reinterpret_cast<const float*>(&u);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += h[i];
  }
  return t;
}

template <int QL>
__global__ void __launch_bounds__(kBnVecThreads) bn_act_fwd_vec_kernel(
    const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ run_mean, float* __restrict__ run_var, long long* __restrict__ nbt, float* __restrict__ mean_out,
    float* __restrict__ invstd_out, int B, int C, int training, int relu, float p_drop, const unsigned long long* rng,
    unsigned long long stream, float momentum, float eps) {
  pdl_prologue();
  constexpr int RL = kBnVecThreads / QL;
  __shared__ F8 s[kBnVecThreads / 32][QL];
  const int tid = threadIdx.x, q = tid % QL, rl = tid / QL;
  const int c4 = blockIdx.x * QL + q, c = c4 * 4, nc4 = C >> 2;
  const bool act = c < C;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4 xv[kBnVecRows];
#pragma unroll
  for (int k = 0; k < kBnVecRows; ++k) {
    const int r = rl + k * RL;
    xv[k] = (act && r < B) ? __ldg(x4 + (size_t)r * nc4 + c4) : zero;
  }
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 g = act ? *reinterpret_cast<const float4*>(gamma + c) : zero;
  const float4 b = act ? *reinterpret_cast<const float4*>(beta + c) : zero;
  const float4 rm0 = act ? *reinterpret_cast<const float4*>(run_mean + c) : zero;
  const float4 rv0 = act ? *reinterpret_cast<const float4*>(run_var + c) : one;
  float4 mean, invstd;
  if (training) {
    // sums of (x - p) and (x - p)^2 around the pivot p = x[0][c] (see the 32-channel kernel)
    const float4 pv = act ? __ldg(x4 + c4) : zero;
    F8 v; v.a = zero; v.b = zero;
#pragma unroll
    for (int k = 0; k < kBnVecRows; ++k)
      if (rl + k * RL < B) {
        const float dx = xv[k].x - pv.x, dy = xv[k].y - pv.y, dz = xv[k].z - pv.z, dw = xv[k].w - pv.w;
        v.a.x += dx; v.a.y += dy; v.a.z += dz; v.a.w += dw;
        v.b.x = fmaf(dx, dx, v.b.x); v.b.y = fmaf(dy, dy, v.b.y); v.b.z = fmaf(dz, dz, v.b.z); v.b.w = fmaf(dw, dw, v.b.w);
      }
    const F8 t = bn_vec_reduce<QL>(v, s, tid);
    const float inv_b = 1.0f / (float)B, unb = (float)B / (float)(B > 1 ? B - 1 : 1);
    float4 var;
    {
      const float d0 = t.a.x * inv_b, d1 = t.a.y * inv_b, d2 = t.a.z * inv_b, d3 = t.a.w * inv_b;
      mean = make_float4(pv.x + d0, pv.y + d1, pv.z + d2, pv.w + d3);
      var = make_float4(fmaxf(t.b.x * inv_b - d0 * d0, 0.f), fmaxf(t.b.y * inv_b - d1 * d1, 0.f),
                        fmaxf(t.b.z * inv_b - d2 * d2, 0.f), fmaxf(t.b.w * inv_b - d3 * d3, 0.f));
    }
    invstd = make_float4(1.0f / sqrtf(var.x + eps), 1.0f / sqrtf(var.y + eps), 1.0f / sqrtf(var.z + eps),
                         1.0f / sqrtf(var.w + eps));
    if (rl == 0 && act) {                                    // running statistics like nn.BatchNorm1d (unbiased variance)
      const float om = 1.0f - momentum;
      *reinterpret_cast<float4*>(run_mean + c) = make_float4(om * rm0.x + momentum * mean.x, om * rm0.y + momentum * mean.y,
                                                             om * rm0.z + momentum * mean.z, om * rm0.w + momentum * mean.w);
      *reinterpret_cast<float4*>(run_var + c) =
          make_float4(om * rv0.x + momentum * var.x * unb, om * rv0.y + momentum * var.y * unb,
                      om * rv0.z + momentum * var.z * unb, om * rv0.w + momentum * var.w * unb);
    }
    if (nbt && blockIdx.x == 0 && tid == 0) *nbt += 1;
  } else {
    mean = rm0;
    invstd = make_float4(1.0f / sqrtf(rv0.x + eps), 1.0f / sqrtf(rv0.y + eps), 1.0f / sqrtf(rv0.z + eps),
                         1.0f / sqrtf(rv0.w + eps));
  }
  if (!act) return;
  if (rl == 0) {
    *reinterpret_cast<float4*>(mean_out + c) = mean;
    *reinterpret_cast<float4*>(invstd_out + c) = invstd;
  }
  const bool drop = training && p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  unsigned long long seed = 0, off = 0;
  if (drop) { seed = rng[0]; off = rng[1]; }
  const Philox ph(seed);
  float4* y4 = reinterpret_cast<float4*>(y);
#pragma unroll
  for (int k = 0; k < kBnVecRows; ++k) {
    const int r = rl + k * RL;
    if (r >= B) continue;
    float4 z;
    z.x = (xv[k].x - mean.x) * invstd.x * g.x + b.x;
    z.y = (xv[k].y - mean.y) * invstd.y * g.y + b.y;
    z.z = (xv[k].z - mean.z) * invstd.z * g.z + b.z;
    z.w = (xv[k].w - mean.w) * invstd.w * g.w + b.w;
    if (relu) { z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f); }
    if (drop) {
      bool kp[4];
      dropout_keep4(ph, off, stream, (unsigned long long)r * C + c, p_drop, kp);
      z.x = kp[0] ? z.x * scale : 0.f; z.y = kp[1] ? z.y * scale : 0.f;
      z.z = kp[2] ? z.z * scale : 0.f; z.w = kp[3] ? z.w * scale : 0.f;
    }
    y4[(size_t)r * nc4 + c4] = z;
  }
}

template <int QL>
__global__ void __launch_bounds__(kBnVecThreads) bn_act_bwd_vec_kernel(
    const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ mean_in, const float* __restrict__ invstd_in,
    float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C, int training, int relu,
    float p_drop, const unsigned long long* rng, unsigned long long stream) {
  pdl_prologue();
  constexpr int RL = kBnVecThreads / QL;
  __shared__ F8 s[kBnVecThreads / 32][QL];
  const int tid = threadIdx.x, q = tid % QL, rl = tid / QL;
  const int c4 = blockIdx.x * QL + q, c = c4 * 4, nc4 = C >> 2;
  const bool act = c < C;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mean = act ? *reinterpret_cast<const float4*>(mean_in + c) : zero;
  const float4 invstd = act ? *reinterpret_cast<const float4*>(invstd_in + c) : zero;
  const float4 g = act ? *reinterpret_cast<const float4*>(gamma + c) : zero;
  const float4 b = act ? *reinterpret_cast<const float4*>(beta + c) : zero;
  const bool drop = training && p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  unsigned long long seed = 0, off = 0;
  if (drop) { seed = rng[0]; off = rng[1]; }
  const Philox ph(seed);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* dy4 = reinterpret_cast<const float4*>(dy);
  float4 gv[kBnVecRows], hv[kBnVecRows];
  F8 v; v.a = zero; v.b = zero;
#pragma unroll
  for (int k = 0; k < kBnVecRows; ++k) {
    const int r = rl + k * RL;
    gv[k] = zero; hv[k] = zero;
    if (act && r < B) {
      const float4 xr = __ldg(x4 + (size_t)r * nc4 + c4);
      float4 gr = __ldg(dy4 + (size_t)r * nc4 + c4);
      float4 xh;
      xh.x = (xr.x - mean.x) * invstd.x; xh.y = (xr.y - mean.y) * invstd.y;
      xh.z = (xr.z - mean.z) * invstd.z; xh.w = (xr.w - mean.w) * invstd.w;
      if (drop) {
        bool kp[4];
        dropout_keep4(ph, off, stream, (unsigned long long)r * C + c, p_drop, kp);
        gr.x = kp[0] ? gr.x * scale : 0.f; gr.y = kp[1] ? gr.y * scale : 0.f;
        gr.z = kp[2] ? gr.z * scale : 0.f; gr.w = kp[3] ? gr.w * scale : 0.f;
      }
      if (relu) {
        if (!(xh.x * g.x + b.x > 0.0f)) gr.x = 0.f;
        if (!(xh.y * g.y + b.y > 0.0f)) gr.y = 0.f;
        if (!(xh.z * g.z + b.z > 0.0f)) gr.z = 0.f;
        if (!(xh.w * g.w + b.w > 0.0f)) gr.w = 0.f;
      }
      gv[k] = gr; hv[k] = xh;
      v.a.x += gr.x; v.a.y += gr.y; v.a.z += gr.z; v.a.w += gr.w;
      v.b.x = fmaf(gr.x, xh.x, v.b.x); v.b.y = fmaf(gr.y, xh.y, v.b.y);
      v.b.z = fmaf(gr.z, xh.z, v.b.z); v.b.w = fmaf(gr.w, xh.w, v.b.w);
    }
  }
  const F8 t = bn_vec_reduce<QL>(v, s, tid);
  if (!act) return;
  if (rl == 0) {
    *reinterpret_cast<float4*>(dbeta + c) = t.a;
    *reinterpret_cast<float4*>(dgamma + c) = t.b;
  }
  const float inv_b = 1.0f / (float)B;
  const float4 m1 = make_float4(t.a.x * inv_b, t.a.y * inv_b, t.a.z * inv_b, t.a.w * inv_b);
  const float4 m2 = make_float4(t.b.x * inv_b, t.b.y * inv_b, t.b.z * inv_b, t.b.w * inv_b);
  const float4 gi = make_float4(g.x * invstd.x, g.y * invstd.y, g.z * invstd.z, g.w * invstd.w);
  float4* dx4 = reinterpret_cast<float4*>(dx);
#pragma unroll
  for (int k = 0; k < kBnVecRows; ++k) {
    const int r = rl + k * RL;
    if (r >= B) continue;
    float4 o;
    if (training) {
      o.x = gi.x * (gv[k].x - m1.x - hv[k].x * m2.x); o.y = gi.y * (gv[k].y - m1.y - hv[k].y * m2.y);
      o.z = gi.z * (gv[k].z - m1.z - hv[k].z * m2.z); o.w = gi.w * (gv[k].w - m1.w - hv[k].w * m2.w);
    } else {
      o.x = gi.x * gv[k].x; o.y = gi.y * gv[k].y; o.z = gi.z * gv[k].z; o.w = gi.w * gv[k].w;
    }
    dx4[(size_t)r * nc4 + c4] = o;
  }
}

// float4 kernels: every array 16-byte aligned, C % 4 == 0, all rows of a thread in registers.  QL = 2 (8 channels per
// CTA) when that still leaves >= 64 CTAs, else one float4 group per CTA.
static int bn_vec_ql(int64_t B, int64_t C, std::initializer_list<const void*> ptrs) {
  if (bn_act_mode() != 0 || (C & 3)) return 0;
  for (const void* p : ptrs) if (!aligned16(p)) return 0;
  const int ql = (C / 4 >= 128) ? 2 : 1;
  if (B > (int64_t)(kBnVecThreads / ql) * kBnVecRows) return (ql == 2 && B <= (int64_t)kBnVecThreads * kBnVecRows) ? 1 : 0;
  return ql;
}

}  // namespace eagcn
using namespace eagcn;

extern "C" int eagcn_set_bn_act_mode(int mode) {
  if (mode < 0 || mode > 1) return EAGCN_E_ARG;
  eagcn::bn_act_mode() = mode;
  return 0;
}
extern "C" int eagcn_get_bn_act_mode(void) { return eagcn::bn_act_mode(); }

extern "C" int eagcn_bn_act_forward(const void* x, void* y, const void* gamma, const void* beta, void* run_mean,
                                    void* run_var, void* nbt, void* mean_out, void* invstd_out, int64_t B, int64_t C,
                                    int training, int relu, double p_drop, const void* rng, int64_t rng_stream,
                                    double momentum, double eps, void* stream) {
  if (!x || !y || !gamma || !beta || !run_mean || !run_var || !mean_out || !invstd_out || B <= 0 || C <= 0)
    return EAGCN_E_ARG;
  if (p_drop < 0.0 || p_drop >= 1.0 || (training && p_drop > 0.0 && !rng) || B * C >= (int64_t)2147483000) return EAGCN_E_ARG;
  EAGCN_PROF("bn_act_fwd_kernel", stream);
  if (const int ql = bn_vec_ql(B, C, {x, y, gamma, beta, run_mean, run_var, mean_out, invstd_out})) {
    auto kv = ql == 2 ? bn_act_fwd_vec_kernel<2> : bn_act_fwd_vec_kernel<1>;
    EAGCN_LAUNCH(kv, (unsigned)((C / 4 + ql - 1) / ql), kBnVecThreads, 0, (cudaStream_t)stream)(
        (const float*)x, (float*)y, (const float*)gamma, (const float*)beta, (float*)run_mean, (float*)run_var,
        (long long*)nbt, (float*)mean_out, (float*)invstd_out, (int)B, (int)C, training ? 1 : 0, relu ? 1 : 0, (float)p_drop,
        (const unsigned long long*)rng, (unsigned long long)rng_stream, (float)momentum, (float)eps);
    EAGCN_LAUNCH_CHECK();
    return 0;
  }
  auto kern = B <= kBnWarps * kBnRegRows ? bn_act_fwd_kernel<true> : bn_act_fwd_kernel<false>;
  EAGCN_LAUNCH(kern, (unsigned)((C + 31) / 32), kBnWarps * 32, 0, (cudaStream_t)stream)(
      (const float*)x, (float*)y, (const float*)gamma, (const float*)beta, (float*)run_mean, (float*)run_var,
      (long long*)nbt, (float*)mean_out, (float*)invstd_out, (int)B, (int)C, training ? 1 : 0, relu ? 1 : 0, (float)p_drop,
      (const unsigned long long*)rng, (unsigned long long)rng_stream, (float)momentum, (float)eps);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int eagcn_bn_act_backward(const void* x, const void* dy, const void* gamma, const void* beta, const void* mean,
                                     const void* invstd, void* dx, void* dgamma, void* dbeta, int64_t B, int64_t C,
                                     int training, int relu, double p_drop, const void* rng, int64_t rng_stream,
                                     void* stream) {
  if (!x || !dy || !gamma || !beta || !mean || !invstd || !dx || !dgamma || !dbeta || B <= 0 || C <= 0) return EAGCN_E_ARG;
  if (p_drop < 0.0 || p_drop >= 1.0 || (training && p_drop > 0.0 && !rng) || B * C >= (int64_t)2147483000) return EAGCN_E_ARG;
  EAGCN_PROF("bn_act_bwd_kernel", stream);
  if (const int ql = bn_vec_ql(B, C, {x, dy, gamma, beta, mean, invstd, dx, dgamma, dbeta})) {
    auto kv = ql == 2 ? bn_act_bwd_vec_kernel<2> : bn_act_bwd_vec_kernel<1>;
    EAGCN_LAUNCH(kv, (unsigned)((C / 4 + ql - 1) / ql), kBnVecThreads, 0, (cudaStream_t)stream)(
        (const float*)x, (const float*)dy, (const float*)gamma, (const float*)beta, (const float*)mean, (const float*)invstd,
        (float*)dx, (float*)dgamma, (float*)dbeta, (int)B, (int)C, training ? 1 : 0, relu ? 1 : 0, (float)p_drop,
        (const unsigned long long*)rng, (unsigned long long)rng_stream);
    EAGCN_LAUNCH_CHECK();
    return 0;
  }
  auto kern = B <= kBnWarps * kBnRegRows ? bn_act_bwd_kernel<true> : bn_act_bwd_kernel<false>;
  EAGCN_LAUNCH(kern, (unsigned)((C + 31) / 32), kBnWarps * 32, 0, (cudaStream_t)stream)(
      (const float*)x, (const float*)dy, (const float*)gamma, (const float*)beta, (const float*)mean, (const float*)invstd,
      (float*)dx, (float*)dgamma, (float*)dbeta, (int)B, (int)C, training ? 1 : 0, relu ? 1 : 0, (float)p_drop,
      (const unsigned long long*)rng, (unsigned long long)rng_stream);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
