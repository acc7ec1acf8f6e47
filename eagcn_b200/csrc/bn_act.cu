// Fused BatchNorm1d (+ ReLU) (+ dropout) over a [B, C] matrix, forward and backward: the element-wise chain between
// the dense layers of the read-out head (reference models.py:112 Graph_BN, :114-116 relu(bn_den1) + dropout,
// :119 relu(bn_den2)).  Stock PyTorch spends 4 kernels per BatchNorm direction plus one per activation on these
// 256-row matrices (launch-bound: ~85 us per step on the Tox21 shape); here one CTA owns 32 channels for ALL rows, so
// statistics (sums centred on a pivot row -- as well conditioned as torch's two-pass batch_norm), normalisation,
// activation and dropout are one launch per direction.  With B <= 256 a thread keeps its 8 rows in registers: one pass
// over memory (the first version looped 32 rows per thread three times and was latency-bound at ~10 us per launch).
#include "common.cuh"

namespace eagcn {

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

}  // namespace eagcn
using namespace eagcn;

extern "C" int eagcn_bn_act_forward(const void* x, void* y, const void* gamma, const void* beta, void* run_mean,
                                    void* run_var, void* nbt, void* mean_out, void* invstd_out, int64_t B, int64_t C,
                                    int training, int relu, double p_drop, const void* rng, int64_t rng_stream,
                                    double momentum, double eps, void* stream) {
  if (!x || !y || !gamma || !beta || !run_mean || !run_var || !mean_out || !invstd_out || B <= 0 || C <= 0)
    return EAGCN_E_ARG;
  if (p_drop < 0.0 || p_drop >= 1.0 || (training && p_drop > 0.0 && !rng) || B * C >= (int64_t)2147483000) return EAGCN_E_ARG;
  EAGCN_PROF("bn_act_fwd_kernel", stream);
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
  auto kern = B <= kBnWarps * kBnRegRows ? bn_act_bwd_kernel<true> : bn_act_bwd_kernel<false>;
  EAGCN_LAUNCH(kern, (unsigned)((C + 31) / 32), kBnWarps * 32, 0, (cudaStream_t)stream)(
      (const float*)x, (const float*)dy, (const float*)gamma, (const float*)beta, (const float*)mean, (const float*)invstd,
      (float*)dx, (float*)dgamma, (float*)dbeta, (int)B, (int)C, training ? 1 : 0, relu ? 1 : 0, (float)p_drop,
      (const unsigned long long*)rng, (unsigned long long)rng_stream);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
