// Fused read-out head of EAGCN (reference models.py:112-120):
//     x0 -> Graph_BN -> den1 -> bn_den1 -> ReLU -> dropout -> den2 (= graph_representation) -> bn_den2 -> ReLU -> den3
// with three bias-free Dense layers (layers.py:360-392) and three nn.BatchNorm1d over the batch rows.
//
// On the reference this is ~12 ATen launches forward and ~25 backward for ~0.3 GFLOP of work -- pure launch
// latency.  Here: ONE kernel forward and ONE kernel backward.  The BatchNorms need whole-batch column
// statistics between the products, so the kernels run as a single resident grid (<= 64 CTAs, far below the
// 148 SMs) and separate their phases with a software grid barrier; every BatchNorm is folded into the operand
// loader of the product that consumes it (normalise / ReLU / dropout on load), column statistics come out of
// the producing product's epilogue as per-row-tile partials that the next phase reduces in fixed order.
// All arithmetic is fp32 FFMA (these products are far too small for the tensor-core pipeline to pay off).
#include "common.cuh"

namespace eagcn {
namespace head {

constexpr int HT = 256;              // threads per CTA
constexpr int TM = 32, TN = 64, TK = 32;
constexpr int kMaxCtas = 64;

struct HeadDev {
  int B, F, D1, D2, NC;
  int training; float p_drop; double eps, momentum;
  const float* x0;
  const float *g0, *b0, *g1, *b1, *g2, *b2;           // BatchNorm affine
  float *rm0, *rv0, *rm1, *rv1, *rm2, *rv2;           // running stats
  long long *nbt0, *nbt1, *nbt2;
  const float *W1, *W2, *W3;                           // [F,D1] [D1,D2] [D2,NC]
  const unsigned long long* rng; unsigned long long rng_stream;
  float *a1, *a2, *out;                                // [B,D1] [B,D2] [B,NC]
  float *mean0, *invstd0, *mean1, *invstd1, *mean2, *invstd2;
  float* part;                                         // [rowtiles][2][max(F,D1,D2)] column partials
  float* part2;                                        // second partial buffer (ping-pong between phases)
  unsigned* bar;                                       // [2] grid barrier (count, generation), zero-initialised once
  // backward
  const float *d_out, *d_a2;                           // [B,NC], optional [B,D2]
  float *g2buf, *g1buf, *dh0;                          // [B,D2] [B,D1] [B,F] workspaces
  float *dx0, *dW1, *dW2, *dW3, *dg0, *db0, *dg1, *db1, *dg2, *db2;
};

__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned gen = atomicAdd(&bar[1], 0u);
    if (atomicAdd(&bar[0], 1u) == nblocks - 1) {
      atomicExch(&bar[0], 0u);
      __threadfence();
      atomicAdd(&bar[1], 1u);
    } else {
      while (atomicAdd(&bar[1], 0u) == gen) __nanosleep(40);
    }
    __threadfence();
  }
  __syncthreads();
}

// C[TM x TN] tile of  sum_k A(m,k) * Bm(k,n);  la(m,k) / lb(k,n) are element loaders (they fuse the BatchNorm /
// ReLU / dropout / gradient formulas), epi(m, n, value) consumes the results.  256 threads: 8 (m) x 32 (n) thread
// grid, 4 x 2 outputs per thread.
template <class LA, class LB, class EPI>
__device__ __forceinline__ void tile_gemm(int m0, int n0, int M, int N, int K, LA la, LB lb, EPI epi,
                                          float (*sA)[TK][TM + 1], float (*sB)[TK][TN + 1],
                                          float* __restrict__ part = nullptr, int ldp = 0) {
  // epi(m, n, value) returns a float2 (p, q); when `part` is given, the column sums of p and q over this tile's rows
  // are stored to part[row_tile][0 / 1][n] (fixed order) -- the BatchNorm statistics of the product's output.
  // Software pipeline: the operand elements of k-block i+1 are fetched into registers (all loads issued back to
  // back, nothing stored in between) while k-block i is multiplied out of shared memory; two smem buffers.
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // n = tx + 32*j (j<2), m = ty*4 + i (i<4)
  constexpr int NA = TM * TK / HT, NB = TK * TN / HT;
  float ra[NA], rb[NB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = threadIdx.x + i * HT;
      const int k = idx % TK, m = idx / TK;
      const int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < M && gk < K) ? la(gm, gk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = threadIdx.x + i * HT;
      const int n = idx % TN, k = idx / TN;
      const int gn = n0 + n, gk = k0 + k;
      rb[i] = (gn < N && gk < K) ? lb(gk, gn) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) { const int idx = threadIdx.x + i * HT; sA[buf][idx % TK][idx / TK] = ra[i]; }
#pragma unroll
    for (int i = 0; i < NB; ++i) { const int idx = threadIdx.x + i * HT; sB[buf][idx / TN][idx % TN] = rb[i]; }
  };
  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
  const int nkb = (K + TK - 1) / TK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) fetch((kb + 1) * TK);
#pragma unroll 8
    for (int k = 0; k < TK; ++k) {
      const float b0v = sB[buf][k][tx], b1v = sB[buf][k][tx + 32];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = sA[buf][k][ty * 4 + i];
        acc[i][0] = fmaf(a, b0v, acc[i][0]);
        acc[i][1] = fmaf(a, b1v, acc[i][1]);
      }
    }
    if (kb + 1 < nkb) stash(buf ^ 1);
    __syncthreads();
  }
  float ps[2] = {0.f, 0.f}, qs[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx + 32 * j;
      if (gm < M && gn < N) { const float2 r = epi(gm, gn, acc[i][j]); ps[j] += r.x; qs[j] += r.y; }
    }
  if (part) {                                                      // 8 row groups -> fixed-order column sums
    float (*red)[TN + 1] = sB[0];                                  // reuse: [0..7] p sums, [8..15] q sums
#pragma unroll
    for (int j = 0; j < 2; ++j) { red[ty][tx + 32 * j] = ps[j]; red[8 + ty][tx + 32 * j] = qs[j]; }
    __syncthreads();
    if (threadIdx.x < TN) {
      const int gn = n0 + threadIdx.x;
      if (gn < N) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += red[w][threadIdx.x]; b += red[8 + w][threadIdx.x]; }
        part[((size_t)(m0 / TM) * 2 + 0) * ldp + gn] = a;
        part[((size_t)(m0 / TM) * 2 + 1) * ldp + gn] = b;
      }
    }
    __syncthreads();
  }
}

// column statistics of an [R, C] row-major matrix restricted to one row tile: part[tile][0][c] = sum, [1] = sum sq
__device__ __forceinline__ void col_partials(const float* __restrict__ X, int R, int C, int tile, float* __restrict__ part,
                                             int ldp) {
  const int r0 = tile * TM, r1 = min(R, r0 + TM);
  for (int c = threadIdx.x; c < C; c += HT) {
    float s = 0.f, q = 0.f;
    for (int r = r0; r < r1; ++r) { const float v = X[(size_t)r * C + c]; s += v; q = fmaf(v, v, q); }
    part[((size_t)tile * 2 + 0) * ldp + c] = s;
    part[((size_t)tile * 2 + 1) * ldp + c] = q;
  }
}

// BatchNorm1d statistics of column c from row-tile partials (fixed order, double)
__device__ __forceinline__ void bn_stats_from_partials(const float* __restrict__ part, int ntile, int ldp, int c, int R,
                                                       double eps, float& mean, float& invstd, double& var_unbiased) {
  double s = 0.0, q = 0.0;
  for (int t = 0; t < ntile; ++t) { s += (double)part[((size_t)t * 2 + 0) * ldp + c]; q += (double)part[((size_t)t * 2 + 1) * ldp + c]; }
  const double m = s / R;
  double var = q / R - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  invstd = (float)(1.0 / sqrt(var + eps));
  var_unbiased = R > 1 ? var * ((double)R / (R - 1)) : var;
}

__device__ __forceinline__ void finalize_bn(const HeadDev& h, int C, int ntile, int ldp, const float* part, float* mean,
                                            float* invstd, float* rm, float* rv, long long* nbt) {
  // every CTA computes all C statistics (tiny), CTA 0 publishes them and updates the running buffers
  for (int c = threadIdx.x; c < C; c += HT) {
    float mu, is;
    if (h.training) {
      double vu;
      bn_stats_from_partials(part, ntile, ldp, c, h.B, h.eps, mu, is, vu);
      if (blockIdx.x == 0) {
        rm[c] = (float)((1.0 - h.momentum) * (double)rm[c] + h.momentum * (double)mu);
        rv[c] = (float)((1.0 - h.momentum) * (double)rv[c] + h.momentum * vu);
      }
    } else {
      mu = rm[c]; is = 1.0f / sqrtf(rv[c] + (float)h.eps);
    }
    if (blockIdx.x == 0) { mean[c] = mu; invstd[c] = is; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && h.training && nbt) nbt[0] += 1;
}

// ======================================================================================================
__global__ void __launch_bounds__(HT) head_fwd_kernel(HeadDev h) {
  pdl_prologue();
  __shared__ float sA[2][TK][TM + 1];
  __shared__ float sB[2][TK][TN + 1];
  extern __shared__ float s_stat[];                    // [2][max(F, D1, D2)] mean / invstd of the operand being normalised
  const int nb = gridDim.x;
  const int rt = (h.B + TM - 1) / TM;                  // row tiles
  const int ldp = max(h.F, max(h.D1, h.D2));
  const bool drop = h.training && h.p_drop > 0.f;
  const float dscale = drop ? 1.0f / (1.0f - h.p_drop) : 1.0f;
  unsigned long long seed = 0, roff = 0;
  if (drop) { seed = h.rng[0]; roff = h.rng[1]; }
  const Philox ph(seed);

  // ---- phase 0: column partials of x0 (Graph_BN statistics) ----
  if (h.training)
    for (int t = blockIdx.x; t < rt; t += nb) col_partials(h.x0, h.B, h.F, t, h.part, ldp);
  grid_sync(h.bar, nb);

  // ---- phase 1: a1 = BN0(x0) @ W1 ----
  finalize_bn(h, h.F, rt, ldp, h.part, h.mean0, h.invstd0, h.rm0, h.rv0, h.nbt0);
  for (int c = threadIdx.x; c < h.F; c += HT) {        // per-CTA copy of the statistics (scale / shift form)
    float mu, is;
    if (h.training) { double vu; bn_stats_from_partials(h.part, rt, ldp, c, h.B, h.eps, mu, is, vu); }
    else { mu = h.rm0[c]; is = 1.0f / sqrtf(h.rv0[c] + (float)h.eps); }
    s_stat[c] = is * h.g0[c];
    s_stat[ldp + c] = h.b0[c] - mu * is * h.g0[c];
  }
  __syncthreads();
  {
    const int nt = (h.D1 + TN - 1) / TN;
    for (int t = blockIdx.x; t < rt * nt; t += nb) {
      const int m0 = (t / nt) * TM, n0 = (t % nt) * TN;
      tile_gemm(m0, n0, h.B, h.D1, h.F,
                [&](int m, int k) { return fmaf(h.x0[(size_t)m * h.F + k], s_stat[k], s_stat[ldp + k]); },
                [&](int k, int n) { return h.W1[(size_t)k * h.D1 + n]; },
                [&](int m, int n, float v) { h.a1[(size_t)m * h.D1 + n] = v; return make_float2(v, v * v); }, sA, sB,
                h.training ? h.part2 : nullptr, ldp);
    }
  }
  grid_sync(h.bar, nb);          // a1 and its column partials (part2) complete

  // ---- phase 2: a2 = dropout(relu(BN1(a1))) @ W2 ----
  finalize_bn(h, h.D1, rt, ldp, h.part2, h.mean1, h.invstd1, h.rm1, h.rv1, h.nbt1);
  for (int c = threadIdx.x; c < h.D1; c += HT) {
    float mu, is;
    if (h.training) { double vu; bn_stats_from_partials(h.part2, rt, ldp, c, h.B, h.eps, mu, is, vu); }
    else { mu = h.rm1[c]; is = 1.0f / sqrtf(h.rv1[c] + (float)h.eps); }
    s_stat[c] = is * h.g1[c];
    s_stat[ldp + c] = h.b1[c] - mu * is * h.g1[c];
  }
  __syncthreads();
  {
    const int nt = (h.D2 + TN - 1) / TN;
    for (int t = blockIdx.x; t < rt * nt; t += nb) {
      const int m0 = (t / nt) * TM, n0 = (t % nt) * TN;
      tile_gemm(m0, n0, h.B, h.D2, h.D1,
                [&](int m, int k) {
                  float v = fmaxf(fmaf(h.a1[(size_t)m * h.D1 + k], s_stat[k], s_stat[ldp + k]), 0.f);
                  if (drop) v = dropout_keep(ph, roff, h.rng_stream, (unsigned long long)m * h.D1 + k, h.p_drop) ? v * dscale : 0.f;
                  return v;
                },
                [&](int k, int n) { return h.W2[(size_t)k * h.D2 + n]; },
                [&](int m, int n, float v) { h.a2[(size_t)m * h.D2 + n] = v; return make_float2(v, v * v); }, sA, sB,
                h.training ? h.part : nullptr, ldp);
    }
  }
  grid_sync(h.bar, nb);          // a2 and its column partials (part) complete

  // ---- phase 3: out = relu(BN2(a2)) @ W3 ----
  finalize_bn(h, h.D2, rt, ldp, h.part, h.mean2, h.invstd2, h.rm2, h.rv2, h.nbt2);
  for (int c = threadIdx.x; c < h.D2; c += HT) {
    float mu, is;
    if (h.training) { double vu; bn_stats_from_partials(h.part, rt, ldp, c, h.B, h.eps, mu, is, vu); }
    else { mu = h.rm2[c]; is = 1.0f / sqrtf(h.rv2[c] + (float)h.eps); }
    s_stat[c] = is * h.g2[c];
    s_stat[ldp + c] = h.b2[c] - mu * is * h.g2[c];
  }
  __syncthreads();
  {
    const int nt = (h.NC + TN - 1) / TN;
    for (int t = blockIdx.x; t < rt * nt; t += nb) {
      const int m0 = (t / nt) * TM, n0 = (t % nt) * TN;
      tile_gemm(m0, n0, h.B, h.NC, h.D2,
                [&](int m, int k) { return fmaxf(fmaf(h.a2[(size_t)m * h.D2 + k], s_stat[k], s_stat[ldp + k]), 0.f); },
                [&](int k, int n) { return h.W3[(size_t)k * h.NC + n]; },
                [&](int m, int n, float v) { h.out[(size_t)m * h.NC + n] = v; return make_float2(0.f, 0.f); }, sA, sB);
    }
  }
}

// ======================================================================================================
// backward helpers: sums S1 = sum_m g, S2 = sum_m g * xhat per column from row-tile partials
__device__ __forceinline__ void bwd_sums(const float* __restrict__ part, int ntile, int ldp, int c, float& s1, float& s2) {
  double a = 0.0, b = 0.0;
  for (int t = 0; t < ntile; ++t) { a += (double)part[((size_t)t * 2 + 0) * ldp + c]; b += (double)part[((size_t)t * 2 + 1) * ldp + c]; }
  s1 = (float)a; s2 = (float)b;
}

// partials of (g, g*xhat) over one row tile, g and the pre-BatchNorm activation given as [R, C] matrices
__device__ __forceinline__ void bwd_partials(const float* __restrict__ G, const float* __restrict__ Xpre, const float* mean,
                                             const float* invstd, int R, int C, int tile, float* __restrict__ part, int ldp) {
  const int r0 = tile * TM, r1 = min(R, r0 + TM);
  for (int c = threadIdx.x; c < C; c += HT) {
    const float mu = mean[c], is = invstd[c];
    float s = 0.f, q = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float g = G[(size_t)r * C + c];
      s += g; q = fmaf(g, (Xpre[(size_t)r * C + c] - mu) * is, q);
    }
    part[((size_t)tile * 2 + 0) * ldp + c] = s;
    part[((size_t)tile * 2 + 1) * ldp + c] = q;
  }
}

__global__ void __launch_bounds__(HT) head_bwd_kernel(HeadDev h) {
  pdl_prologue();
  __shared__ float sA[2][TK][TM + 1];
  __shared__ float sB[2][TK][TN + 1];
  extern __shared__ float s_c[];                       // [4][ldp]: per-column coefficients of the current BatchNorm
  const int nb = gridDim.x;
  const int rt = (h.B + TM - 1) / TM;
  const int ldp = max(h.F, max(h.D1, h.D2));
  const bool drop = h.training && h.p_drop > 0.f;
  const float dscale = drop ? 1.0f / (1.0f - h.p_drop) : 1.0f;
  unsigned long long seed = 0, roff = 0;
  if (drop) { seed = h.rng[0]; roff = h.rng[1]; }
  const Philox ph(seed);
  const float invB = 1.0f / (float)h.B;

  // h2 = relu(BN2(a2)) recomputed from a2; column coefficients of BN2 in s_c[0..1]
  for (int c = threadIdx.x; c < h.D2; c += HT) {
    s_c[c] = h.invstd2[c] * h.g2[c];
    s_c[ldp + c] = h.b2[c] - h.mean2[c] * h.invstd2[c] * h.g2[c];
  }
  __syncthreads();
  // ---- phase 0: dW3 = h2^T d_out ;  g2 = (d_out W3^T) * relu'(BN2(a2)) ----
  {
    const int mt = (h.D2 + TM - 1) / TM, nt = (h.NC + TN - 1) / TN;
    for (int t = blockIdx.x; t < mt * nt; t += nb) {
      const int m0 = (t / nt) * TM, n0 = (t % nt) * TN;
      tile_gemm(m0, n0, h.D2, h.NC, h.B,
                [&](int m, int k) { return fmaxf(fmaf(h.a2[(size_t)k * h.D2 + m], s_c[m], s_c[ldp + m]), 0.f); },
                [&](int k, int n) { return h.d_out[(size_t)k * h.NC + n]; },
                [&](int m, int n, float v) { h.dW3[(size_t)m * h.NC + n] = v; return make_float2(0.f, 0.f); }, sA, sB);
    }
    const int nt2 = (h.D2 + TN - 1) / TN;
    for (int t = blockIdx.x; t < rt * nt2; t += nb) {
      const int m0 = (t / nt2) * TM, n0 = (t % nt2) * TN;
      tile_gemm(m0, n0, h.B, h.D2, h.NC,
                [&](int m, int k) { return h.d_out[(size_t)m * h.NC + k]; },
                [&](int k, int n) { return h.W3[(size_t)n * h.NC + k]; },
                [&](int m, int n, float v) {
                  const float a = h.a2[(size_t)m * h.D2 + n];
                  const float z = fmaf(a, s_c[n], s_c[ldp + n]);
                  const float g = z > 0.f ? v : 0.f;
                  h.g2buf[(size_t)m * h.D2 + n] = g;
                  return make_float2(g, g * ((a - h.mean2[n]) * h.invstd2[n]));
                }, sA, sB, h.part, ldp);
    }
  }
  grid_sync(h.bar, nb);          // g2 and its (sum g, sum g*xhat) partials complete

  // ---- phase 1: da2 (on the fly) ; dW2 = h1^T da2 ; g1 = (da2 W2^T) * drop * relu'(BN1(a1)) ----
  // da2[m,k] = gam2*invstd2*(g2 - S1/B - xhat2*S2/B) [train] | gam2*invstd2*g2 [eval]  (+ d_a2 if given)
  for (int c = threadIdx.x; c < h.D2; c += HT) {
    float s1, s2;
    bwd_sums(h.part, rt, ldp, c, s1, s2);
    if (blockIdx.x == 0) { h.dg2[c] = s2; h.db2[c] = s1; }
    s_c[c] = h.g2[c] * h.invstd2[c];
    s_c[ldp + c] = h.training ? s1 * invB : 0.f;
    s_c[2 * ldp + c] = h.training ? s2 * invB : 0.f;
  }
  __syncthreads();
  auto da2 = [&](int m, int k) {
    const size_t i = (size_t)m * h.D2 + k;
    const float xh = (h.a2[i] - h.mean2[k]) * h.invstd2[k];
    float v = s_c[k] * (h.g2buf[i] - s_c[ldp + k] - xh * s_c[2 * ldp + k]);
    if (h.d_a2) v += h.d_a2[i];
    return v;
  };
  // BN1 forward coefficients for recomputing h1 live in s_c[3*ldp ..] (scale) and part-free registers (shift via mean/invstd)
  {
    const int mt = (h.D1 + TM - 1) / TM, nt = (h.D2 + TN - 1) / TN;
    for (int t = blockIdx.x; t < mt * nt; t += nb) {
      const int m0 = (t / nt) * TM, n0 = (t % nt) * TN;
      tile_gemm(m0, n0, h.D1, h.D2, h.B,
                [&](int m, int k) {                       // h1[k(row), m(col)]
                  const size_t i = (size_t)k * h.D1 + m;
                  const float sc = h.invstd1[m] * h.g1[m];
                  float v = fmaxf(fmaf(h.a1[i], sc, h.b1[m] - h.mean1[m] * sc), 0.f);
                  if (drop) v = dropout_keep(ph, roff, h.rng_stream, (unsigned long long)i, h.p_drop) ? v * dscale : 0.f;
                  return v;
                },
                [&](int k, int n) { return da2(k, n); },
                [&](int m, int n, float v) { h.dW2[(size_t)m * h.D2 + n] = v; return make_float2(0.f, 0.f); }, sA, sB);
    }
    const int nt1 = (h.D1 + TN - 1) / TN;
    for (int t = blockIdx.x; t < rt * nt1; t += nb) {
      const int m0 = (t / nt1) * TM, n0 = (t % nt1) * TN;
      tile_gemm(m0, n0, h.B, h.D1, h.D2,
                [&](int m, int k) { return da2(m, k); },
                [&](int k, int n) { return h.W2[(size_t)n * h.D2 + k]; },
                [&](int m, int n, float v) {
                  const size_t i = (size_t)m * h.D1 + n;
                  const float sc = h.invstd1[n] * h.g1[n];
                  const float a = h.a1[i];
                  const float z = fmaf(a, sc, h.b1[n] - h.mean1[n] * sc);
                  float g = z > 0.f ? v : 0.f;
                  if (drop) g = dropout_keep(ph, roff, h.rng_stream, (unsigned long long)i, h.p_drop) ? g * dscale : 0.f;
                  h.g1buf[i] = g;
                  return make_float2(g, g * ((a - h.mean1[n]) * h.invstd1[n]));
                }, sA, sB, h.part2, ldp);
    }
  }
  grid_sync(h.bar, nb);          // g1 and its partials (part2) complete

  // ---- phase 2: da1 (on the fly) ; dW1 = h0^T da1 ; dh0 = da1 W1^T ----
  for (int c = threadIdx.x; c < h.D1; c += HT) {
    float s1, s2;
    bwd_sums(h.part2, rt, ldp, c, s1, s2);
    if (blockIdx.x == 0) { h.dg1[c] = s2; h.db1[c] = s1; }
    s_c[c] = h.g1[c] * h.invstd1[c];
    s_c[ldp + c] = h.training ? s1 * invB : 0.f;
    s_c[2 * ldp + c] = h.training ? s2 * invB : 0.f;
  }
  __syncthreads();
  auto da1 = [&](int m, int k) {
    const size_t i = (size_t)m * h.D1 + k;
    const float xh = (h.a1[i] - h.mean1[k]) * h.invstd1[k];
    return s_c[k] * (h.g1buf[i] - s_c[ldp + k] - xh * s_c[2 * ldp + k]);
  };
  {
    const int mt = (h.F + TM - 1) / TM, nt = (h.D1 + TN - 1) / TN;
    for (int t = blockIdx.x; t < mt * nt; t += nb) {
      const int m0 = (t / nt) * TM, n0 = (t % nt) * TN;
      tile_gemm(m0, n0, h.F, h.D1, h.B,
                [&](int m, int k) {                       // h0[k(row), m(col)] = BN0(x0)
                  const float sc = h.invstd0[m] * h.g0[m];
                  return fmaf(h.x0[(size_t)k * h.F + m], sc, h.b0[m] - h.mean0[m] * sc);
                },
                [&](int k, int n) { return da1(k, n); },
                [&](int m, int n, float v) { h.dW1[(size_t)m * h.D1 + n] = v; return make_float2(0.f, 0.f); }, sA, sB);
    }
    const int ntf = (h.F + TN - 1) / TN;
    for (int t = blockIdx.x; t < rt * ntf; t += nb) {
      const int m0 = (t / ntf) * TM, n0 = (t % ntf) * TN;
      tile_gemm(m0, n0, h.B, h.F, h.D1,
                [&](int m, int k) { return da1(m, k); },
                [&](int k, int n) { return h.W1[(size_t)n * h.D1 + k]; },
                [&](int m, int n, float v) {
                  h.dh0[(size_t)m * h.F + n] = v;
                  return make_float2(v, v * ((h.x0[(size_t)m * h.F + n] - h.mean0[n]) * h.invstd0[n]));
                }, sA, sB, h.part, ldp);
    }
  }
  grid_sync(h.bar, nb);          // dh0 and its partials complete

  // ---- phase 3: dx0 = Graph_BN backward ----
  for (int c = threadIdx.x; c < h.F; c += HT) {
    float s1, s2;
    bwd_sums(h.part, rt, ldp, c, s1, s2);
    if (blockIdx.x == 0) { h.dg0[c] = s2; h.db0[c] = s1; }
    s_c[c] = h.g0[c] * h.invstd0[c];
    s_c[ldp + c] = h.training ? s1 * invB : 0.f;
    s_c[2 * ldp + c] = h.training ? s2 * invB : 0.f;
  }
  __syncthreads();
  const long long total = (long long)h.B * h.F;
  for (long long i = (long long)blockIdx.x * HT + threadIdx.x; i < total; i += (long long)nb * HT) {
    const int k = (int)(i % h.F);
    const float xh = (h.x0[i] - h.mean0[k]) * h.invstd0[k];
    h.dx0[i] = s_c[k] * (h.dh0[i] - s_c[ldp + k] - xh * s_c[2 * ldp + k]);
  }
}

}  // namespace head
}  // namespace eagcn

// ---- C ABI ------------------------------------------------------------------------------------------
using namespace eagcn::head;

static bool head_to_dev(const eagcn_head_t* a, HeadDev& h, bool bwd) {
  if (!a || a->B <= 0 || a->F <= 0 || a->D1 <= 0 || a->D2 <= 0 || a->NC <= 0) return false;
  h.B = (int)a->B; h.F = (int)a->F; h.D1 = (int)a->D1; h.D2 = (int)a->D2; h.NC = (int)a->NC;
  h.training = a->training ? 1 : 0; h.p_drop = (float)a->p_drop; h.eps = a->eps; h.momentum = a->momentum;
  h.x0 = (const float*)a->x0;
  h.g0 = (const float*)a->bn_w[0]; h.b0 = (const float*)a->bn_b[0];
  h.g1 = (const float*)a->bn_w[1]; h.b1 = (const float*)a->bn_b[1];
  h.g2 = (const float*)a->bn_w[2]; h.b2 = (const float*)a->bn_b[2];
  h.rm0 = (float*)a->bn_rm[0]; h.rv0 = (float*)a->bn_rv[0]; h.rm1 = (float*)a->bn_rm[1]; h.rv1 = (float*)a->bn_rv[1];
  h.rm2 = (float*)a->bn_rm[2]; h.rv2 = (float*)a->bn_rv[2];
  h.nbt0 = (long long*)a->bn_nbt[0]; h.nbt1 = (long long*)a->bn_nbt[1]; h.nbt2 = (long long*)a->bn_nbt[2];
  h.W1 = (const float*)a->W[0]; h.W2 = (const float*)a->W[1]; h.W3 = (const float*)a->W[2];
  h.rng = (const unsigned long long*)a->rng; h.rng_stream = (unsigned long long)a->rng_stream;
  h.a1 = (float*)a->a1; h.a2 = (float*)a->a2; h.out = (float*)a->out;
  h.mean0 = (float*)a->mean[0]; h.invstd0 = (float*)a->invstd[0]; h.mean1 = (float*)a->mean[1];
  h.invstd1 = (float*)a->invstd[1]; h.mean2 = (float*)a->mean[2]; h.invstd2 = (float*)a->invstd[2];
  h.part = (float*)a->part; h.bar = (unsigned*)a->bar;
  h.part2 = h.part ? h.part + eagcn_head_part_floats(a->B, a->F, a->D1, a->D2) / 2 : nullptr;
  h.d_out = (const float*)a->d_out; h.d_a2 = (const float*)a->d_a2;
  h.g2buf = (float*)a->g2buf; h.g1buf = (float*)a->g1buf; h.dh0 = (float*)a->dh0;
  h.dx0 = (float*)a->dx0; h.dW1 = (float*)a->dW[0]; h.dW2 = (float*)a->dW[1]; h.dW3 = (float*)a->dW[2];
  h.dg0 = (float*)a->dbn_w[0]; h.db0 = (float*)a->dbn_b[0]; h.dg1 = (float*)a->dbn_w[1]; h.db1 = (float*)a->dbn_b[1];
  h.dg2 = (float*)a->dbn_w[2]; h.db2 = (float*)a->dbn_b[2];
  bool ok = h.x0 && h.g0 && h.b0 && h.g1 && h.b1 && h.g2 && h.b2 && h.rm0 && h.rv0 && h.rm1 && h.rv1 && h.rm2 && h.rv2 &&
            h.W1 && h.W2 && h.W3 && h.a1 && h.a2 && h.mean0 && h.invstd0 && h.mean1 && h.invstd1 && h.mean2 &&
            h.invstd2 && h.part && h.bar;
  if (h.training && h.p_drop > 0.f && !h.rng) ok = false;
  if (h.p_drop < 0.f || h.p_drop >= 1.f) ok = false;
  if (!bwd) ok = ok && h.out;
  else ok = ok && h.d_out && h.g2buf && h.g1buf && h.dh0 && h.dx0 && h.dW1 && h.dW2 && h.dW3 && h.dg0 && h.db0 && h.dg1 &&
            h.db1 && h.dg2 && h.db2;
  return ok;
}

static int head_grid(const HeadDev& h) {
  const int rt = (h.B + TM - 1) / TM;
  int work = rt * ((h.F + TN - 1) / TN);                       // the widest phase (dh0 tiles)
  const int w2 = ((h.F + TM - 1) / TM) * ((h.D1 + TN - 1) / TN);
  if (w2 > work) work = w2;
  return work < 1 ? 1 : (work > kMaxCtas ? kMaxCtas : work);
}

extern "C" int64_t eagcn_head_part_floats(int64_t B, int64_t F, int64_t D1, int64_t D2) {
  const int64_t ld = F > D1 ? (F > D2 ? F : D2) : (D1 > D2 ? D1 : D2);
  return 2 * ((B + TM - 1) / TM) * 2 * ld;          // two ping-pong buffers
}

extern "C" int eagcn_head_forward(const eagcn_head_t* args, void* stream) {
  HeadDev h;
  if (!head_to_dev(args, h, false)) return EAGCN_E_ARG;
  const int ldp = h.F > h.D1 ? (h.F > h.D2 ? h.F : h.D2) : (h.D1 > h.D2 ? h.D1 : h.D2);
  const size_t smem = (size_t)2 * ldp * sizeof(float);
  if (smem > 160 * 1024) return EAGCN_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  if (e != cudaSuccess) return (int)e;
  EAGCN_PROF("head_fwd_kernel", stream);
  EAGCN_LAUNCH(head_fwd_kernel, head_grid(h), HT, smem, (cudaStream_t)stream)(h);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int eagcn_head_backward(const eagcn_head_t* args, void* stream) {
  HeadDev h;
  if (!head_to_dev(args, h, true)) return EAGCN_E_ARG;
  const int ldp = h.F > h.D1 ? (h.F > h.D2 ? h.F : h.D2) : (h.D1 > h.D2 ? h.D1 : h.D2);
  const size_t smem = (size_t)3 * ldp * sizeof(float);
  if (smem > 160 * 1024) return EAGCN_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  if (e != cudaSuccess) return (int)e;
  EAGCN_PROF("head_bwd_kernel", stream);
  EAGCN_LAUNCH(head_bwd_kernel, head_grid(h), HT, smem, (cudaStream_t)stream)(h);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
