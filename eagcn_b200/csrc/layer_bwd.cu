// Backward of one GraphConv_Layer on packed rows (replaces the autograd replay of reference
// layers.py:293-325 driven by train.py:333).  Nothing big is stored by forward except Z = H.W_all and the
// pre-BatchNorm Y; attention weights are recomputed from the uint8 edge codes and 1/R.
//
//   bnbwd-partial : g = dX * dropout * relu'  ;  per-channel  S1 = sum g,  S2 = sum g*xhat   (xhat = BN-normalised Y)
//   bnbwd-apply   : dY = gamma*invstd * (g - S1/M - xhat*S2/M)     (training; M = B*N padded positions)
//                   dY = gamma*invstd * g                           (eval: running statistics)
//                   Padded rows receive the constant dY_pad = -gamma*invstd*(S1/M + xhat_pad*S2/M); they only
//                   feed d(bias), and sum over ALL positions of dY is identically 0 in training, so
//                   d(bias) = 0 there (the reference's autograd returns rounding noise around 0).
//   agg-bwd       : with Y_v[i] = sum_j A_v[i,j] Z_v[j] + b,  A = U/R:
//                     dU[i,j]  = dY_v[i].(Z_v[j] - (Y_v[i]-b)) / R_i       (quotient rule; sum_j A[i,j] Z[j] = Y-b)
//                     d a_v[c] = sum_{edges of type c} dU * s(1-s),   d r_v = sum_i dU[i,i] * s_r(1-s_r)
//                     Q_v[j]   = sum_i A_v[i,j] dY_v[i]   (transposed aggregation through the reverse-edge codes)
//   gemm          : dH = Q . W_all^T ,  dW_all = H^T . Q   (split-K, fixed-order reduction)
#include "common.cuh"

namespace eagcn {

int gemm_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int Mcap, int N, int K,
            const int* Mdev, cudaStream_t st);
int gemm_tn(const float* A, int lda, const float* B, int ldb, float* C, int M, int N, int Kcap, const int* Kdev,
            float* ws, long long ws_floats, cudaStream_t st, int* nsplit_out);
long long gemm_tn_workspace_floats(int M, int N, int Kcap);
int splitk_reduce(const float* ws, float* C, long long n, int ns, cudaStream_t st);
bool layer_ok(const eagcn_plan_t* plan, const eagcn_layer_t* l);
struct StatEpilogue;

template <int VEC>
__device__ __forceinline__ void ld4(const float* __restrict__ row, int q, int lane, int lim, float (&o)[4]) {
  if (VEC == 4) {
    const int c = q * 128 + lane * 4;
    if (c < lim) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(row + c));
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else { o[0] = o[1] = o[2] = o[3] = 0.0f; }
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int c = q * 128 + u * 32 + lane; o[u] = c < lim ? __ldg(row + c) : 0.0f; }
  }
}
template <int VEC>
__device__ __forceinline__ void st4(float* __restrict__ row, int q, int lane, int lim, const float (&o)[4]) {
  if (VEC == 4) {
    const int c = q * 128 + lane * 4;
    if (c < lim) *reinterpret_cast<float4*>(row + c) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int c = q * 128 + u * 32 + lane; if (c < lim) row[c] = o[u]; }
  }
}

// g for one element (shared by partial / apply): dX * keep/(1-p) * [BN output > 0]
__device__ __forceinline__ float grad_through_act(float dx, float y, float mean, float invstd, float gamma, float beta,
                                                  int training, float p_drop, const unsigned long long* rng,
                                                  unsigned long long stream, unsigned long long idx, float& xhat) {
  xhat = (y - mean) * invstd;
  const float z = xhat * gamma + beta;
  float g = z > 0.0f ? dx : 0.0f;
  if (training && p_drop > 0.0f) {
    const Philox ph(rng[0]);
    g = dropout_keep(ph, rng[1], stream, idx, p_drop) ? g * (1.0f / (1.0f - p_drop)) : 0.0f;
  }
  return g;
}

// grid (row tiles, ceil(C/128)); thread -> channel, 64 rows serial per 2 half-tiles
__global__ void __launch_bounds__(256) bn_bwd_partial_kernel(PlanDev p, const float* __restrict__ dX,
                                                             const float* __restrict__ Y, const float* __restrict__ ball,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             float* __restrict__ partial, int C, int training, float p_drop,
                                                             const unsigned long long* rng, unsigned long long stream) {
  __shared__ float s[2][2][128];
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int tile = blockIdx.x;
  if (tile * kStatRows >= T) return;
  const int cl = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int c = blockIdx.y * 128 + cl;
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    const float mu = mean[c], is = invstd[c], g_ = ball[C + c], b_ = ball[2 * C + c];
    const int r0 = tile * kStatRows + half * (kStatRows / 2);
    const int r1 = min(T, r0 + kStatRows / 2);
    for (int t = r0; t < r1; ++t) {
      const long long idx = (long long)t * C + c;
      float xh;
      const float g = grad_through_act(dX[idx], Y[idx], mu, is, g_, b_, training, p_drop, rng, stream,
                                       (unsigned long long)idx, xh);
      s1 += g; s2 = fmaf(g, xh, s2);
    }
  }
  s[half][0][cl] = s1; s[half][1][cl] = s2;
  __syncthreads();
  if (half == 0 && c < C) {
    partial[((size_t)tile * 2 + 0) * C + c] = s[0][0][cl] + s[1][0][cl];
    partial[((size_t)tile * 2 + 1) * C + c] = s[0][1][cl] + s[1][1][cl];
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(PlanDev p, const float* __restrict__ dX,
                                                           const float* __restrict__ Y, const float* __restrict__ ball,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const double* __restrict__ bsums, float* __restrict__ dY, int C,
                                                           int training, float p_drop, const unsigned long long* rng,
                                                           unsigned long long stream, double M) {
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)p.t_cap * C) return;
  const int t = (int)(idx / C), c = (int)(idx - (long long)t * C);
  if (t >= T) { dY[idx] = 0.0f; return; }
  float xh;
  const float gam = ball[C + c];
  const float g = grad_through_act(dX[idx], Y[idx], mean[c], invstd[c], gam, ball[2 * C + c], training, p_drop, rng,
                                   stream, (unsigned long long)idx, xh);
  float r = g;
  if (training) r = g - (float)(bsums[c] / M) - xh * (float)(bsums[C + c] / M);
  dY[idx] = gam * invstd[c] * r;
}

// ---- float4 forms (C % 4 == 0) --------------------------------------------------------------------
__device__ __forceinline__ void grad_through_act4(const float4 dx, const float4 y, const float4 mu, const float4 is,
                                                  const float4 ga, const float4 be, bool drop, float scale, const Philox& ph,
                                                  unsigned long long off, unsigned long long stream,
                                                  unsigned long long idx, float p_drop, float (&g)[4], float (&xh)[4]) {
  xh[0] = (y.x - mu.x) * is.x; xh[1] = (y.y - mu.y) * is.y; xh[2] = (y.z - mu.z) * is.z; xh[3] = (y.w - mu.w) * is.w;
  g[0] = xh[0] * ga.x + be.x > 0.f ? dx.x : 0.f; g[1] = xh[1] * ga.y + be.y > 0.f ? dx.y : 0.f;
  g[2] = xh[2] * ga.z + be.z > 0.f ? dx.z : 0.f; g[3] = xh[3] * ga.w + be.w > 0.f ? dx.w : 0.f;
  if (drop) {
    bool k[4];
    dropout_keep4(ph, off, stream, idx, p_drop, k);
#pragma unroll
    for (int u = 0; u < 4; ++u) g[u] = k[u] ? g[u] * scale : 0.f;
  }
}

// grid (stat tiles, ceil(C/4/32)); block (32 channel-groups x 8 row-groups): a thread sums 1/8 of the tile's rows
// for its 4 channels, the 8 row-groups are then combined through shared memory in fixed order
__global__ void __launch_bounds__(256) bn_bwd_partial_vec_kernel(PlanDev p, const float* __restrict__ dX,
                                                                 const float* __restrict__ Y, const float* __restrict__ ball,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 float* __restrict__ partial, int C, int training, float p_drop,
                                                                 const unsigned long long* rng, unsigned long long stream) {
  __shared__ float4 s_red[2][8][32];
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int tile = blockIdx.x;
  if (tile * kStatRows >= T) return;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + cx) * 4;
  const bool act = c < C;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (act) {
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(ball + C + c), be = *reinterpret_cast<const float4*>(ball + 2 * C + c);
    const bool drop = training && p_drop > 0.0f;
    const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
    unsigned long long seed = 0, off = 0;
    if (drop) { seed = rng[0]; off = rng[1]; }
    const Philox ph(seed);
    const int r0 = tile * kStatRows + ry * (kStatRows / 8);
    const int r1 = min(T, r0 + kStatRows / 8);
#pragma unroll 4
    for (int t = r0; t < r1; ++t) {
      const size_t idx = (size_t)t * C + c;
      const float4 dx = __ldg(reinterpret_cast<const float4*>(dX + idx)), y = __ldg(reinterpret_cast<const float4*>(Y + idx));
      float g[4], xh[4];
      grad_through_act4(dx, y, mu, is, ga, be, drop, scale, ph, off, stream, (unsigned long long)idx, p_drop, g, xh);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s1[u] += g[u]; s2[u] = fmaf(g[u], xh[u], s2[u]); }
    }
  }
  s_red[0][ry][cx] = make_float4(s1[0], s1[1], s1[2], s1[3]);
  s_red[1][ry][cx] = make_float4(s2[0], s2[1], s2[2], s2[3]);
  __syncthreads();
  if (ry < 2 && act) {
    float4 a = s_red[ry][0][cx];
#pragma unroll
    for (int w = 1; w < 8; ++w) { const float4 b = s_red[ry][w][cx]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
    *reinterpret_cast<float4*>(partial + ((size_t)tile * 2 + ry) * C + c) = a;
  }
}

// grid (ceil(C/4/128), ceil(t_cap/kEltRows))
__global__ void __launch_bounds__(128) bn_bwd_apply_vec_kernel(PlanDev p, const float* __restrict__ dX,
                                                               const float* __restrict__ Y, const float* __restrict__ ball,
                                                               const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               const double* __restrict__ bsums, float* __restrict__ dY, int C,
                                                               int training, float p_drop, const unsigned long long* rng,
                                                               unsigned long long stream, double M) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (c >= C) return;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
  const float4 ga = *reinterpret_cast<const float4*>(ball + C + c), be = *reinterpret_cast<const float4*>(ball + 2 * C + c);
  const bool drop = training && p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  unsigned long long seed = 0, off = 0;
  if (drop) { seed = rng[0]; off = rng[1]; }
  const Philox ph(seed);
  float m1[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
  if (training) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { m1[u] = (float)(bsums[c + u] / M); m2[u] = (float)(bsums[C + c + u] / M); }
  }
  const float gi[4] = {ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w};
  const int r0 = blockIdx.y * kEltRows, r1 = min(p.t_cap, r0 + kEltRows);
#pragma unroll 4
  for (int t = r0; t < r1; ++t) {
    const size_t idx = (size_t)t * C + c;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T) {
      const float4 dx = __ldg(reinterpret_cast<const float4*>(dX + idx)), y = __ldg(reinterpret_cast<const float4*>(Y + idx));
      float g[4], xh[4];
      grad_through_act4(dx, y, mu, is, ga, be, drop, scale, ph, off, stream, (unsigned long long)idx, p_drop, g, xh);
      float r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = gi[u] * (training ? (g[u] - m1[u] - xh[u] * m2[u]) : g[u]);
      o = make_float4(r[0], r[1], r[2], r[3]);
    }
    *reinterpret_cast<float4*>(dY + idx) = o;
  }
}

// dvec = [dbias | dgamma | dbeta]
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ ball, const float* __restrict__ invstd,
                                                              const double* __restrict__ bsums, float* __restrict__ dvec,
                                                              int C, int training) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const double s1 = bsums[c], s2 = bsums[C + c];
  dvec[c] = training ? 0.0f : (float)((double)ball[C + c] * (double)invstd[c] * s1);
  dvec[C + c] = (float)s2;
  dvec[2 * C + c] = (float)s1;
}

// grid (row tiles, V); kAggWarps warps x kAggRows rows.  Fo_v handled in super-chunks of 512 channels (4 x 128).
template <int VEC>
__global__ void __launch_bounds__(kAggThreads) agg_bwd_kernel(PlanDev p, LayerDev L, const float* __restrict__ Z,
                                                      const float* __restrict__ Y, const float* __restrict__ dY,
                                                      const float* __restrict__ ball, const float* __restrict__ sig,
                                                      const float* __restrict__ invR, float* __restrict__ Q,
                                                      float* __restrict__ dpart) {
  __shared__ float s_hist[kAggWarps][EAGCN_SIG_STRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.y, tile = blockIdx.x;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  {   // slack rows [T, t_cap) of Q stay defined (zeros): the split-K dW GEMM reads whole 32-row blocks
    const int fo_ = L.fo[v], off_ = L.off[v];
    for (int r = 0; r < kAggRows; ++r) {
      const int t = tile * kStatRows + warp * kAggRows + r;
      if (t >= T && t < p.t_cap)
        for (int c = lane; c < fo_; c += 32) Q[(size_t)t * L.fo_tot + off_ + c] = 0.0f;
    }
  }
  if (tile * kStatRows >= T) return;
  for (int i = threadIdx.x; i < kAggWarps * EAGCN_SIG_STRIDE; i += kAggThreads) (&s_hist[0][0])[i] = 0.0f;
  __syncthreads();
  const int fo = L.fo[v], off = L.off[v], ld = L.fo_tot;
  const float* sg = sig + v * EAGCN_SIG_STRIDE;
  const float sig_r = sg[256];
  const uint8_t* code = p.code + (size_t)v * p.e_cap;
  const uint8_t* rcode = p.rcode + (size_t)v * p.e_cap;
  const float* iR = invR + (size_t)v * p.t_cap;
  const int nsc = (fo + 511) / 512;
  for (int r = 0; r < kAggRows; ++r) {
    const int t = tile * kStatRows + warp * kAggRows + r;
    if (t >= T) break;
    const int e0 = p.row_ptr[t], e1 = p.row_ptr[t + 1];
    const float invR_t = iR[t];
    const float* dYt = dY + (size_t)t * ld + off;
    // ---- c_t = dY_t . (Y_t - b),  d_self = dY_t . Z_t   (full Fo_v dots) ----
    float ct = 0.f, dself = 0.f;
    for (int q = 0; q < (fo + 127) / 128; ++q) {
      float a[4], y[4], b[4], z[4];
      ld4<VEC>(dYt, q, lane, fo, a);
      ld4<VEC>(Y + (size_t)t * ld + off, q, lane, fo, y);
      ld4<VEC>(ball + off, q, lane, fo, b);
      ld4<VEC>(Z + (size_t)t * ld + off, q, lane, fo, z);
#pragma unroll
      for (int u = 0; u < 4; ++u) { ct = fmaf(a[u], y[u] - b[u], ct); dself = fmaf(a[u], z[u], dself); }
    }
    ct = warp_sum(ct); dself = warp_sum(dself);
    if (lane == 0) s_hist[warp][256] += (dself - ct) * invR_t * sig_r * (1.0f - sig_r);
    const float a_self = sig_r * invR_t;
    // ---- edges, 32 at a time ----
    for (int eb = e0; eb < e1; eb += 32) {   // active rows have deg >= 1
      const int e = eb + lane;
      int j_e = 0, c_e = 0; float aq_e = 0.f, s_e = 0.f, d_e = 0.f;
      if (e < e1) {
        j_e = p.col[e]; c_e = code[e];
        s_e = sg[c_e];
        aq_e = sg[rcode[e]] * iR[j_e];               // A_v[j -> t]: edge (j,t) has type rcode, row sum R_j
      }
      const int cnt = min(32, e1 - eb);
      for (int sc = 0; sc < nsc; ++sc) {
        float dyt[4][4], qacc[4][4];
        const int nq = min(4, (fo - sc * 512 + 127) / 128);
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          if (qq < nq) {
            ld4<VEC>(dYt, sc * 4 + qq, lane, fo, dyt[qq]);
#pragma unroll
            for (int u = 0; u < 4; ++u) qacc[qq][u] = 0.0f;
          }
        }
#pragma unroll 2
        for (int k = 0; k < cnt; ++k) {
          const int j = __shfl_sync(0xffffffffu, j_e, k);
          const float aq = __shfl_sync(0xffffffffu, aq_e, k);
          float dot = 0.0f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            if (qq < nq) {
              float zj[4], gj[4];
              ld4<VEC>(Z + (size_t)j * ld + off, sc * 4 + qq, lane, fo, zj);
              ld4<VEC>(dY + (size_t)j * ld + off, sc * 4 + qq, lane, fo, gj);
#pragma unroll
              for (int u = 0; u < 4; ++u) { dot = fmaf(dyt[qq][u], zj[u], dot); qacc[qq][u] = fmaf(aq, gj[u], qacc[qq][u]); }
            }
          }
          dot = warp_sum(dot);
          if (lane == k) d_e += dot;
        }
        // Q rows: first edge chunk initialises with the self term, later chunks accumulate
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          if (qq < nq) {
            float* qrow = Q + (size_t)t * ld + off;
            float base[4];
            if (eb == e0) {
#pragma unroll
              for (int u = 0; u < 4; ++u) base[u] = a_self * dyt[qq][u];
            } else {
              ld4<VEC>(qrow, sc * 4 + qq, lane, fo, base);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) base[u] += qacc[qq][u];
            st4<VEC>(qrow, sc * 4 + qq, lane, fo, base);
          }
        }
      }
      // attention-logit gradients of this chunk's edges, serialised -> deterministic
      const float ds_e = (d_e - ct) * invR_t * s_e * (1.0f - s_e);
      for (int k = 0; k < cnt; ++k) {
        const float dsk = __shfl_sync(0xffffffffu, ds_e, k);
        const int ck = __shfl_sync(0xffffffffu, c_e, k);
        if (lane == 0) s_hist[warp][ck] += dsk;
      }
    }
  }
  __syncthreads();
  float* out = dpart + ((size_t)tile * L.V + v) * EAGCN_SIG_STRIDE;
  for (int i = threadIdx.x; i < EAGCN_SIG_STRIDE; i += kAggThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kAggWarps; ++w) a += s_hist[w][i];
    out[i] = a;
  }
}

// ---- shared-memory tile variant with the BatchNorm/ReLU/dropout backward folded in -----------------------------
struct BnBwdArgs {
  const float* dX; const float* mean; const float* invstd; const double* bsums;
  const unsigned long long* rng; unsigned long long stream; double M; float p_drop; int training;
};
inline size_t agg_bwd_tile_smem(int fo) { return (size_t)(2 * kStatRows * fo + 6 * fo) * sizeof(float); }

// dY of 4 consecutive channels (c4-th float4 of the view slab) of row t: exactly bn_bwd_apply_vec_kernel's arithmetic.
// sP = per-view parameter rows [mean | invstd | gamma | beta | S1/M | S2/M], each nc4 float4 long.
__device__ __forceinline__ float4 dy_at(const BnBwdArgs& bn, const float* __restrict__ Y, const float4* sP, int nc4, int ld,
                                        int off, int t, int c4, bool drop, float scale, const Philox& ph,
                                        unsigned long long rng_off) {
  const size_t idx = (size_t)t * ld + off + c4 * 4;
  const float4 dx = __ldg(reinterpret_cast<const float4*>(bn.dX + idx)), y = __ldg(reinterpret_cast<const float4*>(Y + idx));
  const float4 mu = sP[c4], is = sP[nc4 + c4], ga = sP[2 * nc4 + c4], be = sP[3 * nc4 + c4];
  const float4 m1 = sP[4 * nc4 + c4], m2 = sP[5 * nc4 + c4];
  float g[4], xh[4];
  grad_through_act4(dx, y, mu, is, ga, be, drop, scale, ph, rng_off, bn.stream, (unsigned long long)idx, bn.p_drop, g, xh);
  const float gi[4] = {ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w};
  const float a1[4] = {m1.x, m1.y, m1.z, m1.w}, a2[4] = {m2.x, m2.y, m2.z, m2.w};
  float r[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) r[u] = gi[u] * (bn.training ? (g[u] - a1[u] - xh[u] * a2[u]) : g[u]);
  return make_float4(r[0], r[1], r[2], r[3]);
}

// out-of-line copy for the rare neighbour outside the tile: keeps the Philox state out of the row loop's registers
__device__ __noinline__ float4 dy_at_halo(const BnBwdArgs& bn, const float* __restrict__ Y, const float4* sP, int nc4, int ld,
                                          int off, int t, int c4) {
  const bool drop = bn.training && bn.p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - bn.p_drop) : 1.0f;
  unsigned long long seed = 0, rng_off = 0;
  if (drop) { seed = bn.rng[0]; rng_off = bn.rng[1]; }
  const Philox ph(seed);
  return dy_at(bn, Y, sP, nc4, ld, off, t, c4, drop, scale, ph, rng_off);
}

// dY slab of a tile -> shared memory (out of line: its Philox/parameter registers do not add to the row loop's)
__device__ __noinline__ void stage_dy_slab(const BnBwdArgs& bn, const float* __restrict__ Y, const float4* sP, float4* sG,
                                           int nc4, int ld, int off, int t0, int nrows) {
  const bool drop = bn.training && bn.p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - bn.p_drop) : 1.0f;
  unsigned long long seed = 0, rng_off = 0;
  if (drop) { seed = bn.rng[0]; rng_off = bn.rng[1]; }
  const Philox ph(seed);
#pragma unroll 2
  for (int i = threadIdx.x; i < nrows * nc4; i += kAggThreads) {
    const int r = i / nc4, c = i - r * nc4;
    sG[i] = dy_at(bn, Y, sP, nc4, ld, off, t0 + r, c, drop, scale, ph, rng_off);
  }
}

// grid (row tiles, V); fo_v <= 128*NQ, float4 layout.  The CTA builds its kStatRows x fo_v slabs of Z (cp.async) and of
// dY (computed from dX, Y and the reduced BatchNorm sums while staging -- bn_bwd_apply and the dY round trip through
// memory disappear) in shared memory, prefetches the edge metadata of each warp's rows (one edge per lane), and then
// runs the row loop of agg_bwd_kernel out of shared memory.  Neighbours outside the tile recompute dY on the fly.
template <int NQ>
__global__ void __launch_bounds__(kAggThreads, 3) agg_bwd_tile_kernel(PlanDev p, LayerDev L, BnBwdArgs bn,
                                                                   const float* __restrict__ Z, const float* __restrict__ Y,
                                                                   const float* __restrict__ ball,
                                                                   const float* __restrict__ sig,
                                                                   const float* __restrict__ invR, float* __restrict__ Q,
                                                                   float* __restrict__ dpart) {
  extern __shared__ __align__(16) float tile_smem_b[];
  __shared__ float s_hist[kAggWarps][EAGCN_SIG_STRIDE];
  __shared__ int s_rp[kStatRows + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v = blockIdx.y, tile = blockIdx.x, t0 = tile * kStatRows;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int fo = L.fo[v], off = L.off[v], ld = L.fo_tot, nc4 = fo >> 2;
  for (int r = 0; r < kAggRows; ++r) {   // slack rows [T, t_cap) of Q stay defined (zeros) for the split-K dW GEMM
    const int t = t0 + warp * kAggRows + r;
    if (t >= T && t < p.t_cap)
      for (int c = lane; c < nc4; c += 32) reinterpret_cast<float4*>(Q + (size_t)t * ld + off)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (t0 >= T) return;
  const int nrows = min(kStatRows, T - t0);
  float4* sZ = reinterpret_cast<float4*>(tile_smem_b);   // [kStatRows][nc4]
  float4* sG = sZ + kStatRows * nc4;                     // [kStatRows][nc4]  dY
  float4* sP = sG + kStatRows * nc4;                     // [6][nc4]
  for (int i = tid; i < nrows * nc4; i += kAggThreads) {
    const int r = i / nc4, c = i - r * nc4;
    cp_async16(sZ + i, Z + (size_t)(t0 + r) * ld + off + c * 4);
  }
  cp_async_commit();
  {
    float* P = reinterpret_cast<float*>(sP);
    for (int c = tid; c < fo; c += kAggThreads) {
      P[c] = bn.mean[off + c]; P[fo + c] = bn.invstd[off + c];
      P[2 * fo + c] = ball[ld + off + c]; P[3 * fo + c] = ball[2 * ld + off + c];
      P[4 * fo + c] = bn.training ? (float)(bn.bsums[off + c] / bn.M) : 0.0f;
      P[5 * fo + c] = bn.training ? (float)(bn.bsums[ld + off + c] / bn.M) : 0.0f;
    }
  }
  if (tid <= nrows) s_rp[tid] = p.row_ptr[t0 + tid];
  for (int i = tid; i < kAggWarps * EAGCN_SIG_STRIDE; i += kAggThreads) (&s_hist[0][0])[i] = 0.0f;
  __syncthreads();
  const float* sg = sig + v * EAGCN_SIG_STRIDE;
  const float sig_r = sg[256];
  const uint8_t* code = p.code + (size_t)v * p.e_cap;
  const uint8_t* rcode = p.rcode + (size_t)v * p.e_cap;
  const float* iR = invR + (size_t)v * p.t_cap;
  // edge metadata of this warp's rows: consecutive in the CSR arrays, one edge per lane when there are <= 32
  const int wr0 = warp * kAggRows, wr1 = min(nrows, wr0 + kAggRows);
  int ea = 0, eb = 0;
  if (wr0 < nrows) { ea = s_rp[wr0]; eb = s_rp[wr1]; }
  const bool pre = eb - ea <= 32;
  const bool mine = pre && ea + lane < eb;
  int j_p = 0, c_p = 0, rc_p = 0;
  if (mine) { j_p = p.col[ea + lane]; c_p = code[ea + lane]; rc_p = rcode[ea + lane]; }
  const float invR_l = wr0 + lane < wr1 ? iR[t0 + wr0 + lane] : 0.0f;
  stage_dy_slab(bn, Y, sP, sG, nc4, ld, off, t0, nrows);
  float s_p = 0.f, aq_p = 0.f;
  if (mine) { s_p = sg[c_p]; aq_p = sg[rc_p] * iR[j_p]; }
  cp_async_wait_all();
  __syncthreads();
  for (int r = wr0; r < wr1; ++r) {
    const int t = t0 + r;
    const int e0 = s_rp[r], e1 = s_rp[r + 1];
    const float invR_t = __shfl_sync(0xffffffffu, invR_l, r - wr0);
    // ---- c_t = dY_t . (Y_t - b),  d_self = dY_t . Z_t ----
    float dyt[NQ][4];
    float ct = 0.f, dself = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c4 = q * 32 + lane;
      if (c4 < nc4) {
        const float4 a = sG[r * nc4 + c4], z = sZ[r * nc4 + c4];
        const float4 y = __ldg(reinterpret_cast<const float4*>(Y + (size_t)t * ld + off) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(ball + off) + c4);
        dyt[q][0] = a.x; dyt[q][1] = a.y; dyt[q][2] = a.z; dyt[q][3] = a.w;
        ct = fmaf(a.x, y.x - b.x, ct); dself = fmaf(a.x, z.x, dself);
        ct = fmaf(a.y, y.y - b.y, ct); dself = fmaf(a.y, z.y, dself);
        ct = fmaf(a.z, y.z - b.z, ct); dself = fmaf(a.z, z.z, dself);
        ct = fmaf(a.w, y.w - b.w, ct); dself = fmaf(a.w, z.w, dself);
      } else { dyt[q][0] = dyt[q][1] = dyt[q][2] = dyt[q][3] = 0.0f; }
    }
    ct = warp_sum(ct); dself = warp_sum(dself);
    if (lane == 0) s_hist[warp][256] += (dself - ct) * invR_t * sig_r * (1.0f - sig_r);
    const float a_self = sig_r * invR_t;
    float qacc[NQ][4];
#pragma unroll
    for (int q = 0; q < NQ; ++q) { qacc[q][0] = qacc[q][1] = qacc[q][2] = qacc[q][3] = 0.0f; }
    for (int ch = e0; ch < e1; ch += 32) {   // active rows have deg >= 1
      int j_e = j_p, c_e = c_p, lo = e0 - ea;
      float aq_e = aq_p, s_e = s_p;
      if (!pre) {
        const int e = ch + lane;
        lo = 0; j_e = 0; c_e = 0; aq_e = 0.f; s_e = 0.f;
        if (e < e1) { j_e = p.col[e]; c_e = code[e]; s_e = sg[c_e]; aq_e = sg[rcode[e]] * iR[j_e]; }
      }
      const int cnt = min(32, e1 - ch);
      float d_e = 0.f;
      for (int k = 0; k < cnt; ++k) {
        const int j = __shfl_sync(0xffffffffu, j_e, lo + k);
        const float aq = __shfl_sync(0xffffffffu, aq_e, lo + k);   // A_v[j -> t]: edge (j,t) has type rcode, row sum R_j
        const int jr = j - t0;
        const bool in_tile = (unsigned)jr < (unsigned)nrows;
        float dot = 0.0f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int c4 = q * 32 + lane;
          if (c4 < nc4) {
            float4 zj, gj;
            if (in_tile) { zj = sZ[jr * nc4 + c4]; gj = sG[jr * nc4 + c4]; }
            else {
              zj = __ldg(reinterpret_cast<const float4*>(Z + (size_t)j * ld + off) + c4);
              gj = dy_at_halo(bn, Y, sP, nc4, ld, off, j, c4);
            }
            dot = fmaf(dyt[q][0], zj.x, dot); qacc[q][0] = fmaf(aq, gj.x, qacc[q][0]);
            dot = fmaf(dyt[q][1], zj.y, dot); qacc[q][1] = fmaf(aq, gj.y, qacc[q][1]);
            dot = fmaf(dyt[q][2], zj.z, dot); qacc[q][2] = fmaf(aq, gj.z, qacc[q][2]);
            dot = fmaf(dyt[q][3], zj.w, dot); qacc[q][3] = fmaf(aq, gj.w, qacc[q][3]);
          }
        }
        dot = warp_sum(dot);
        if (lane == lo + k) d_e += dot;
      }
      // attention-logit gradients of this chunk's edges, serialised -> deterministic
      const float ds_e = (d_e - ct) * invR_t * s_e * (1.0f - s_e);
      for (int k = 0; k < cnt; ++k) {
        const float dsk = __shfl_sync(0xffffffffu, ds_e, lo + k);
        const int ck = __shfl_sync(0xffffffffu, c_e, lo + k);
        if (lane == 0) s_hist[warp][ck] += dsk;
      }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c4 = q * 32 + lane;
      if (c4 < nc4)
        reinterpret_cast<float4*>(Q + (size_t)t * ld + off)[c4] =
            make_float4(a_self * dyt[q][0] + qacc[q][0], a_self * dyt[q][1] + qacc[q][1],
                        a_self * dyt[q][2] + qacc[q][2], a_self * dyt[q][3] + qacc[q][3]);
    }
  }
  __syncthreads();
  float* out = dpart + ((size_t)tile * L.V + v) * EAGCN_SIG_STRIDE;
  for (int i = tid; i < EAGCN_SIG_STRIDE; i += kAggThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kAggWarps; ++w) a += s_hist[w][i];
    out[i] = a;
  }
}

template <int NQ>
static int launch_agg_bwd_tile(dim3 grid, size_t smem, cudaStream_t st, const PlanDev& p, const LayerDev& L,
                               const BnBwdArgs& bn, const eagcn_work_t* w) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(agg_bwd_tile_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)agg_bwd_tile_smem(128 * NQ));
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  agg_bwd_tile_kernel<NQ><<<grid, kAggThreads, smem, st>>>(p, L, bn, (const float*)w->Z, (const float*)w->Y,
                                                           (const float*)w->ball, (const float*)w->sig,
                                                           (const float*)w->invR, (float*)w->Q, (float*)w->partial);
  return 0;
}

// datt[v][i] = sum over live tiles (fixed order)
__global__ void __launch_bounds__(256) datt_reduce_kernel(PlanDev p, const float* __restrict__ dpart,
                                                          float* __restrict__ datt, int V) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= V * EAGCN_SIG_STRIDE) return;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int ntile = (T + kStatRows - 1) / kStatRows;
  double a = 0.0;
  for (int t = 0; t < ntile; ++t) a += (double)dpart[(size_t)t * V * EAGCN_SIG_STRIDE + i];
  datt[i] = (float)a;
}

// One launch finishing the layer backward:
//  (a) dW: sum the split-K partials [ns][fin][C] in z order and store them VIEW-BLOCKED -- view v's [fin, fo_v]
//      block contiguous at offset fin*off_v -- so that each GraphConv_block's weight gradient is a contiguous tensor;
//  (b) d att / d self_r: sum the per-tile partials in tile order.
__global__ void __launch_bounds__(256) bwd_post_kernel(PlanDev p, LayerDev L, const float* __restrict__ wpart, int ns,
                                                       float* __restrict__ dwall, const float* __restrict__ dpart,
                                                       float* __restrict__ datt, int nblk_w) {
  const int C = L.fo_tot;
  if ((int)blockIdx.x < nblk_w) {
    const long long n = (long long)L.fin * C;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    float s = 0.0f;
#pragma unroll 4
    for (int z = 0; z < ns; ++z) s += __ldg(wpart + (long long)z * n + idx);
    const int k = (int)(idx / C), c = (int)(idx - (long long)k * C);
    int v = 0;
    while (v + 1 < L.V && c >= L.off[v + 1]) ++v;
    dwall[(long long)L.fin * L.off[v] + (long long)k * L.fo[v] + (c - L.off[v])] = s;
  } else {
    // one warp per table entry: lanes take every 32nd tile (independent loads), fixed-order shuffle tree
    const int i = ((int)blockIdx.x - nblk_w) * 8 + (threadIdx.x >> 5);
    if (i >= L.V * EAGCN_SIG_STRIDE) return;
    const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
    const int ntile = (T + kStatRows - 1) / kStatRows;
    double a = 0.0;
#pragma unroll 4
    for (int t = threadIdx.x & 31; t < ntile; t += 32) a += (double)__ldg(dpart + (size_t)t * L.V * EAGCN_SIG_STRIDE + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) datt[i] = (float)a;
  }
}

static bool vec4_ok_b(const eagcn_layer_t* l) {
  if (l->fo_tot % 4) return false;
  for (int v = 0; v < l->V; ++v) if ((l->fo[v] % 4) || (l->off[v] % 4)) return false;
  return true;
}

}  // namespace eagcn
using namespace eagcn;

extern "C" int64_t eagcn_gemm_workspace_bytes(int64_t fin, int64_t fo_tot, int64_t t_cap) {
  return gemm_tn_workspace_floats((int)fin, (int)fo_tot, (int)t_cap) * (int64_t)sizeof(float);
}
extern "C" int64_t eagcn_partial_floats(int64_t t_cap, int64_t fo_tot, int64_t V) {
  const int64_t tiles = eagcn_stat_tiles(t_cap);
  const int64_t a = tiles * 2 * fo_tot, b = tiles * V * EAGCN_SIG_STRIDE;
  return a > b ? a : b;
}

extern "C" int eagcn_layer_backward_a(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w,
                                      void* stream) {
  if (!plan_ok(plan) || !w || !layer_ok(plan, layer)) return EAGCN_E_ARG;
  if (!w->dX || !w->Y || !w->ball || !w->mean || !w->invstd || !w->partial || !w->bsums) return EAGCN_E_ARG;
  if ((w->training & 1) && w->p_drop > 0.0 && !w->rng) return EAGCN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PlanDev p = to_dev(plan);
  const int C = (int)layer->fo_tot;
  if ((C & 3) == 0 && aligned16(w->dX) && aligned16(w->Y)) {
    dim3 grid((unsigned)eagcn_stat_tiles(p.t_cap), (C / 4 + 31) / 32);
    EAGCN_PROF("bn_bwd_partial_kernel", st);
    bn_bwd_partial_vec_kernel<<<grid, 256, 0, st>>>(p, (const float*)w->dX, (const float*)w->Y, (const float*)w->ball,
                                                    (const float*)w->mean, (const float*)w->invstd, (float*)w->partial,
                                                    C, (w->training & 1) ? 1 : 0, (float)w->p_drop,
                                                    (const unsigned long long*)w->rng,
                                                    (unsigned long long)w->rng_stream);
    EAGCN_LAUNCH_CHECK();
  } else {
    dim3 grid((unsigned)eagcn_stat_tiles(p.t_cap), (C + 127) / 128);
    EAGCN_PROF("bn_bwd_partial_kernel", st);
    bn_bwd_partial_kernel<<<grid, 256, 0, st>>>(p, (const float*)w->dX, (const float*)w->Y, (const float*)w->ball,
                                                (const float*)w->mean, (const float*)w->invstd, (float*)w->partial, C,
                                                (w->training & 1) ? 1 : 0, (float)w->p_drop,
                                                (const unsigned long long*)w->rng, (unsigned long long)w->rng_stream);
    EAGCN_LAUNCH_CHECK();
  }
  // reduce the tile partials and emit dbias / dgamma / dbeta from THIS rank's sums: under data parallelism the
  // host all-reduces `bsums` afterwards (global-batch BatchNorm), but parameter gradients stay per-rank
  // contributions -- the flat gradient all-reduce sums them exactly once
  if (!w->dvec) return EAGCN_E_ARG;
  LayerDev L = to_dev(layer, plan);
  StatEpilogue ep{2, (const float*)w->ball, nullptr, (float*)w->invstd, (float*)w->dvec,
                  (w->training & 1) ? 1 : 0, 0.0, 0.0, 0.0};
  EAGCN_PROF("stat_reduce_kernel", st);
  stat_reduce_kernel<<<(C + 31) / 32, 32 * kStatLanes, 0, st>>>(p, L, (const float*)w->partial, (double*)w->bsums, C, ep);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int eagcn_layer_backward_b(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w,
                                      void* stream) {
  if (!plan_ok(plan) || !w || !layer_ok(plan, layer)) return EAGCN_E_ARG;
  if (!w->dX || !w->Y || !w->Z || !w->H || !w->ball || !w->mean || !w->invstd || !w->partial || !w->bsums || !w->dY ||
      !w->Q || !w->dwall || !w->dvec || !w->datt || !w->wall || !w->sig || !w->invR || !w->gemm_ws)
    return EAGCN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PlanDev p = to_dev(plan);
  LayerDev L = to_dev(layer, plan);
  const int C = L.fo_tot;
  const double M = (double)(w->m_total > 0 ? w->m_total : plan->B * plan->N);
  const long long total = (long long)p.t_cap * C;
  dim3 grid((unsigned)eagcn_stat_tiles(p.t_cap), L.V);
  int fo_max = 0;
  for (int v = 0; v < L.V; ++v) fo_max = L.fo[v] > fo_max ? L.fo[v] : fo_max;
  if (agg_mode() == 0 && vec4_ok_b(layer) && fo_max <= kTileMaxFo && aligned16(w->dX) && aligned16(w->Y) &&
      aligned16(w->Z) && aligned16(w->Q) && aligned16(w->ball)) {
    // fused: dY is produced inside the aggregation kernel's shared-memory tile (w->dY is not written)
    BnBwdArgs bn{(const float*)w->dX, (const float*)w->mean, (const float*)w->invstd, (const double*)w->bsums,
                 (const unsigned long long*)w->rng, (unsigned long long)w->rng_stream, M, (float)w->p_drop,
                 (w->training & 1) ? 1 : 0};
    const size_t smem = agg_bwd_tile_smem(fo_max);
    int lrc;
    EAGCN_PROF("agg_bwd_kernel", st);
    if (fo_max <= 128) lrc = launch_agg_bwd_tile<1>(grid, smem, st, p, L, bn, w);
    else if (fo_max <= 256) lrc = launch_agg_bwd_tile<2>(grid, smem, st, p, L, bn, w);
    else if (fo_max <= 384) lrc = launch_agg_bwd_tile<3>(grid, smem, st, p, L, bn, w);
    else lrc = launch_agg_bwd_tile<4>(grid, smem, st, p, L, bn, w);
    if (lrc) { ::eagcn::prof_end(); return lrc; }
  } else {
  if ((C & 3) == 0 && aligned16(w->dX) && aligned16(w->Y) && aligned16(w->dY)) {
    dim3 grid((C / 4 + 127) / 128, (p.t_cap + kEltRows - 1) / kEltRows);
    EAGCN_PROF("bn_bwd_apply_kernel", st);
    bn_bwd_apply_vec_kernel<<<grid, 128, 0, st>>>(
        p, (const float*)w->dX, (const float*)w->Y, (const float*)w->ball, (const float*)w->mean,
        (const float*)w->invstd, (const double*)w->bsums, (float*)w->dY, C, (w->training & 1) ? 1 : 0, (float)w->p_drop,
        (const unsigned long long*)w->rng, (unsigned long long)w->rng_stream, M);
    EAGCN_LAUNCH_CHECK();
  } else {
    EAGCN_PROF("bn_bwd_apply_kernel", st);
    bn_bwd_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        p, (const float*)w->dX, (const float*)w->Y, (const float*)w->ball, (const float*)w->mean, (const float*)w->invstd,
        (const double*)w->bsums, (float*)w->dY, C, (w->training & 1) ? 1 : 0, (float)w->p_drop,
        (const unsigned long long*)w->rng, (unsigned long long)w->rng_stream, M);
    EAGCN_LAUNCH_CHECK();
  }
  if (vec4_ok_b(layer)) {
    EAGCN_PROF("agg_bwd_kernel", st);
    agg_bwd_kernel<4><<<grid, kAggThreads, 0, st>>>(p, L, (const float*)w->Z, (const float*)w->Y, (const float*)w->dY,
                                            (const float*)w->ball, (const float*)w->sig, (const float*)w->invR,
                                            (float*)w->Q, (float*)w->partial);
  } else {
    EAGCN_PROF("agg_bwd_kernel", st);
    agg_bwd_kernel<1><<<grid, kAggThreads, 0, st>>>(p, L, (const float*)w->Z, (const float*)w->Y, (const float*)w->dY,
                                            (const float*)w->ball, (const float*)w->sig, (const float*)w->invR,
                                            (float*)w->Q, (float*)w->partial);
  }
  }
  EAGCN_LAUNCH_CHECK();
  int rc = 0;
  if (!w->dH) {
    // the layer input needs no gradient (first layer fed by data): skip dH = Q W^T altogether
  } else if (gemm_mode() != 1 && w->wsplit && tc::tc_supported((const float*)w->Q, C, (const float*)w->wsplit, C, C))
    rc = tc::gemm_tc_nt((const float*)w->Q, C, (const float*)w->wsplit, C, (float*)w->dH, L.fin, p.t_cap, L.fin, C,
                        p.counts + EAGCN_CNT_T, st, "gemm_tc_nt", (const float*)w->wsplit + (size_t)L.fin * C);
  else
    rc = gemm_nt((const float*)w->Q, C, (const float*)w->wall, C, (float*)w->dH, L.fin, p.t_cap, L.fin, C,
                 p.counts + EAGCN_CNT_T, st);
  if (rc) return rc;
  int ns = 0;
  if (gemm_mode() == 0 && tc::tc_supported((const float*)w->H, L.fin, (const float*)w->Q, C, p.t_cap))
    rc = tc::gemm_tc_tn((const float*)w->H, L.fin, (const float*)w->Q, C, (float*)w->gemm_ws,
                        w->gemm_ws_bytes / (long long)sizeof(float), L.fin, C, p.t_cap, p.counts + EAGCN_CNT_T, &ns, st);
  else
    rc = gemm_tn((const float*)w->H, L.fin, (const float*)w->Q, C, nullptr, L.fin, C, p.t_cap, p.counts + EAGCN_CNT_T,
                 (float*)w->gemm_ws, w->gemm_ws_bytes / (long long)sizeof(float), st, &ns);
  if (rc) return rc;
  const int nblk_w = (int)(((long long)L.fin * C + 255) / 256);
  const int nblk_a = (L.V * EAGCN_SIG_STRIDE + 7) / 8;
  EAGCN_PROF("bwd_post_kernel", st);
  bwd_post_kernel<<<nblk_w + nblk_a, 256, 0, st>>>(p, L, (const float*)w->gemm_ws, ns, (float*)w->dwall,
                                                   (const float*)w->partial, (float*)w->datt, nblk_w);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
