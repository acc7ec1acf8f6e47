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
#include <cstdlib>
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
  pdl_prologue();
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
  pdl_prologue();
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

constexpr int kTicketTiles = 192;   // most 32-row tiles for which the in-kernel reduction of the backward sums is used
// grid (stat tiles, ceil(C/4/32)); block (32 channel-groups x 8 row-groups): a thread sums 1/8 of the tile's rows
// for its 4 channels, the 8 row-groups are then combined through shared memory in fixed order
template <int MINB>
__global__ void __launch_bounds__(256, MINB) bn_bwd_partial_vec_kernel(PlanDev p, const float* __restrict__ dX,
                                                                 const float* __restrict__ Y, const float* __restrict__ ball,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 float* __restrict__ partial, int C, int training, float p_drop,
                                                                 const unsigned long long* rng, unsigned long long stream,
                                                                 float* __restrict__ G, int* __restrict__ tickets,
                                                                 double* __restrict__ bsums, float* __restrict__ dvec) {
  pdl_prologue();
  __shared__ __align__(16) float4 s_red[2][8][32];         // 8 KB; the last CTA reuses it as double [8][32][4]
  __shared__ int s_last;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int tile = blockIdx.x;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + cx) * 4;
  const bool act = c < C;
  if (T == 0 && tile == 0 && tickets && ry < 3 && act) {   // empty batch: no tile is live, the sums are zero
#pragma unroll
    for (int u = 0; u < 4; ++u) { dvec[ry * C + c + u] = 0.f; if (ry < 2) bsums[ry * C + c + u] = 0.0; }
  }
  if (tile * kStatRows >= T) return;
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
      // g = dX * relu' * dropout: kept for the tile aggregation kernel, which then needs no Philox replay
      if (G) *reinterpret_cast<float4*>(G + idx) = make_float4(g[0], g[1], g[2], g[3]);
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
  if (!tickets) return;                       // separate stat_reduce launch (global-batch statistics keep it simple too)
  // ---- "last CTA reduces": the CTA that completes a 128-channel block's last tile sums that block's tile partials in
  // FIXED tile order (so the result does not depend on which CTA happens to be last) -> bsums (double) and dbias /
  // dgamma / dbeta: the stat_reduce launch between this kernel and the aggregation backward disappears
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ntile = (T + kStatRows - 1) / kStatRows;
    const int prev = atomicAdd(&tickets[blockIdx.y], 1);
    s_last = prev == ntile - 1;
    if (s_last) tickets[blockIdx.y] = 0;      // leave the ticket array zero for the next call
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const int ntile = (T + kStatRows - 1) / kStatRows;
    double a1[4] = {0.0, 0.0, 0.0, 0.0}, a2[4] = {0.0, 0.0, 0.0, 0.0};
    if (act) {
#pragma unroll 8
      for (int t = ry; t < ntile; t += 8) {
        const float4 x1 = __ldcg(reinterpret_cast<const float4*>(partial + ((size_t)t * 2 + 0) * C + c));
        const float4 x2 = __ldcg(reinterpret_cast<const float4*>(partial + ((size_t)t * 2 + 1) * C + c));
        a1[0] += (double)x1.x; a1[1] += (double)x1.y; a1[2] += (double)x1.z; a1[3] += (double)x1.w;
        a2[0] += (double)x2.x; a2[1] += (double)x2.y; a2[2] += (double)x2.z; a2[3] += (double)x2.w;
      }
    }
    double (*s_fin)[32][4] = reinterpret_cast<double (*)[32][4]>(&s_red[0][0][0]);   // [8][32][4], one sum at a time
#pragma unroll
    for (int k = 0; k < 2; ++k) {             // k = 0: sum g, k = 1: sum g * xhat
      __syncthreads();                        // s_red (k = 0) / the previous round (k = 1) fully consumed
#pragma unroll
      for (int u = 0; u < 4; ++u) s_fin[ry][cx][u] = k == 0 ? a1[u] : a2[u];
      __syncthreads();
      if (ry == k && act) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c + u >= C) break;
          double acc = 0.0;
#pragma unroll
          for (int w = 0; w < 8; ++w) acc += s_fin[w][cx][u];
          bsums[k * C + c + u] = acc;
          if (k == 0) {                       // dvec = [dbias | dgamma | dbeta]
            dvec[c + u] = training ? 0.0f : (float)((double)ball[C + c + u] * (double)invstd[c + u] * acc);
            dvec[2 * C + c + u] = (float)acc;
          } else {
            dvec[C + c + u] = (float)acc;
          }
        }
      }
    }
  }
}

// grid (ceil(C/4/128), ceil(t_cap/kEltRows))
__global__ void __launch_bounds__(128) bn_bwd_apply_vec_kernel(PlanDev p, const float* __restrict__ dX,
                                                               const float* __restrict__ Y, const float* __restrict__ ball,
                                                               const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               const double* __restrict__ bsums, float* __restrict__ dY, int C,
                                                               int training, float p_drop, const unsigned long long* rng,
                                                               unsigned long long stream, double M) {
  pdl_prologue();
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

// grid (row tiles, V); kAggWarps warps x kAggRows rows.  Fo_v handled in super-chunks of 512 channels (4 x 128).
template <int VEC>
__global__ void __launch_bounds__(kAggThreads) agg_bwd_kernel(PlanDev p, LayerDev L, const float* __restrict__ Z,
                                                      const float* __restrict__ Y, const float* __restrict__ dY,
                                                      const float* __restrict__ ball, const float* __restrict__ sig,
                                                      const float* __restrict__ invR, float* __restrict__ Q,
                                                      float* __restrict__ dpart) {
  pdl_prologue();
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

// ---- shared-memory tile variant, BatchNorm backward folded in ---------------------------------------------------
// The warp-per-row kernel above is instruction-bound (Nsight: ~30k warp instructions per 32-row tile and view,
// lanes idle for fo_v = 80 / 140, a 5-step shuffle reduction per edge).  This variant splits the work of a tile into
// phases whose thread mappings fit each job:
//   stage : Z slab (one bulk/TMA copy per row) and dY slab -> shared memory.  dY = a*g - (c1 + xhat*c2) per element from
//           g = dX*relu'*dropout (written by bn_bwd_partial, no Philox replay here), Y and the reduced BatchNorm sums:
//           bn_bwd_apply and the dY round trip through memory disappear;
//   dots  : d_e = dY_t . Z_j for every edge of the tile and d_self = dY_t . Z_t, 8 lanes per dot (4 dots per warp
//           in flight, 3 shuffle steps);  c_t = dY_t . (Y_t - b) follows from them as sum_j A[t,j] d_(t,j);
//   Q     : (row group, float4 channel) items, no cross-lane traffic:  Q_t = a_self dY_t + sum_e A[j->t] dY_j;
//   hist  : d att / d self_r partials: one thread per table entry scans the tile's edges in order (deterministic).
// Rows are processed in chunks whose edge lists fit the staged arrays (one chunk unless the graph is dense);
// neighbours outside the tile are read through L2 (dY recomputed on the fly).
struct BnFold {
  const float* G; const float* mean; const float* invstd; const double* bsums; double M; int training;
};
constexpr int kTileMaxN = 4096;       // padded molecule size up to which the tile kernel's edge staging is sized
inline int tile_edge_cap(int N) { return N > kTileEdgeCap ? ((N + 31) / 32) * 32 : kTileEdgeCap; }
inline size_t agg_bwd_tile_smem(int fo, int cap) {
  return (size_t)(2 * kStatRows * fo + 5 * fo) * sizeof(float) + (size_t)cap * 20 + (size_t)(cap + kStatRows) * 4;
}

__device__ __forceinline__ float4 dy_fold(const float4 g, const float4 y, const float4* sP, int nc4, int c4) {
  const float4 mu = sP[c4], is = sP[nc4 + c4], a = sP[2 * nc4 + c4], c1 = sP[3 * nc4 + c4], c2 = sP[4 * nc4 + c4];
  float4 r;
  r.x = fmaf(a.x, g.x, -fmaf((y.x - mu.x) * is.x, c2.x, c1.x));
  r.y = fmaf(a.y, g.y, -fmaf((y.y - mu.y) * is.y, c2.y, c1.y));
  r.z = fmaf(a.z, g.z, -fmaf((y.z - mu.z) * is.z, c2.z, c1.z));
  r.w = fmaf(a.w, g.w, -fmaf((y.w - mu.w) * is.w, c2.w, c1.w));
  return r;
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}

__global__ void __launch_bounds__(kAggThreads, 4) agg_bwd_tile_kernel(PlanDev p, LayerDev L, BnFold bn,
                                                                   const float* __restrict__ Z, const float* __restrict__ Y,
                                                                   const float* __restrict__ ball,
                                                                   const float* __restrict__ sig,
                                                                   const float* __restrict__ invR, float* __restrict__ Q,
                                                                   float* __restrict__ dpart, int cap) {
  pdl_prologue();
  extern __shared__ __align__(16) float tile_smem_b[];
  __shared__ int s_rp[kStatRows + 1];
  __shared__ float s_invR[kStatRows], s_ct[kStatRows];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v = blockIdx.y, tile = blockIdx.x, t0 = tile * kStatRows;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int fo = L.fo[v], off = L.off[v], ld = L.fo_tot, nc4 = fo >> 2;
  for (int r = warp; r < kStatRows; r += kAggWarps) {   // slack rows [T, t_cap) of Q stay defined (zeros): dW GEMM input
    const int t = t0 + r;
    if (t >= T && t < p.t_cap)
      for (int c = lane; c < nc4; c += 32) reinterpret_cast<float4*>(Q + (size_t)t * ld + off)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (t0 >= T) return;
  const int nrows = min(kStatRows, T - t0);
  float4* sZ = reinterpret_cast<float4*>(tile_smem_b);       // [kStatRows][nc4]
  float4* sG = sZ + kStatRows * nc4;                         // [kStatRows][nc4]   dY
  float4* sP = sG + kStatRows * nc4;                         // [5][nc4]           mean | invstd | a | c1 | c2
  float* s_aq = reinterpret_cast<float*>(sP + 5 * nc4);      // [cap]  A[j -> t] = sigma(rcode) / R_j
  float* s_s = s_aq + cap;                                   // [cap]  sigma(code)
  float* s_ds = s_s + cap;                                   // [cap]  d logit of the edge
  int* s_j = reinterpret_cast<int*>(s_ds + cap);             // [cap]  neighbour row
  int* s_cr = s_j + cap;                                     // [cap]  code | row-in-tile << 8
  float* s_d = reinterpret_cast<float*>(s_cr + cap);         // [cap + kStatRows]  dots: rows first, then edges
  // (row group, float4 channel) mapping of the dY staging and of the Q phase
  const int nrg = __float2int_rz(__fdividef((float)kAggThreads + 0.5f, (float)nc4));
  const int rg = __float2int_rz(__fdividef((float)tid + 0.5f, (float)nc4)), c4 = tid - rg * nc4;
  // The thread's first kPreRows rows of g and Y are requested NOW: their DRAM latency then runs under the staging of
  // the fold parameters / row pointers and the barrier below (Nsight: 30 % of the kernel's samples sat on these loads
  // when they were issued two rows at a time after the barrier).
  constexpr int kPreRows = 4;
  float4 pg[kPreRows], py[kPreRows];
#pragma unroll
  for (int k = 0; k < kPreRows; ++k) {
    const int r = rg + k * nrg;
    if (rg < nrg && r < nrows) {
      const size_t rowoff = (size_t)(t0 + r) * ld + off;
      pg[k] = __ldg(reinterpret_cast<const float4*>(bn.G + rowoff) + c4);
      py[k] = __ldg(reinterpret_cast<const float4*>(Y + rowoff) + c4);
    } else {
      pg[k] = make_float4(0.f, 0.f, 0.f, 0.f); py[k] = pg[k];
    }
  }
  // ---- stage: Z slab, BatchNorm fold parameters, row pointers ----
  slab_load(sZ, Z, t0, nrows, ld, off, fo, &s_bar);
  {
    float* P = reinterpret_cast<float*>(sP);
    for (int c = tid; c < fo; c += kAggThreads) {
      const float is = bn.invstd[off + c], gi = ball[ld + off + c] * is;
      P[c] = bn.mean[off + c]; P[fo + c] = is; P[2 * fo + c] = gi;
      P[3 * fo + c] = bn.training ? gi * (float)(bn.bsums[off + c] / bn.M) : 0.0f;
      P[4 * fo + c] = bn.training ? gi * (float)(bn.bsums[ld + off + c] / bn.M) : 0.0f;
    }
  }
  const float* iR = invR + (size_t)v * p.t_cap;
  if (tid <= nrows) s_rp[tid] = p.row_ptr[t0 + tid];
  if (tid < nrows) s_invR[tid] = iR[t0 + tid];
  __syncthreads();
  // ---- stage: dY slab; a thread keeps one float4 channel's fold parameters in registers and walks rows ----
  if (rg < nrg) {
    const float4 mu = sP[c4], is = sP[nc4 + c4], a = sP[2 * nc4 + c4], c1 = sP[3 * nc4 + c4], c2 = sP[4 * nc4 + c4];
#pragma unroll
    for (int k = 0; k < kPreRows; ++k) {
      const int r = rg + k * nrg;
      if (r < nrows) {
        const float4 g = pg[k], y = py[k];
        float4 o;
        o.x = fmaf(a.x, g.x, -fmaf((y.x - mu.x) * is.x, c2.x, c1.x));
        o.y = fmaf(a.y, g.y, -fmaf((y.y - mu.y) * is.y, c2.y, c1.y));
        o.z = fmaf(a.z, g.z, -fmaf((y.z - mu.z) * is.z, c2.z, c1.z));
        o.w = fmaf(a.w, g.w, -fmaf((y.w - mu.w) * is.w, c2.w, c1.w));
        sG[r * nc4 + c4] = o;
      }
    }
#pragma unroll 2
    for (int r = rg + kPreRows * nrg; r < nrows; r += nrg) {
      const size_t rowoff = (size_t)(t0 + r) * ld + off;
      const float4 g = __ldg(reinterpret_cast<const float4*>(bn.G + rowoff) + c4);
      const float4 y = __ldg(reinterpret_cast<const float4*>(Y + rowoff) + c4);
      float4 o;
      o.x = fmaf(a.x, g.x, -fmaf((y.x - mu.x) * is.x, c2.x, c1.x));
      o.y = fmaf(a.y, g.y, -fmaf((y.y - mu.y) * is.y, c2.y, c1.y));
      o.z = fmaf(a.z, g.z, -fmaf((y.z - mu.z) * is.z, c2.z, c1.z));
      o.w = fmaf(a.w, g.w, -fmaf((y.w - mu.w) * is.w, c2.w, c1.w));
      sG[r * nc4 + c4] = o;
    }
  }
  const float* sg = sig + v * EAGCN_SIG_STRIDE;
  const float sig_r = sg[256];
  const uint8_t* code = p.code + (size_t)v * p.e_cap;
  const uint8_t* rcode = p.rcode + (size_t)v * p.e_cap;
  const int nbins = L.chan[v] + 1;                           // codes 0..C_v (C_v: all-zero relation vector)
  float hist = 0.0f;                                         // thread c < nbins: d att[c];  thread 255: d self_r
  slab_wait(&s_bar);
  int r0 = 0;
  while (r0 < nrows) {
    // rows [r0, r1): the longest run whose edges fit the staged arrays (a single row always fits: deg < N <= cap)
    int r1 = nrows;
    if (s_rp[nrows] - s_rp[r0] > cap) {                      // dense graphs only
      r1 = r0 + 1;
      while (r1 < nrows && s_rp[r1 + 1] - s_rp[r0] <= cap) ++r1;
    }
    const int eA = s_rp[r0], nE = s_rp[r1] - eA, nR = r1 - r0;
    __syncthreads();                                         // previous chunk fully consumed; slabs visible (first pass)
    for (int i = tid; i < nE; i += kAggThreads) {
      const int j = p.col[eA + i];
      s_j[i] = j; s_s[i] = sg[code[eA + i]]; s_aq[i] = sg[rcode[eA + i]] * iR[j];
    }
    if (tid < nR) {                                          // code | owning row, written by the row's thread
      const int r = r0 + tid;
      for (int e = s_rp[r] - eA; e < s_rp[r + 1] - eA; ++e) s_cr[e] = (int)code[eA + e] | (r << 8);
    }
    __syncthreads();
    // ---- dots: index d < nR -> d_self of row r0+d, else edge d-nR; 8 lanes per dot ----
    {
      const int grp = tid >> 3, sub = tid & 7;
      const int ndots = nR + nE;
      for (int base = 0; base < ndots; base += kAggThreads / 8) {
        const int d = base + grp;
        float acc = 0.0f;
        if (d < ndots) {
          int r, j;
          if (d < nR) { r = r0 + d; j = t0 + r; } else { r = s_cr[d - nR] >> 8; j = s_j[d - nR]; }
          const float4* dy = sG + r * nc4;
          const int jr = j - t0;
          if ((unsigned)jr < (unsigned)nrows) {
            const float4* z = sZ + jr * nc4;
            for (int c = sub; c < nc4; c += 8) acc = dot4(dy[c], z[c], acc);
          } else {
            const float4* z = reinterpret_cast<const float4*>(Z + (size_t)j * ld + off);
            for (int c = sub; c < nc4; c += 8) acc = dot4(dy[c], __ldg(z + c), acc);
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (sub == 0 && d < ndots) s_d[d] = acc;
      }
    }
    // ---- Q rows (independent of the dots): (row group, float4 channel) items ----
    if (rg < nrg) {
      for (int r = r0 + rg; r < r1; r += nrg) {
        const float a_self = sig_r * s_invR[r];
        const float4 gt = sG[r * nc4 + c4];
        float q0 = a_self * gt.x, q1 = a_self * gt.y, q2 = a_self * gt.z, q3 = a_self * gt.w;
        for (int e = s_rp[r] - eA; e < s_rp[r + 1] - eA; ++e) {
          const float aq = s_aq[e];
          const int j = s_j[e], jr = j - t0;
          float4 gj;
          if ((unsigned)jr < (unsigned)nrows) gj = sG[jr * nc4 + c4];
          else {
            const size_t ro = (size_t)j * ld + off;
            gj = dy_fold(__ldg(reinterpret_cast<const float4*>(bn.G + ro) + c4),
                         __ldg(reinterpret_cast<const float4*>(Y + ro) + c4), sP, nc4, c4);
          }
          q0 = fmaf(aq, gj.x, q0); q1 = fmaf(aq, gj.y, q1); q2 = fmaf(aq, gj.z, q2); q3 = fmaf(aq, gj.w, q3);
        }
        reinterpret_cast<float4*>(Q + (size_t)(t0 + r) * ld + off)[c4] = make_float4(q0, q1, q2, q3);
      }
    }
    __syncthreads();
    // ---- per row: c_t = dY_t . (Y_t - b) = sum_j A[t,j] d_(t,j);  d logits of the row's edges and of its self loop ----
    if (tid < nR) {
      const int r = r0 + tid;
      const float ir = s_invR[r], dself = s_d[tid];
      const int ea = s_rp[r] - eA, eb = s_rp[r + 1] - eA;
      float ct = sig_r * ir * dself;
      for (int e = ea; e < eb; ++e) ct = fmaf(s_s[e] * ir, s_d[nR + e], ct);
      for (int e = ea; e < eb; ++e) { const float se = s_s[e]; s_ds[e] = (s_d[nR + e] - ct) * ir * se * (1.0f - se); }
      s_ct[r] = (dself - ct) * ir * sig_r * (1.0f - sig_r);
    }
    __syncthreads();
    // ---- table-entry owners scan the chunk's edges in order ----
    if (tid < nbins) {
      for (int e = 0; e < nE; ++e) if ((s_cr[e] & 255) == tid) hist += s_ds[e];
    } else if (tid == kAggThreads - 1) {
      for (int r = r0; r < r1; ++r) hist += s_ct[r];
    }
    r0 = r1;
  }
  float* out = dpart + ((size_t)tile * L.V + v) * EAGCN_SIG_STRIDE;
  out[tid] = tid < nbins ? hist : 0.0f;                      // entries 0..255 (nbins <= 255)
  if (tid == kAggThreads - 1) out[256] = hist;
}

static bool tile_bwd_ok(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w);

// One launch finishing the layer backward:
//  (a) dW: sum the split-K partials [ns][fin][C] in z order and store them VIEW-BLOCKED -- view v's [fin, fo_v]
//      block contiguous at offset fin*off_v -- so that each GraphConv_block's weight gradient is a contiguous tensor;
//  (b) d att / d self_r: sum the per-tile partials in tile order.
__global__ void __launch_bounds__(256) bwd_post_kernel(PlanDev p, LayerDev L, const float* __restrict__ wpart, int ns,
                                                       float* __restrict__ dwall, const float* __restrict__ dpart,
                                                       float* __restrict__ datt, int nblk_w) {
  pdl_prologue();
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

static bool tile_bwd_ok(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w) {
  if (agg_mode() != 0 || !vec4_ok_b(layer) || plan->N > kTileMaxN) return false;
  int fo_max = 0;
  for (int v = 0; v < layer->V; ++v) fo_max = (int)layer->fo[v] > fo_max ? (int)layer->fo[v] : fo_max;
  if (fo_max > kTileMaxFo || agg_bwd_tile_smem(fo_max, tile_edge_cap((int)plan->N)) > 200 * 1024) return false;
  return aligned16(w->dX) && aligned16(w->Y) && aligned16(w->dY) && aligned16(w->Z) && aligned16(w->Q) && aligned16(w->ball);
}

}  // namespace eagcn
using namespace eagcn;

extern "C" int64_t eagcn_gemm_workspace_bytes(int64_t fin, int64_t fo_tot, int64_t t_cap) {
  return gemm_tn_workspace_floats((int)fin, (int)fo_tot, (int)t_cap) * (int64_t)sizeof(float);
}
extern "C" int64_t eagcn_partial_floats(int64_t t_cap, int64_t fo_tot, int64_t V) {
  // + 4 tiles: the fused forward kernel stores its per-row-tile sums as doubles for up to 2 * t_cap / 128 + 1
  // molecule-aligned tiles (4 floats per tile and channel) in this buffer
  const int64_t tiles = eagcn_stat_tiles(t_cap) + 4;
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
  if (!w->dvec) return EAGCN_E_ARG;
  if ((C & 3) == 0 && aligned16(w->dX) && aligned16(w->Y)) {
    dim3 grid((unsigned)eagcn_stat_tiles(p.t_cap), (C / 4 + 31) / 32);
    // "last CTA reduces" pays per CTA (fence + ticket before it retires) and per tile (one CTA sums them all): it beats the
    // separate reduction launch for a few hundred tiles per SM wave (Tox21 batch: +0.3 %) and loses beyond
    // (Lipophilicity B = 512, 432 tiles x 8 channel blocks: -2.6 %, profiles/r02_dp2_notes.md r02w)
    int* tickets = eagcn_stat_tiles(p.t_cap) <= kTicketTiles ? (int*)w->tickets : nullptr;
    EAGCN_PROF("bn_bwd_partial_kernel", st);
    // compiled for 4 resident CTAs per SM (64 registers, 16 B of spills) by default: a third more loads in flight for this
    // bandwidth/latency-shaped kernel, 785.7 k -> 793.3 k molecules/s (r02ad); EAGCN_PARTIAL_MINB=0 selects the 80-register build
    static int minb = -1;
    if (minb < 0) { const char* e = getenv("EAGCN_PARTIAL_MINB"); minb = (e && e[0] == '0') ? 3 : 4; }
    auto kern = minb == 4 ? bn_bwd_partial_vec_kernel<4> : bn_bwd_partial_vec_kernel<3>;
    EAGCN_LAUNCH(kern, grid, 256, 0, st)(p, (const float*)w->dX, (const float*)w->Y, (const float*)w->ball,
                                                    (const float*)w->mean, (const float*)w->invstd, (float*)w->partial,
                                                    C, (w->training & 1) ? 1 : 0, (float)w->p_drop,
                                                    (const unsigned long long*)w->rng,
                                                    (unsigned long long)w->rng_stream,
                                                    tile_bwd_ok(plan, layer, w) ? (float*)w->dY : nullptr,
                                                    tickets, (double*)w->bsums, (float*)w->dvec);
    EAGCN_LAUNCH_CHECK();
    if (tickets) return 0;                              // the sums were reduced by the kernel's last CTAs
  } else {
    dim3 grid((unsigned)eagcn_stat_tiles(p.t_cap), (C + 127) / 128);
    EAGCN_PROF("bn_bwd_partial_kernel", st);
    EAGCN_LAUNCH(bn_bwd_partial_kernel, grid, 256, 0, st)(p, (const float*)w->dX, (const float*)w->Y, (const float*)w->ball,
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
  EAGCN_LAUNCH(stat_reduce_kernel, (C + 31) / 32, 32 * kStatLanes, 0, st)(p, L, (const float*)w->partial, (double*)w->bsums, C, ep, 0);
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
  const int phase = w->phase ? (int)w->phase : 7;     // bit0: aggregation backward, bit1: dH, bit2: dW + parameter sums
  if (phase & 1) {
  if (tile_bwd_ok(plan, layer, w)) {
    // fused: w->dY holds g = dX*relu'*dropout (written by backward_a); dY itself only ever exists in shared memory
    BnFold bn{(const float*)w->dY, (const float*)w->mean, (const float*)w->invstd, (const double*)w->bsums, M,
              (w->training & 1) ? 1 : 0};
    const int cap = tile_edge_cap(p.N);
    const size_t smem = agg_bwd_tile_smem(fo_max, cap);
    static size_t smem_set = 0;
    if (smem > smem_set) {
      cudaError_t e = cudaFuncSetAttribute(agg_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      smem_set = smem;
    }
    EAGCN_PROF("agg_bwd_kernel", st);
    EAGCN_LAUNCH(agg_bwd_tile_kernel, grid, kAggThreads, smem, st)(p, L, bn, (const float*)w->Z, (const float*)w->Y,
                                                         (const float*)w->ball, (const float*)w->sig,
                                                         (const float*)w->invR, (float*)w->Q, (float*)w->partial, cap);
  } else {
  if ((C & 3) == 0 && aligned16(w->dX) && aligned16(w->Y) && aligned16(w->dY)) {
    dim3 grid((C / 4 + 127) / 128, (p.t_cap + kEltRows - 1) / kEltRows);
    EAGCN_PROF("bn_bwd_apply_kernel", st);
    EAGCN_LAUNCH(bn_bwd_apply_vec_kernel, grid, 128, 0, st)(
        p, (const float*)w->dX, (const float*)w->Y, (const float*)w->ball, (const float*)w->mean,
        (const float*)w->invstd, (const double*)w->bsums, (float*)w->dY, C, (w->training & 1) ? 1 : 0, (float)w->p_drop,
        (const unsigned long long*)w->rng, (unsigned long long)w->rng_stream, M);
    EAGCN_LAUNCH_CHECK();
  } else {
    EAGCN_PROF("bn_bwd_apply_kernel", st);
    EAGCN_LAUNCH(bn_bwd_apply_kernel, (unsigned)((total + 255) / 256), 256, 0, st)(
        p, (const float*)w->dX, (const float*)w->Y, (const float*)w->ball, (const float*)w->mean, (const float*)w->invstd,
        (const double*)w->bsums, (float*)w->dY, C, (w->training & 1) ? 1 : 0, (float)w->p_drop,
        (const unsigned long long*)w->rng, (unsigned long long)w->rng_stream, M);
    EAGCN_LAUNCH_CHECK();
  }
  if (vec4_ok_b(layer)) {
    EAGCN_PROF("agg_bwd_kernel", st);
    EAGCN_LAUNCH((agg_bwd_kernel<4>), grid, kAggThreads, 0, st)(p, L, (const float*)w->Z, (const float*)w->Y, (const float*)w->dY,
                                            (const float*)w->ball, (const float*)w->sig, (const float*)w->invR,
                                            (float*)w->Q, (float*)w->partial);
  } else {
    EAGCN_PROF("agg_bwd_kernel", st);
    EAGCN_LAUNCH((agg_bwd_kernel<1>), grid, kAggThreads, 0, st)(p, L, (const float*)w->Z, (const float*)w->Y, (const float*)w->dY,
                                            (const float*)w->ball, (const float*)w->sig, (const float*)w->invR,
                                            (float*)w->Q, (float*)w->partial);
  }
  }
  EAGCN_LAUNCH_CHECK();
  }
  int rc = 0;
  if (!(phase & 2) || !w->dH) {
    // the layer input needs no gradient (first layer fed by data): skip dH = Q W^T altogether
  } else if (gemm_mode() != 1 && w->wsplit && tc::tc_supported((const float*)w->Q, C, (const float*)w->wsplit, C, C))
    rc = tc::gemm_tc_nt((const float*)w->Q, C, (const float*)w->wsplit, C, (float*)w->dH, L.fin, p.t_cap, L.fin, C,
                        p.counts + EAGCN_CNT_T, st, "gemm_tc_nt", (const float*)w->wsplit + (size_t)L.fin * C);
  else
    rc = gemm_nt((const float*)w->Q, C, (const float*)w->wall, C, (float*)w->dH, L.fin, p.t_cap, L.fin, C,
                 p.counts + EAGCN_CNT_T, st);
  if (rc) return rc;
  if (!(phase & 4)) return 0;
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
  EAGCN_LAUNCH(bwd_post_kernel, nblk_w + nblk_a, 256, 0, st)(p, L, (const float*)w->gemm_ws, ns, (float*)w->dwall,
                                                   (const float*)w->partial, (float*)w->datt, nblk_w);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
