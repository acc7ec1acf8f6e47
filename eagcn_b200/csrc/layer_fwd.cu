// Forward of one GraphConv_Layer (reference layers.py:293-325, structure 'Concate') on packed rows.
//
//   prep   : W_all = [W_1 | ... | W_V] (fin x fo_tot), bias/gamma/beta vectors, sigmoid(att) tables
//   gemm   : Z = H . W_all                       (layers.py:40; reassociated: A_v (H W_v) == (A_v H) W_v)
//   agg    : per row i, view v -- one warp per row:
//              w_e  = sigmoid(a_v[code_v(e)])                         layers.py:82-83 (1x1 conv == lookup)
//              R    = sum_e w_e + sigmoid(r_v) + (N - deg_i) * 1e-9   layers.py:84,87  (self loop + tiny)
//              Y_v  = sum_e (w_e/R) Z_v[j_e] + (sigmoid(r_v)/R) Z_v[i] + b_v      layers.py:90,39,43
//            + per-channel sum / sum-of-squares of (Y - b) for BatchNorm (warp-shuffle row reductions,
//              fixed-order cross-warp / cross-tile reduction -> deterministic)
//   bn     : statistics over ALL B*N padded positions (layers.py:408-412): the (B*N - T) rows that are
//            padding / bond-less hold exactly Y = b, so they add 0 to both centred sums and only enter
//            through the population size M.  Running stats updated like nn.BatchNorm1d.
//   apply  : X = dropout(relu(BN(Y)))            (layers.py:93-94), rows ordered as the concat of
//            layers.py:313; padded rows are never materialised here (rows_scatter writes their zeros).
//
// The 1e-9 "mask_tiny" weights stay in the normaliser R; their contribution to the aggregate
// (<= N*1e-9*|H|, measured <= 1.2e-7 absolute) is below fp32 resolution of the
// reference's own summation and is not materialised.
#include "common.cuh"

namespace eagcn {

int gemm_nn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int Mcap, int N, int K,
            const int* Mdev, cudaStream_t st);
inline int& gemm_mode() { static int m = 0; return m; }
inline int& agg_mode() { static int m = 0; return m; }   // 0: shared-memory tile kernels when eligible, 1: generic
inline int& fuse_mode() { static int m = 0; return m; }  // 0: statistics reduction fused into the BatchNorm apply kernel, 1: separate
namespace fz {
bool fwd_fused_ok(const eagcn_plan_t* plan, const eagcn_layer_t* l, const eagcn_work_t* w, int gemm_mode_, int agg_mode_);
int layer_fwd_fused(const eagcn_plan_t* plan, const PlanDev& p, const LayerDev& L, const eagcn_work_t* w, cudaStream_t st);
}

// ---------------------------------------------------------------------------------------------
// wallT: [2][fo_tot][fin] and wsplit: [2][fin][fo_tot] hold the exact TF32 split of every weight
// (hi = 13 low mantissa bits cleared, lo = w - hi) so the tensor-core GEMMs need not split B in shared memory
__global__ void __launch_bounds__(256) prep_params_kernel(LayerDev L, float* __restrict__ wall, float* __restrict__ wallT,
                                                          float* __restrict__ wsplit, float* __restrict__ ball,
                                                          float* __restrict__ sig) {
  pdl_prologue();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long nW = (long long)L.fin * L.fo_tot;
  if (idx < nW) {
    const int k = (int)(idx / L.fo_tot), c = (int)(idx - (long long)k * L.fo_tot);
    int v = 0;
    while (v + 1 < L.V && c >= L.off[v + 1]) ++v;
    const float wv = __ldg(L.W[v] + (long long)k * L.fo[v] + (c - L.off[v]));
    wall[idx] = wv;
    const float hi = __uint_as_float(__float_as_uint(wv) & 0xFFFFE000u), lo = wv - hi;
    if (wallT) { wallT[(long long)c * L.fin + k] = hi; wallT[nW + (long long)c * L.fin + k] = lo; }
    if (wsplit) { wsplit[idx] = hi; wsplit[nW + idx] = lo; }
  }
  if (idx < L.fo_tot) {
    const int c = (int)idx;
    int v = 0;
    while (v + 1 < L.V && c >= L.off[v + 1]) ++v;
    const int cc = c - L.off[v];
    ball[c] = L.bias[v][cc];
    ball[L.fo_tot + c] = L.gamma[v][cc];
    ball[2 * L.fo_tot + c] = L.beta[v][cc];
    ball[3 * L.fo_tot + c] = 0.0f;
  }
  if (idx < (long long)L.V * EAGCN_SIG_STRIDE) {
    const int v = (int)(idx / EAGCN_SIG_STRIDE), c = (int)(idx - (long long)v * EAGCN_SIG_STRIDE);
    float s = 0.5f;                                   // code == C_v (all-zero relation vector): sigmoid(0)
    if (c < L.chan[v]) s = sigmoidf_(L.att_w[v][c]);
    if (c == 256) s = sigmoidf_(L.self_r[v][0]);
    sig[idx] = s;
  }
}

// lane <-> channel mapping of one 128-channel chunk q: VEC=4 -> lane*4+u (float4), VEC=1 -> u*32+lane
template <int VEC>
__device__ __forceinline__ int chan_of(int q, int lane, int u) {
  return VEC == 4 ? q * 128 + lane * 4 + u : q * 128 + u * 32 + lane;
}
template <int VEC>
__device__ __forceinline__ void load4(const float* __restrict__ row, int q, int lane, int lim, float (&o)[4]) {
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
__device__ __forceinline__ void store4(float* __restrict__ row, int q, int lane, int lim, const float (&o)[4]) {
  if (VEC == 4) {
    const int c = q * 128 + lane * 4;
    if (c < lim) *reinterpret_cast<float4*>(row + c) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int c = q * 128 + u * 32 + lane; if (c < lim) row[c] = o[u]; }
  }
}

// Single-pass variant for fo_v <= 128*NQ (float4 layout): every lane keeps NQ float4 accumulators, so the row's
// edge list is walked once (one sigmoid-table lookup and one neighbour-index broadcast per edge) instead of once
// per 128-channel chunk.  grid (row tiles of kStatRows, V); kAggWarps warps x kAggRows rows.
template <int NQ>
__global__ void __launch_bounds__(kAggThreads) agg_fwd_onepass_kernel(PlanDev p, LayerDev L, const float* __restrict__ Z,
                                                                      const float* __restrict__ ball,
                                                                      const float* __restrict__ sig, float* __restrict__ Y,
                                                                      float* __restrict__ invR, float* __restrict__ partial,
                                                                      int n_pad, int want_stats) {
  pdl_prologue();
  __shared__ float s_red[kAggWarps][2][NQ * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.y;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int tile = blockIdx.x;
  if (tile * kStatRows >= T) return;
  const int fo = L.fo[v], off = L.off[v], ld = L.fo_tot;
  const float* sg = sig + v * EAGCN_SIG_STRIDE;
  const float sig_r = sg[256];
  const uint8_t* code = p.code + (size_t)v * p.e_cap;
  float s1[NQ][4], s2[NQ][4], bias4[NQ][4];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    load4<4>(ball + off, q, lane, fo, bias4[q]);
#pragma unroll
    for (int u = 0; u < 4; ++u) { s1[q][u] = 0.f; s2[q][u] = 0.f; }
  }
  for (int r = 0; r < kAggRows; ++r) {
    const int t = tile * kStatRows + warp * kAggRows + r;
    if (t >= T) break;
    const int e0 = p.row_ptr[t], e1 = p.row_ptr[t + 1];
    const int deg = e1 - e0;
    float sw = 0.0f;                                               // attention row sum (layers.py:84,87)
    for (int e = e0 + lane; e < e1; e += 32) sw += sg[code[e]];
    sw = warp_sum(sw);
    const float R = sw + sig_r + (float)(n_pad - deg) * EAGCN_TINY;
    if (lane == 0) invR[(size_t)v * p.t_cap + t] = 1.0f / R;
    const float a_self = sig_r / R;
    float acc[NQ][4];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float zi[4];
      load4<4>(Z + (size_t)t * ld + off, q, lane, fo, zi);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[q][u] = a_self * zi[u];
    }
    for (int eb = e0; eb < e1; eb += 32) {                         // aggregate (layers.py:90,39)
      const int e = eb + lane;
      float a_e = 0.0f; int j_e = 0;
      if (e < e1) { a_e = sg[code[e]] / R; j_e = p.col[e]; }
      const int cnt = min(32, e1 - eb);
#pragma unroll 2
      for (int k = 0; k < cnt; ++k) {
        const float a = __shfl_sync(0xffffffffu, a_e, k);
        const int j = __shfl_sync(0xffffffffu, j_e, k);
        const float* zr = Z + (size_t)j * ld + off;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          float zj[4];
          load4<4>(zr, q, lane, fo, zj);
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[q][u] = fmaf(a, zj[u], acc[q][u]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        y[u] = acc[q][u] + bias4[q][u];
        s1[q][u] += acc[q][u]; s2[q][u] = fmaf(acc[q][u], acc[q][u], s2[q][u]);
      }
      store4<4>(Y + (size_t)t * ld + off, q, lane, fo, y);
    }
  }
  if (want_stats) {                                                // cross-warp reduction in fixed order
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s_red[warp][0][q * 128 + lane * 4 + u] = s1[q][u];
        s_red[warp][1][q * 128 + lane * 4 + u] = s2[q][u];
      }
    __syncthreads();
    for (int c = threadIdx.x; c < fo; c += kAggThreads) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w = 0; w < kAggWarps; ++w) { a += s_red[w][0][c]; b += s_red[w][1][c]; }
      partial[((size_t)tile * 2 + 0) * ld + off + c] = a;
      partial[((size_t)tile * 2 + 1) * ld + off + c] = b;
    }
  }
}

// ---- shared-memory tile variant -------------------------------------------------------------------------
// one bulk (TMA) copy per slab row instead of fo/4 16-byte cp.async's: global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
// stage rows [t0, t0+nrows) x [off, off+fo) of a [*, ld] matrix as a dense [nrows][fo] slab; call from all threads
// BEFORE any divergence, then slab_wait() before reading.  Needs (fo*4) % 16 == 0 and 16-byte aligned rows.
__device__ __forceinline__ void slab_load(float4* slab, const float* __restrict__ src, int t0, int nrows, int ld, int off,
                                          int fo, uint64_t* bar) {
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) tc::mbar_expect_tx(bar, (uint32_t)nrows * (uint32_t)fo * 4u);
    __syncwarp();
    if ((int)threadIdx.x < nrows)
      bulk_g2s(slab + (size_t)threadIdx.x * (fo >> 2), src + (size_t)(t0 + threadIdx.x) * ld + off, (uint32_t)fo * 4u, bar);
  }
}
__device__ __forceinline__ void slab_wait(uint64_t* bar) { tc::mbar_wait(bar, 0); }

constexpr int kTileEdgeCap = 512;     // edges of one row tile staged in shared memory (beyond: read through L2)
constexpr int kTileMaxFo = 512;       // widest view the tile kernels take (float4 layout)

inline int tile_row_groups(int fo) { return kAggThreads / (fo / 4); }
inline size_t agg_fwd_tile_smem(int fo) { return (size_t)(kStatRows * fo + tile_row_groups(fo) * 2 * fo) * sizeof(float); }

// Tile variant (float4 layout, fo_v <= kTileMaxFo).  Atoms of a molecule are consecutive packed rows, so nearly
// every neighbour of a row lives in the same 32-row tile: the CTA stages its kStatRows x fo_v slab of Z with one bulk
// (TMA) copy per row (instead of one dependent L2 round trip per row and edge) together with the tile's
// edge list, then (row group, float4 channel) items -- no idle lanes for fo_v = 80 / 140 -- walk the edges out of
// shared memory.  Neighbours outside the tile (molecules straddling a tile boundary) are read through L2.
// The attention row sums are accumulated edge by edge (the generic kernel: lane-strided + shuffle tree), so Y can
// differ from the generic kernel's in the last bit; the statistics partials are summed per row group.
__global__ void __launch_bounds__(kAggThreads, 5) agg_fwd_tile_kernel(PlanDev p, LayerDev L, const float* __restrict__ Z,
                                                                   const float* __restrict__ ball,
                                                                   const float* __restrict__ sig, float* __restrict__ Y,
                                                                   float* __restrict__ invR, float* __restrict__ partial,
                                                                   int n_pad, int want_stats) {
  pdl_prologue();
  extern __shared__ __align__(16) float tile_smem[];
  __shared__ int s_rp[kStatRows + 1];
  __shared__ float2 s_e[kTileEdgeCap];     // per staged edge: (a_e = sigma(code) / R_row, neighbour row as int bits)
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ float s_R[kStatRows];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v = blockIdx.y, tile = blockIdx.x, t0 = tile * kStatRows;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  if (t0 >= T) return;
  const int nrows = min(kStatRows, T - t0);
  const int fo = L.fo[v], off = L.off[v], ld = L.fo_tot, nc4 = fo >> 2;
  float4* sZ = reinterpret_cast<float4*>(tile_smem);             // [kStatRows][nc4]
  float* s_red = tile_smem + kStatRows * fo;                     // [row groups][2][fo]
  slab_load(sZ, Z, t0, nrows, ld, off, fo, &s_bar);
  if (tid <= nrows) s_rp[tid] = p.row_ptr[t0 + tid];
  __syncthreads();
  const float* sg = sig + v * EAGCN_SIG_STRIDE;
  const float sig_r = sg[256];
  const uint8_t* code = p.code + (size_t)v * p.e_cap;
  const int e_lo = s_rp[0], ne = s_rp[nrows] - e_lo;
  for (int i = tid; i < min(ne, kTileEdgeCap); i += kAggThreads)
    s_e[i] = make_float2(sg[code[e_lo + i]], __int_as_float(p.col[e_lo + i]));
  __syncthreads();
  // attention row sums (layers.py:84,87): one thread per row walks its (few) edges in order
  if (tid < nrows) {
    const int a0 = s_rp[tid] - e_lo, a1 = s_rp[tid + 1] - e_lo;
    float sw = 0.0f;
    for (int e = a0; e < a1; ++e) sw += e < kTileEdgeCap ? s_e[e].x : sg[code[e_lo + e]];
    const float R = sw + sig_r + (float)(n_pad - (a1 - a0)) * EAGCN_TINY;
    invR[(size_t)v * p.t_cap + t0 + tid] = 1.0f / R;
    s_R[tid] = R;
    for (int e = a0; e < min(a1, kTileEdgeCap); ++e) s_e[e].x = s_e[e].x / R;
  }
  __syncthreads();
  slab_wait(&s_bar);
  const int nrg = __float2int_rz(__fdividef((float)kAggThreads + 0.5f, (float)nc4));
  const int rg = __float2int_rz(__fdividef((float)tid + 0.5f, (float)nc4)), c4 = tid - rg * nc4;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (rg < nrg) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(ball + off) + c4);
    for (int r = rg; r < nrows; r += nrg) {                      // aggregate (layers.py:90,39)
      const float R = s_R[r];
      const float a_self = sig_r / R;
      const float4 zt = sZ[r * nc4 + c4];
      float acc[4] = {a_self * zt.x, a_self * zt.y, a_self * zt.z, a_self * zt.w};
      const int a0 = s_rp[r] - e_lo, a1 = s_rp[r + 1] - e_lo;
      for (int e = a0; e < a1; ++e) {
        float a; int j;
        if (e < kTileEdgeCap) { const float2 ed = s_e[e]; a = ed.x; j = __float_as_int(ed.y); }
        else { a = sg[code[e_lo + e]] / R; j = p.col[e_lo + e]; }
        const int jr = j - t0;
        const float4 zj = (unsigned)jr < (unsigned)nrows
                              ? sZ[jr * nc4 + c4]
                              : __ldg(reinterpret_cast<const float4*>(Z + (size_t)j * ld + off) + c4);
        acc[0] = fmaf(a, zj.x, acc[0]); acc[1] = fmaf(a, zj.y, acc[1]);
        acc[2] = fmaf(a, zj.z, acc[2]); acc[3] = fmaf(a, zj.w, acc[3]);
      }
      reinterpret_cast<float4*>(Y + (size_t)(t0 + r) * ld + off)[c4] =
          make_float4(acc[0] + b4.x, acc[1] + b4.y, acc[2] + b4.z, acc[3] + b4.w);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s1[u] += acc[u]; s2[u] = fmaf(acc[u], acc[u], s2[u]); }
    }
  }
  if (want_stats) {                                              // cross-row-group reduction in fixed order
    if (rg < nrg) {
      *reinterpret_cast<float4*>(s_red + (rg * 2 + 0) * fo + c4 * 4) = make_float4(s1[0], s1[1], s1[2], s1[3]);
      *reinterpret_cast<float4*>(s_red + (rg * 2 + 1) * fo + c4 * 4) = make_float4(s2[0], s2[1], s2[2], s2[3]);
    }
    __syncthreads();
    for (int c = tid; c < fo; c += kAggThreads) {
      float a = 0.f, b = 0.f;
      for (int g = 0; g < nrg; ++g) { a += s_red[(g * 2 + 0) * fo + c]; b += s_red[(g * 2 + 1) * fo + c]; }
      partial[((size_t)tile * 2 + 0) * ld + off + c] = a;
      partial[((size_t)tile * 2 + 1) * ld + off + c] = b;
    }
  }
}

// grid (row tiles of kStatRows, V).  kAggWarps warps x kAggRows rows.
template <int VEC>
__global__ void __launch_bounds__(kAggThreads) agg_fwd_kernel(PlanDev p, LayerDev L, const float* __restrict__ Z,
                                                      const float* __restrict__ ball, const float* __restrict__ sig,
                                                      float* __restrict__ Y, float* __restrict__ invR,
                                                      float* __restrict__ partial, int n_pad, int want_stats) {
  pdl_prologue();
  __shared__ float s_red[kAggWarps][2][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.y;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int tile = blockIdx.x;
  if (tile * kStatRows >= T) return;
  const int fo = L.fo[v], off = L.off[v], ld = L.fo_tot;
  const float* sg = sig + v * EAGCN_SIG_STRIDE;
  const float sig_r = sg[256];
  const uint8_t* code = p.code + (size_t)v * p.e_cap;
  const int nq = (fo + 127) / 128;
  for (int q = 0; q < nq; ++q) {
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    float bias4[4];
    load4<VEC>(ball + off, q, lane, fo, bias4);
    for (int r = 0; r < kAggRows; ++r) {
      const int t = tile * kStatRows + warp * kAggRows + r;
      if (t >= T) break;
      const int e0 = p.row_ptr[t], e1 = p.row_ptr[t + 1];
      const int deg = e1 - e0;
      // ---- attention row sum (layers.py:84,87) ----
      float sw = 0.0f;
      for (int e = e0 + lane; e < e1; e += 32) sw += sg[code[e]];
      sw = warp_sum(sw);
      const float R = sw + sig_r + (float)(n_pad - deg) * EAGCN_TINY;
      if (q == 0 && lane == 0) invR[(size_t)v * p.t_cap + t] = 1.0f / R;
      // ---- aggregate (layers.py:90,39) ----
      float acc[4], zi[4];
      load4<VEC>(Z + (size_t)t * ld + off, q, lane, fo, zi);
      const float a_self = sig_r / R;
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = a_self * zi[u];
      for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        float a_e = 0.0f; int j_e = 0;
        if (e < e1) { a_e = sg[code[e]] / R; j_e = p.col[e]; }
        const int cnt = min(32, e1 - eb);
        for (int k = 0; k < cnt; ++k) {
          const float a = __shfl_sync(0xffffffffu, a_e, k);
          const int j = __shfl_sync(0xffffffffu, j_e, k);
          float zj[4];
          load4<VEC>(Z + (size_t)j * ld + off, q, lane, fo, zj);
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fmaf(a, zj[u], acc[u]);
        }
      }
      float y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { y[u] = acc[u] + bias4[u]; s1[u] += acc[u]; s2[u] = fmaf(acc[u], acc[u], s2[u]); }
      store4<VEC>(Y + (size_t)t * ld + off, q, lane, fo, y);
    }
    if (want_stats) {
      // cross-warp reduction in fixed order
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cl = VEC == 4 ? lane * 4 + u : u * 32 + lane;
        s_red[warp][0][cl] = s1[u]; s_red[warp][1][cl] = s2[u];
      }
      __syncthreads();
      if (threadIdx.x < 128) {
        const int c = q * 128 + threadIdx.x;
        if (c < fo) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w = 0; w < kAggWarps; ++w) { a += s_red[w][0][threadIdx.x]; b += s_red[w][1][threadIdx.x]; }
          partial[((size_t)tile * 2 + 0) * ld + off + c] = a;
          partial[((size_t)tile * 2 + 1) * ld + off + c] = b;
        }
      }
      __syncthreads();
    }
  }
}

// per-channel BatchNorm finalize from the batch sums (training) or the running statistics (eval)
__device__ __forceinline__ void bn_finalize_channel(const LayerDev& L, const float* __restrict__ ball, int c, double s1,
                                                    double s2, float* __restrict__ mean, float* __restrict__ invstd,
                                                    int training, double M, double eps, double momentum) {
  int v = 0;
  while (v + 1 < L.V && c >= L.off[v + 1]) ++v;
  const int cc = c - L.off[v];
  if (training) {
    const double b = (double)ball[c];
    const double m1 = s1 / M;                            // mean of (Y - b) over all positions
    double var = s2 / M - m1 * m1;                       // biased variance (shift-invariant)
    if (var < 0.0) var = 0.0;
    const double mu = b + m1;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + eps));
    const double unb = M > 1.0 ? var * (M / (M - 1.0)) : var;
    L.run_mean[v][cc] = (float)((1.0 - momentum) * (double)L.run_mean[v][cc] + momentum * mu);
    L.run_var[v][cc] = (float)((1.0 - momentum) * (double)L.run_var[v][cc] + momentum * unb);
    if (cc == 0 && L.nbt[v]) L.nbt[v][0] += 1;
  } else {
    mean[c] = L.run_mean[v][cc];
    invstd[c] = 1.0f / sqrtf(L.run_var[v][cc] + (float)eps);
  }
}

struct StatEpilogue {             // what stat_reduce_kernel does with the reduced sums of a channel
  int kind;                       // 0: store sums only, 1: + forward BatchNorm finalize, 2: + backward dvec
  const float* ball; float* mean; float* invstd; float* dvec;
  int training; double M, eps, momentum;
};

// sums[k][c] = sum over live tiles of partial[tile][k][c]   (double, fixed order), then the per-channel epilogue.
// Block = 32 channels x 32 tile lanes: a thread adds every 32nd tile (independent loads, ~5 per thread at Tox21 sizes,
// instead of a 19-deep dependent chain), the 32 lanes are then combined through shared memory in lane order.
constexpr int kStatLanes = 32;
__global__ void __launch_bounds__(32 * kStatLanes) stat_reduce_kernel(PlanDev p, LayerDev L,
                                                                      const float* __restrict__ partial,
                                                                      double* __restrict__ sums, int C, StatEpilogue ep,
                                                                      int aligned_tiles) {
  pdl_prologue();
  __shared__ double s[kStatLanes][2][32];
  const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  // partials per 32-row tile, or per molecule-aligned row tile when the fused forward kernel produced them
  const int ntile = aligned_tiles ? p.counts[EAGCN_CNT_TILES] : (T + kStatRows - 1) / kStatRows;
  double a = 0.0, b = 0.0;
  if (c < C) {
    if (aligned_tiles) {             // the fused forward kernel writes its (fewer, larger) tile partials in double
      const double* pd = reinterpret_cast<const double*>(partial);
#pragma unroll 4
      for (int t = ty; t < ntile; t += kStatLanes) {
        a += __ldg(pd + ((size_t)t * 2 + 0) * C + c);
        b += __ldg(pd + ((size_t)t * 2 + 1) * C + c);
      }
    } else {
#pragma unroll 4
      for (int t = ty; t < ntile; t += kStatLanes) {
        a += (double)__ldg(partial + ((size_t)t * 2 + 0) * C + c);
        b += (double)__ldg(partial + ((size_t)t * 2 + 1) * C + c);
      }
    }
  }
  s[ty][0][cx] = a; s[ty][1][cx] = b;
  __syncthreads();
  if (ty < 2 && c < C) {                 // warp 0 finishes the S1 column, warp 1 the S2 column
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < kStatLanes; ++w) acc += s[w][ty][cx];
    sums[ty * C + c] = acc;
    s[0][ty][cx] = acc;
  }
  __syncthreads();
  if (ty == 0 && c < C) {
    const double aa = s[0][0][cx], bb = s[0][1][cx];
    if (ep.kind == 1) {
      bn_finalize_channel(L, ep.ball, c, aa, bb, ep.mean, ep.invstd, ep.training, ep.M, ep.eps, ep.momentum);
    } else if (ep.kind == 2) {      // dvec = [dbias | dgamma | dbeta]
      ep.dvec[c] = ep.training ? 0.0f : (float)((double)ep.ball[C + c] * (double)ep.invstd[c] * aa);
      ep.dvec[C + c] = (float)bb;
      ep.dvec[2 * C + c] = (float)aa;
    }
  }
}

// Training-mode forward with per-replica statistics: stat_reduce (kind 1) and bn_apply in ONE launch.  grid (ceil(C/32), R):
// every CTA of a channel block repeats the (cheap, L2-resident) fixed-order reduction of the tile partials for its 32
// channels -- the same order as stat_reduce_kernel, so the statistics are bit-identical -- and then normalises its share
// of the rows; the CTAs with blockIdx.y == 0 also publish sums / mean / invstd and update the running statistics.
// Saves one dependent launch per layer (the step is bound by launch count and latency).
__global__ void __launch_bounds__(32 * kStatLanes) bn_stat_apply_kernel(
    PlanDev p, LayerDev L, const float* __restrict__ partial, double* __restrict__ sums, const float* __restrict__ Y,
    const float* __restrict__ ball, float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ X, int C,
    double M, double eps, double momentum, float p_drop, const unsigned long long* rng, unsigned long long rng_stream,
    int aligned_tiles) {
  pdl_prologue();
  __shared__ double s[kStatLanes][2][32];
  __shared__ __align__(16) float s_mu[32], s_is[32];
  const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int ntile = aligned_tiles ? p.counts[EAGCN_CNT_TILES] : (T + kStatRows - 1) / kStatRows;
  // apply-phase mapping: 8 float4 channel lanes x 128 row lanes over this CTA's share of the rows.  The thread's first
  // kPre rows of Y are requested NOW, so that their memory latency runs under the reduction below.
  constexpr int kPre = 6;
  const int c4l = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int cc = blockIdx.x * 32 + c4l * 4;
  const int per = (p.t_cap + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(p.t_cap, r0 + per);
  float4 yv[kPre];
#pragma unroll
  for (int k = 0; k < kPre; ++k) {
    const int t = r0 + rl + 128 * k;
    yv[k] = (cc < C && t < r1 && t < T) ? __ldg(reinterpret_cast<const float4*>(Y + (size_t)t * C + cc))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  double a = 0.0, b = 0.0;
  if (c < C) {
    if (aligned_tiles) {             // the fused forward kernel writes its (fewer, larger) tile partials in double
      const double* pd = reinterpret_cast<const double*>(partial);
#pragma unroll 4
      for (int t = ty; t < ntile; t += kStatLanes) {
        a += __ldg(pd + ((size_t)t * 2 + 0) * C + c);
        b += __ldg(pd + ((size_t)t * 2 + 1) * C + c);
      }
    } else {
#pragma unroll 4
      for (int t = ty; t < ntile; t += kStatLanes) {
        a += (double)__ldg(partial + ((size_t)t * 2 + 0) * C + c);
        b += (double)__ldg(partial + ((size_t)t * 2 + 1) * C + c);
      }
    }
  }
  s[ty][0][cx] = a; s[ty][1][cx] = b;
  __syncthreads();
  if (ty < 2 && c < C) {                 // warp 0 finishes the S1 column, warp 1 the S2 column
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < kStatLanes; ++w) acc += s[w][ty][cx];
    if (blockIdx.y == 0) sums[ty * C + c] = acc;
    s[0][ty][cx] = acc;
  }
  __syncthreads();
  if (ty == 0) {
    float mu = 0.f, is = 0.f;
    if (c < C) {
      const double s1 = s[0][0][cx], s2 = s[0][1][cx];
      const double bb = (double)ball[c];
      const double m1 = s1 / M;                            // mean of (Y - b) over all positions
      double var = s2 / M - m1 * m1;                       // biased variance (shift-invariant)
      if (var < 0.0) var = 0.0;
      const double mud = bb + m1;
      mu = (float)mud;
      is = (float)(1.0 / sqrt(var + eps));
      if (blockIdx.y == 0) {                               // publish + running statistics (bn_finalize_channel)
        int v = 0;
        while (v + 1 < L.V && c >= L.off[v + 1]) ++v;
        const int cc = c - L.off[v];
        mean[c] = mu; invstd[c] = is;
        const double unb = M > 1.0 ? var * (M / (M - 1.0)) : var;
        L.run_mean[v][cc] = (float)((1.0 - momentum) * (double)L.run_mean[v][cc] + momentum * mud);
        L.run_var[v][cc] = (float)((1.0 - momentum) * (double)L.run_var[v][cc] + momentum * unb);
        if (cc == 0 && L.nbt[v]) L.nbt[v][0] += 1;
      }
    }
    s_mu[cx] = mu; s_is[cx] = is;
  }
  __syncthreads();
  // ---- apply ----
  if (cc >= C) return;
  const float4 mu = *reinterpret_cast<const float4*>(s_mu + c4l * 4), is = *reinterpret_cast<const float4*>(s_is + c4l * 4);
  const float4 ga = *reinterpret_cast<const float4*>(ball + C + cc), be = *reinterpret_cast<const float4*>(ball + 2 * C + cc);
  const bool drop = p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  unsigned long long seed = 0, off = 0;
  if (drop) { seed = rng[0]; off = rng[1]; }
  const Philox ph(seed);
  for (int k = 0, t = r0 + rl; t < r1; t += 128, ++k) {
    const size_t idx = (size_t)t * C + cc;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T) {
      float4 y;
      switch (k) {                                          // registers for the prefetched rows, memory beyond
        case 0: y = yv[0]; break; case 1: y = yv[1]; break; case 2: y = yv[2]; break;
        case 3: y = yv[3]; break; case 4: y = yv[4]; break; case 5: y = yv[5]; break;
        default: y = __ldg(reinterpret_cast<const float4*>(Y + idx)); break;
      }
      o.x = fmaxf((y.x - mu.x) * is.x * ga.x + be.x, 0.f);
      o.y = fmaxf((y.y - mu.y) * is.y * ga.y + be.y, 0.f);
      o.z = fmaxf((y.z - mu.z) * is.z * ga.z + be.z, 0.f);
      o.w = fmaxf((y.w - mu.w) * is.w * ga.w + be.w, 0.f);
      if (drop) {
        bool k[4];
        dropout_keep4(ph, off, rng_stream, (unsigned long long)idx, p_drop, k);
        o.x = k[0] ? o.x * scale : 0.f; o.y = k[1] ? o.y * scale : 0.f;
        o.z = k[2] ? o.z * scale : 0.f; o.w = k[3] ? o.w * scale : 0.f;
      }
    }
    *reinterpret_cast<float4*>(X + idx) = o;
  }
}

// BatchNorm finalize from already reduced (and possibly all-reduced) sums
__global__ void __launch_bounds__(256) bn_finalize_kernel(LayerDev L, const float* __restrict__ ball,
                                                          const double* __restrict__ sums, float* __restrict__ mean,
                                                          float* __restrict__ invstd, int training, double M,
                                                          double eps, double momentum) {
  pdl_prologue();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= L.fo_tot) return;
  bn_finalize_channel(L, ball, c, training ? sums[c] : 0.0, training ? sums[L.fo_tot + c] : 0.0, mean, invstd, training, M,
                      eps, momentum);
}

// X = dropout(relu((Y - mean) * invstd * gamma + beta)); one thread per 4 channels (scalar tail-safe)
__global__ void __launch_bounds__(256) bn_apply_kernel(PlanDev p, const float* __restrict__ Y,
                                                       const float* __restrict__ ball, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, float* __restrict__ X, int C,
                                                       int training, float p_drop, const unsigned long long* rng,
                                                       unsigned long long rng_stream) {
  pdl_prologue();
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;   // one element per thread
  const long long total = (long long)p.t_cap * C;
  if (idx >= total) return;
  const int t = (int)(idx / C), c = (int)(idx - (long long)t * C);
  if (t >= T) { X[idx] = 0.0f; return; }
  const float g = ball[C + c], b = ball[2 * C + c];
  float x = (Y[idx] - mean[c]) * invstd[c] * g + b;
  x = fmaxf(x, 0.0f);
  if (training && p_drop > 0.0f) {
    const Philox ph(rng[0]);
    x = dropout_keep(ph, rng[1], rng_stream, (unsigned long long)idx, p_drop) ? x * (1.0f / (1.0f - p_drop)) : 0.0f;
  }
  X[idx] = x;
}

// float4 form (C % 4 == 0): grid (ceil(C/4/128), ceil(t_cap/kEltRows)); a thread owns 4 channels for kEltRows rows
__global__ void __launch_bounds__(128) bn_apply_vec_kernel(PlanDev p, const float* __restrict__ Y,
                                                           const float* __restrict__ ball, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, float* __restrict__ X, int C,
                                                           int training, float p_drop, const unsigned long long* rng,
                                                           unsigned long long rng_stream) {
  pdl_prologue();
  const int c4 = blockIdx.x * 128 + threadIdx.x;
  if (c4 * 4 >= C) return;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const int c = c4 * 4;
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
  const float4 ga = *reinterpret_cast<const float4*>(ball + C + c), be = *reinterpret_cast<const float4*>(ball + 2 * C + c);
  const bool drop = training && p_drop > 0.0f;
  const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  unsigned long long seed = 0, off = 0;
  if (drop) { seed = rng[0]; off = rng[1]; }
  const Philox ph(seed);
  const int r0 = blockIdx.y * kEltRows, r1 = min(p.t_cap, r0 + kEltRows);
#pragma unroll 4
  for (int t = r0; t < r1; ++t) {
    const size_t idx = (size_t)t * C + c;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T) {
      const float4 y = __ldg(reinterpret_cast<const float4*>(Y + idx));
      o.x = fmaxf((y.x - mu.x) * is.x * ga.x + be.x, 0.f);
      o.y = fmaxf((y.y - mu.y) * is.y * ga.y + be.y, 0.f);
      o.z = fmaxf((y.z - mu.z) * is.z * ga.z + be.z, 0.f);
      o.w = fmaxf((y.w - mu.w) * is.w * ga.w + be.w, 0.f);
      if (drop) {
        bool k[4];
        dropout_keep4(ph, off, rng_stream, (unsigned long long)idx, p_drop, k);
        o.x = k[0] ? o.x * scale : 0.f; o.y = k[1] ? o.y * scale : 0.f;
        o.z = k[2] ? o.z * scale : 0.f; o.w = k[3] ? o.w * scale : 0.f;
      }
    }
    *reinterpret_cast<float4*>(X + idx) = o;
  }
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(long long total, float p_drop,
                                                           const unsigned long long* rng, unsigned long long rng_stream,
                                                           uint8_t* __restrict__ keep) {
  pdl_prologue();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const Philox ph(rng[0]);
  keep[idx] = dropout_keep(ph, rng[1], rng_stream, (unsigned long long)idx, p_drop) ? 1 : 0;
}

static bool vec4_ok(const eagcn_layer_t* l) {
  if (l->fo_tot % 4) return false;
  for (int v = 0; v < l->V; ++v) if ((l->fo[v] % 4) || (l->off[v] % 4)) return false;
  return true;
}

bool layer_ok(const eagcn_plan_t* plan, const eagcn_layer_t* l) {
  if (!l || l->V != plan->V || l->fin <= 0 || l->fo_tot <= 0) return false;
  long long s = 0;
  for (int v = 0; v < l->V; ++v) {
    if (l->fo[v] <= 0 || l->off[v] != s) return false;
    s += l->fo[v];
    if (!l->att_w[v] || !l->self_r[v] || !l->W[v] || !l->bias[v] || !l->gamma[v] || !l->beta[v] || !l->run_mean[v] ||
        !l->run_var[v])
      return false;
    if (plan->chan[v] > 254) return false;
  }
  return s == l->fo_tot;
}

// wall / wallT / wsplit / ball / sig of one layer from the module parameters (they change every optimiser step)
static int launch_prep_params(const LayerDev& L, const eagcn_work_t* w, cudaStream_t st) {
  const int C = L.fo_tot;
  long long n = (long long)L.fin * C;
  if (n < (long long)L.V * EAGCN_SIG_STRIDE) n = (long long)L.V * EAGCN_SIG_STRIDE;
  if (n < C) n = C;
  EAGCN_PROF("prep_params_kernel", st);
  EAGCN_LAUNCH(prep_params_kernel, (unsigned)((n + 255) / 256), 256, 0, st)(L, (float*)w->wall, (float*)w->wallT, (float*)w->wsplit, (float*)w->ball, (float*)w->sig);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

}  // namespace eagcn
using namespace eagcn;

extern "C" int eagcn_layer_prepare(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w, void* stream) {
  // plan: only V and chan[] are read (the sigmoid tables have C_v entries); no plan array is touched
  if (!plan || plan->V <= 0 || plan->V > EAGCN_MAX_VIEWS || !w || !layer || layer->V != plan->V || layer->fin <= 0 ||
      layer->fo_tot <= 0)
    return EAGCN_E_ARG;
  if (!w->wall || !w->ball || !w->sig) return EAGCN_E_ARG;
  for (int v = 0; v < layer->V; ++v)
    if (!layer->att_w[v] || !layer->self_r[v] || !layer->W[v] || !layer->bias[v] || !layer->gamma[v] || !layer->beta[v] ||
        plan->chan[v] > 254)
      return EAGCN_E_ARG;
  LayerDev L = to_dev(layer, plan);
  return launch_prep_params(L, w, (cudaStream_t)stream);
}

extern "C" int64_t eagcn_stat_tiles(int64_t t_cap) { return (t_cap + kStatRows - 1) / kStatRows; }

extern "C" int eagcn_layer_forward_a(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w,
                                     void* stream) {
  if (!plan_ok(plan) || !w || !layer_ok(plan, layer)) return EAGCN_E_ARG;
  if (!w->H || !w->Z || !w->Y || !w->invR || !w->wall || !w->ball || !w->sig || !w->partial || !w->sums)
    return EAGCN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PlanDev p = to_dev(plan);
  LayerDev L = to_dev(layer, plan);
  const int C = L.fo_tot;
  int rc;
  if (!(w->training & 4)) {              // bit2: eagcn_layer_prepare already ran for these buffers (on any stream)
    rc = launch_prep_params(L, w, st);
    if (rc) return rc;
  }
  const int want = (w->training & 1) ? 1 : 0;
  const bool host_allreduce = (w->training & 2) != 0;   // global-batch BatchNorm: host sums the partial sums over ranks
  if (fz::fwd_fused_ok(plan, layer, w, gemm_mode(), agg_mode())) {
    // projection + score + normalise + aggregate + bias + statistics partials in ONE launch (layer_fused.cu)
    rc = fz::layer_fwd_fused(plan, p, L, w, st);
    if (rc) return rc;
    if (want && host_allreduce) {
      StatEpilogue ep{0, nullptr, nullptr, nullptr, nullptr, 0, 0.0, 0.0, 0.0};
      EAGCN_PROF("stat_reduce_kernel", st);
      EAGCN_LAUNCH(stat_reduce_kernel, (C + 31) / 32, 32 * kStatLanes, 0, st)(p, L, (const float*)w->partial, (double*)w->sums, C, ep, 1);
      EAGCN_LAUNCH_CHECK();
    }
    return 0;
  }
  if (gemm_mode() != 1 && w->wallT && tc::tc_supported((const float*)w->H, L.fin, (const float*)w->wallT, L.fin, L.fin))
    rc = tc::gemm_tc_nt((const float*)w->H, L.fin, (const float*)w->wallT, L.fin, (float*)w->Z, C, p.t_cap, C, L.fin,
                        p.counts + EAGCN_CNT_T, st, "gemm_tc_nn", (const float*)w->wallT + (size_t)L.fin * C);
  else
    rc = gemm_nn((const float*)w->H, L.fin, (const float*)w->wall, C, (float*)w->Z, C, p.t_cap, C, L.fin,
                 p.counts + EAGCN_CNT_T, st);
  if (rc) return rc;
  dim3 grid((unsigned)eagcn_stat_tiles(p.t_cap), L.V);
  const int n_pad = (int)(w->n_pad > 0 ? w->n_pad : plan->N);
  int fo_max = 0;
  for (int v = 0; v < L.V; ++v) fo_max = L.fo[v] > fo_max ? L.fo[v] : fo_max;
  if (agg_mode() == 0 && vec4_ok(layer) && fo_max <= kTileMaxFo && aligned16(w->Z) && aligned16(w->Y) && aligned16(w->ball)) {
    size_t smem = 0;
    for (int v = 0; v < L.V; ++v) smem = agg_fwd_tile_smem(L.fo[v]) > smem ? agg_fwd_tile_smem(L.fo[v]) : smem;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(agg_fwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)agg_fwd_tile_smem(kTileMaxFo));
      if (e != cudaSuccess) return (int)e;
      attr_set = true;
    }
    EAGCN_PROF("agg_fwd_kernel", st);
    EAGCN_LAUNCH(agg_fwd_tile_kernel, grid, kAggThreads, smem, st)(p, L, (const float*)w->Z, (const float*)w->ball,
                                                         (const float*)w->sig, (float*)w->Y, (float*)w->invR,
                                                         (float*)w->partial, n_pad, want);
  } else if (vec4_ok(layer) && fo_max <= 256) {
    EAGCN_PROF("agg_fwd_kernel", st);
    if (fo_max <= 128)
      EAGCN_LAUNCH((agg_fwd_onepass_kernel<1>), grid, kAggThreads, 0, st)(p, L, (const float*)w->Z, (const float*)w->ball,
                                                              (const float*)w->sig, (float*)w->Y, (float*)w->invR,
                                                              (float*)w->partial, n_pad, want);
    else
      EAGCN_LAUNCH((agg_fwd_onepass_kernel<2>), grid, kAggThreads, 0, st)(p, L, (const float*)w->Z, (const float*)w->ball,
                                                              (const float*)w->sig, (float*)w->Y, (float*)w->invR,
                                                              (float*)w->partial, n_pad, want);
  } else if (vec4_ok(layer)) {
    EAGCN_PROF("agg_fwd_kernel", st);
    EAGCN_LAUNCH((agg_fwd_kernel<4>), grid, kAggThreads, 0, st)(p, L, (const float*)w->Z, (const float*)w->ball, (const float*)w->sig,
                                            (float*)w->Y, (float*)w->invR, (float*)w->partial, n_pad, want);
  } else {
    EAGCN_PROF("agg_fwd_kernel", st);
    EAGCN_LAUNCH((agg_fwd_kernel<1>), grid, kAggThreads, 0, st)(p, L, (const float*)w->Z, (const float*)w->ball, (const float*)w->sig,
                                            (float*)w->Y, (float*)w->invR, (float*)w->partial, n_pad, want);
  }
  EAGCN_LAUNCH_CHECK();
  if (want && host_allreduce) {
    StatEpilogue ep{0, nullptr, nullptr, nullptr, nullptr, 0, 0.0, 0.0, 0.0};
    EAGCN_PROF("stat_reduce_kernel", st);
    EAGCN_LAUNCH(stat_reduce_kernel, (C + 31) / 32, 32 * kStatLanes, 0, st)(p, L, (const float*)w->partial, (double*)w->sums, C, ep, 0);
    EAGCN_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int eagcn_layer_forward_b(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w,
                                     void* stream) {
  if (!plan_ok(plan) || !w || !layer_ok(plan, layer)) return EAGCN_E_ARG;
  if (!w->Y || !w->X || !w->ball || !w->sums || !w->mean || !w->invstd) return EAGCN_E_ARG;
  if ((w->training & 1) && w->p_drop > 0.0 && !w->rng) return EAGCN_E_ARG;
  if (w->p_drop < 0.0 || w->p_drop >= 1.0) return EAGCN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PlanDev p = to_dev(plan);
  LayerDev L = to_dev(layer, plan);
  const int C = L.fo_tot;
  const double M = (double)(w->m_total > 0 ? w->m_total : plan->B * plan->N);
  const int training = (w->training & 1) ? 1 : 0;
  // how part A indexed the statistics partials (same pure eligibility test as eagcn_layer_forward_a)
  const int aligned = (w->H && w->Z && fz::fwd_fused_ok(plan, layer, w, gemm_mode(), agg_mode())) ? 1 : 0;
  if (training && !(w->training & 2) && fuse_mode() == 0 && (C & 3) == 0 && aligned16(w->Y) && aligned16(w->X) &&
      aligned16(w->ball)) {
    // per-replica statistics, float4 layout: reduction of the tile partials + finalize + normalise/ReLU/dropout in ONE launch
    if (!w->partial) return EAGCN_E_ARG;
    const int nblk = (C + 31) / 32;
    int R = 148 / nblk;                   // 1024-thread CTAs at ~60 registers: one per SM -> a single wave
    R = R < 1 ? 1 : (R > 16 ? 16 : R);
    EAGCN_PROF("bn_stat_apply_kernel", st);
    EAGCN_LAUNCH(bn_stat_apply_kernel, dim3(nblk, R), 32 * kStatLanes, 0, st)(
        p, L, (const float*)w->partial, (double*)w->sums, (const float*)w->Y, (const float*)w->ball, (float*)w->mean,
        (float*)w->invstd, (float*)w->X, C, M, w->eps, w->momentum, (float)w->p_drop, (const unsigned long long*)w->rng,
        (unsigned long long)w->rng_stream, aligned);
    EAGCN_LAUNCH_CHECK();
    return 0;
  }
  if (training && !(w->training & 2)) {
    // per-replica statistics: reduce the tile partials and finalize in one kernel
    if (!w->partial) return EAGCN_E_ARG;
    StatEpilogue ep{1, (const float*)w->ball, (float*)w->mean, (float*)w->invstd, nullptr, 1, M, w->eps, w->momentum};
    EAGCN_PROF("stat_reduce_kernel", st);
    EAGCN_LAUNCH(stat_reduce_kernel, (C + 31) / 32, 32 * kStatLanes, 0, st)(p, L, (const float*)w->partial, (double*)w->sums, C, ep, aligned);
    EAGCN_LAUNCH_CHECK();
  } else {
    EAGCN_PROF("bn_finalize_kernel", st);
    EAGCN_LAUNCH(bn_finalize_kernel, (C + 255) / 256, 256, 0, st)(L, (const float*)w->ball, (const double*)w->sums, (float*)w->mean,
                                                        (float*)w->invstd, training, M, w->eps, w->momentum);
    EAGCN_LAUNCH_CHECK();
  }
  const long long total = (long long)p.t_cap * C;
  if ((C & 3) == 0 && aligned16(w->Y) && aligned16(w->X)) {
    dim3 grid((C / 4 + 127) / 128, (p.t_cap + kEltRows - 1) / kEltRows);
    EAGCN_PROF("bn_apply_kernel", st);
    EAGCN_LAUNCH(bn_apply_vec_kernel, grid, 128, 0, st)(p, (const float*)w->Y, (const float*)w->ball, (const float*)w->mean,
                                              (const float*)w->invstd, (float*)w->X, C, training,
                                              (float)w->p_drop, (const unsigned long long*)w->rng,
                                              (unsigned long long)w->rng_stream);
    EAGCN_LAUNCH_CHECK();
  } else {
    EAGCN_PROF("bn_apply_kernel", st);
    EAGCN_LAUNCH(bn_apply_kernel, (unsigned)((total + 255) / 256), 256, 0, st)(
        p, (const float*)w->Y, (const float*)w->ball, (const float*)w->mean, (const float*)w->invstd, (float*)w->X, C,
        training, (float)w->p_drop, (const unsigned long long*)w->rng, (unsigned long long)w->rng_stream);
    EAGCN_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int eagcn_dropout_mask(const eagcn_plan_t* plan, const eagcn_work_t* w, int64_t fo_tot, void* keep_out,
                                  void* stream) {
  if (!plan_ok(plan) || !w || !w->rng || !keep_out || fo_tot <= 0) return EAGCN_E_ARG;
  const long long total = (long long)plan->t_cap * fo_tot;
  EAGCN_PROF("dropout_mask_kernel", (cudaStream_t)stream);
  EAGCN_LAUNCH(dropout_mask_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream)(
      total, (float)w->p_drop, (const unsigned long long*)w->rng, (unsigned long long)w->rng_stream, (uint8_t*)keep_out);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int eagcn_dropout_mask_flat(const void* rng, int64_t rng_stream, double p_drop, int64_t total, void* keep_out,
                                       void* stream) {
  if (!rng || !keep_out || total <= 0 || p_drop < 0.0 || p_drop >= 1.0) return EAGCN_E_ARG;
  EAGCN_PROF("dropout_mask_kernel", stream);
  EAGCN_LAUNCH(dropout_mask_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream)(
      total, (float)p_drop, (const unsigned long long*)rng, (unsigned long long)rng_stream, (uint8_t*)keep_out);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
