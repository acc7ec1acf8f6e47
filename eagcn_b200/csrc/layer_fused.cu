// ONE kernel for part A of GraphConv_Layer.forward (reference layers.py:81-92 for all five views):
//
//   projection      Z = H . W_all                      layers.py:40    tcgen05 (3xTF32, fp32-faithful), TMA-fed, TMEM
//   score           s_e = sigmoid(a_v[type_v(e)])       layers.py:82-83 table lookup on the uint8 edge codes
//   normalise       R_i = sum_e s_e + sigmoid(r_v) + (N - deg_i) * 1e-9 ;  A = s / R      layers.py:84-90
//   aggregate+bias  Y_v[i] = sum_e A_e Z_v[j_e] + A_self Z_v[i] + b_v                     layers.py:39,43
//   BatchNorm partial sums of (Y - b) per row tile                                           layers.py:408-412
//
// One CTA = one MOLECULE-ALIGNED row tile (plan.tile_row: whole molecules, <= 128 packed rows) x one chunk of BN <= 256
// output channels.  Warp roles:
//   warp 0       TMA producer (H rows of the tile + the pre-split weight chunk, k-block by k-block)
//   warp 1       single-thread tcgen05.mma issuer, accumulators in TMEM (main + hi/lo correction)
//   warps 2..9   hi/lo split of the activation tile during the main loop; afterwards the epilogue: TMEM -> registers ->
//                a [128][BN] tile of Z in the (now idle) pipeline shared memory, aggregation over it, Y, statistics
//   warps 10..13 "graph warps": everything of the layer that does NOT depend on the projection runs WHILE the tensor
//                cores work -- edge codes -> sigmoid lookup, attention row sums, normalised edge weights, 1/R -- and,
//                once the Z tile is in shared memory, its one store to HBM (the activation backward needs)
// Because a tile holds whole molecules every neighbour row of every row is in the CTA's own Z tile: the aggregation
// reads nothing but shared memory, Z is never re-read from HBM in forward, and in inference (bit4 of work.training)
// it is not written at all.  Replaces the round-1 pair gemm_tc (Z to HBM) + agg_fwd_tile (Z back from HBM).
#include "common.cuh"

namespace eagcn {
namespace fz {
using namespace tc;

constexpr int kEdgeCap = 1024;       // edges of a run of rows staged at a time (a tile normally has ~300)
constexpr int kEpiThreads = 256;     // warps 2..9
constexpr int kGraphThreads = 128;   // warps 10..13
constexpr int kFusedThreads = 64 + kEpiThreads + kGraphThreads;
constexpr int kRedFloats = 2 * (2 * 1024 + 64);   // nrg * ncols <= 1024 (+ slack) doubles, twice (sum, sum of squares)

struct FusedArgs {
  int C, K, BN, stages;
  uint32_t tmem_cols;
  int n_pad, want_stats, save_z;
  int nv_max;                        // most views any column chunk overlaps (sizes the per-view edge-weight staging)
  int passes;                        // 3: fp32-faithful 3xTF32 (default); 1: single TF32 pass (tc::tc_passes())
  uint32_t graph_off;                // byte offset of the graph warps' staging region in dynamic shared memory
  const float* ball;
  const float* sig;
  float* Z; float* Y; float* invR; float* partial;
};

// named barriers (0 = __syncthreads): producers bar.arrive, consumers bar.sync
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void graph_sync() { asm volatile("bar.sync 5, 128;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_id(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_id(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int kBarPrep = 2, kBarUsed = 3, kBarZs = 4, kBarBoth = kEpiThreads + kGraphThreads;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

__global__ void __launch_bounds__(kFusedThreads, 1)
layer_fwd_fused_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                       const __grid_constant__ CUtensorMap mapB2, PlanDev p, LayerDev L, FusedArgs g) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_ready[MAX_STAGES], bar_empty[MAX_STAGES], bar_acc;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int s_rp[EAGCN_ROW_TILE + 1];

  const int tile = blockIdx.y;
  // three independent loads (tile_row has t_cap/32 + 2 entries, more than the grid's tile bound: always readable)
  const int ntiles = p.counts[EAGCN_CNT_TILES];
  const int row0 = p.tile_row[tile];
  const int row1 = p.tile_row[tile + 1];
  if (tile >= ntiles) return;                                      // grid.y is a host-side upper bound
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nr = min(row1 - row0, EAGCN_ROW_TILE);                 // owned rows (whole molecules)
  const int n0 = blockIdx.x * g.BN;
  const int bk = BK;
  const int num_kb = (g.K + bk - 1) / bk;

  // 1024-byte aligned start (SWIZZLE_128B atoms) as an OFFSET into the __shared__ array: an integer round trip of the
  // pointer would lose the address space and turn every access below into a generic LD.E / ST.E instead of LDS / STS
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t bytesA = BM * bk * 4, bytesB = (uint32_t)g.BN * bk * 4;
  const bool one = g.passes == 1;                       // single TF32 pass: [A | B] per stage, no lo copies
  const uint32_t stage_bytes = one ? bytesA + bytesB : 2 * bytesA + 2 * bytesB;
  auto sA_hi = [&](int s) { return base + (size_t)s * stage_bytes; };
  auto sA_lo = [&](int s) { return base + (size_t)s * stage_bytes + bytesA; };
  auto sB_hi = [&](int s) { return base + (size_t)s * stage_bytes + (one ? bytesA : 2 * bytesA); };
  auto sB_lo = [&](int s) { return base + (size_t)s * stage_bytes + 2 * bytesA + bytesB; };
  // graph warps' region (NOT overlapping the pipeline stages: filled while the main loop runs)
  int* s_col = reinterpret_cast<int*>(base + g.graph_off);                      // [kEdgeCap] neighbour row inside the tile
  float* s_w = reinterpret_cast<float*>(s_col + kEdgeCap);                      // [nv_max][kEdgeCap] normalised edge weight
  float* s_aself = s_w + (size_t)g.nv_max * kEdgeCap;                           // [nv_max][128]  sigmoid(r_v) / R
  float* s_sig = s_aself + (size_t)g.nv_max * EAGCN_ROW_TILE;                   // [nv_max][257]  sigmoid tables of the chunk's views

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_ready[s], kXformThreads / 32);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x - 64 <= nr) s_rp[threadIdx.x - 64] = p.row_ptr[row0 + threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_smem;
  const uint32_t tmem_c = tmem_d + (uint32_t)g.BN;

  const int ncols = min(g.BN, g.C - n0);           // live columns of this chunk (multiple of 4)
  // views overlapping [n0, n0 + ncols): a contiguous range [va, vb]
  int va = 0;
  while (va + 1 < L.V && L.off[va + 1] <= n0) ++va;
  int vb = va;
  while (vb + 1 < L.V && L.off[vb + 1] < n0 + ncols) ++vb;
  const int nv = vb - va + 1;
  const int ZLD = g.BN + 4;                         // row stride of the Z tile: conflict-free float4 rows
  float* Zs = reinterpret_cast<float*>(base);       // [128][ZLD], over the pipeline stages once the main loop is done

  if (warp == 0) {
    // ===================== TMA producer: H rows [row0, row0+128) and the weight chunk, k-block by k-block ============
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (kb >= g.stages) mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_expect_tx(&bar_full[s], bytesA + (one ? 1 : 2) * bytesB);
        tma_load_2d(&mapA, &bar_full[s], sA_hi(s), kb * bk, row0);
        tma_load_2d(&mapB, &bar_full[s], sB_hi(s), kb * bk, n0);
        if (!one) tma_load_2d(&mapB2, &bar_full[s], sB_lo(s), kb * bk, n0);
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&bar_ready[s], ph);
        tc_fence_after();
        const uint64_t dAh = make_desc(smem_u32(sA_hi(s)), bk), dAl = make_desc(smem_u32(sA_lo(s)), bk);
        const uint64_t dBh = make_desc(smem_u32(sB_hi(s)), bk), dBl = make_desc(smem_u32(sB_lo(s)), bk);
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          const uint64_t adv = (uint64_t)((k * UK * 4) >> 4);
          const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
          if (!one) {
            umma_tf32(tmem_c, dAl + adv, dBh + adv, idesc, acc);
            umma_tf32(tmem_c, dAh + adv, dBl + adv, idesc, 1u);
          }
          umma_tf32(tmem_d, dAh + adv, dBh + adv, idesc, acc);
        }
        umma_commit(&bar_empty[s]);
        if (kb == num_kb - 1) umma_commit(&bar_acc);
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 2 + kEpiThreads / 32) {
    // ===================== graph warps: attention weights of the tile, concurrently with the projection =============
    const int gt = threadIdx.x - 64 - kEpiThreads;                 // 0..127
    for (int i = gt; i < nv * EAGCN_SIG_STRIDE; i += kGraphThreads) s_sig[i] = __ldg(g.sig + va * EAGCN_SIG_STRIDE + i);
    int r0 = 0;
    bool first = true;
    while (r0 < nr) {
      int r1 = nr;                                 // longest run of rows whose edges fit the staging arrays
      if (s_rp[nr] - s_rp[r0] > kEdgeCap) {
        r1 = r0 + 1;
        while (r1 < nr && s_rp[r1 + 1] - s_rp[r0] <= kEdgeCap) ++r1;
      }
      const int eA = s_rp[r0], nE = min(s_rp[r1] - eA, kEdgeCap), nR = r1 - r0;
      if (!first) bar_sync_id(kBarUsed, kBarBoth);                 // the epilogue warps are done with the previous run
      graph_sync();                                                // sigmoid tables staged (first run) / run boundary
      for (int i = gt; i < nE; i += kGraphThreads) {
        s_col[i] = min(max(p.col[eA + i] - row0, 0), EAGCN_ROW_TILE - 1);
        for (int q = 0; q < nv; ++q)               // sigmoid(a_v[type_v(e)])  (layers.py:82-83)
          s_w[q * kEdgeCap + i] = s_sig[q * EAGCN_SIG_STRIDE + p.code[(size_t)(va + q) * p.e_cap + eA + i]];
      }
      graph_sync();
      for (int it = gt; it < nR * nv; it += kGraphThreads) {       // one thread per (row, view): row sum in edge order (layers.py:84,87)
        const int q = it / nR, r = r0 + (it - q * nR);
        const int v = va + q;
        const float sig_r = s_sig[q * EAGCN_SIG_STRIDE + 256];
        float* wq = s_w + q * kEdgeCap;
        const int a0 = s_rp[r] - eA, a1 = min(s_rp[r + 1] - eA, kEdgeCap);
        float sw = 0.0f;
        for (int e = a0; e < a1; ++e) sw += wq[e];
        const float R = sw + sig_r + (float)(g.n_pad - (a1 - a0)) * EAGCN_TINY;
        s_aself[q * EAGCN_ROW_TILE + r] = sig_r / R;
        if (L.off[v] >= n0) g.invR[(size_t)v * p.t_cap + row0 + r] = 1.0f / R;     // the chunk holding the view's first column
        for (int e = a0; e < a1; ++e) wq[e] = wq[e] / R;
      }
      bar_arrive_id(kBarPrep, kBarBoth);                           // this run's weights are ready for the epilogue warps
      first = false;
      r0 = r1;
    }
    // ---- Z -> HBM once, for the backward pass (coalesced: a warp writes whole rows, two rows in flight) ----
    bar_sync_id(kBarZs, kBarBoth);
    if (g.save_z) {
      const int n4 = ncols >> 2, gw = gt >> 5;
      for (int r = gw; r < nr; r += 2 * (kGraphThreads / 32)) {
        const int rb = r + kGraphThreads / 32;
        const bool two = rb < nr;
        const float4* src0 = reinterpret_cast<const float4*>(Zs + r * ZLD);
        const float4* src1 = reinterpret_cast<const float4*>(Zs + (two ? rb : r) * ZLD);
        float4* dst0 = reinterpret_cast<float4*>(g.Z + (size_t)(row0 + r) * g.C + n0);
        float4* dst1 = reinterpret_cast<float4*>(g.Z + (size_t)(row0 + rb) * g.C + n0);
        for (int c = lane; c < n4; c += 32) {
          const float4 x0 = src0[c];
          const float4 x1 = src1[c];
          dst0[c] = x0;
          if (two) dst1[c] = x1;
        }
      }
    }
  } else {
    // ===================== hi/lo split of the activation tile, then the layer epilogue =====================
    const int et = threadIdx.x - 64;
    {
      int s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&bar_full[s], ph);
        if (!one) split_tile(reinterpret_cast<uint4*>(sA_hi(s)), reinterpret_cast<uint4*>(sA_lo(s)), (int)(bytesA / 16), et);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_ready[s]);
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
    }
    mbar_wait(&bar_acc, 0);                       // all MMAs retired: accumulators complete, pipeline buffers idle
    tc_fence_after();
    float* s_redf = Zs + EAGCN_ROW_TILE * ZLD;    // statistics exchange: [row groups][2][ncols] doubles (8-byte aligned: ZLD % 4 == 0)
    // ---- TMEM -> shared memory: Z tile = main + correction accumulator ----
    {
      const int quad = warp & 3, half = (warp - 2) >> 2;
      const int row = quad * 32 + lane;
      for (int c0 = half * 32; c0 < g.BN; c0 += 64) {
        uint32_t r[32], q[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
        tmem_ld32(taddr, r);
        if (!one) tmem_ld32(taddr + (uint32_t)g.BN, q);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (one) {
#pragma unroll
          for (int j = 0; j < 32; ++j) q[j] = 0u;                 // +0.0f: no correction accumulator in the single-pass mode
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (c0 + 4 * j < g.BN) {
            float4 o;
            o.x = __uint_as_float(r[4 * j]) + __uint_as_float(q[4 * j]);
            o.y = __uint_as_float(r[4 * j + 1]) + __uint_as_float(q[4 * j + 1]);
            o.z = __uint_as_float(r[4 * j + 2]) + __uint_as_float(q[4 * j + 2]);
            o.w = __uint_as_float(r[4 * j + 3]) + __uint_as_float(q[4 * j + 3]);
            *reinterpret_cast<float4*>(Zs + row * ZLD + c0 + 4 * j) = o;
          }
        }
      }
    }
    tc_fence_before();
    bar_arrive_id(kBarZs, kBarBoth);               // the graph warps may start the store of Z
    epi_sync();
    // ---- aggregation, bias, BatchNorm partials: (row group, float4 column) items over all views of the chunk ----
    const int nc4 = ncols >> 2;                                       // float4 columns of the chunk (<= 64)
    const int nrg = kEpiThreads / nc4;                                // row groups
    const int rg = et / nc4, c4 = et - rg * nc4;
    const bool worker = rg < nrg;
    int vi = 0;                                                       // view of this thread's float4 column
    if (worker) { while (va + vi < vb && L.off[va + vi + 1] <= n0 + 4 * c4) ++vi; }
    const float* wrow = s_w + (size_t)vi * kEdgeCap;
    const float* asf = s_aself + vi * EAGCN_ROW_TILE;
    const float* zcol = Zs + 4 * c4;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (worker) b4 = __ldg(reinterpret_cast<const float4*>(g.ball + n0) + c4);
    double d1[4] = {0.0, 0.0, 0.0, 0.0}, d2[4] = {0.0, 0.0, 0.0, 0.0};
    int r0 = 0;
    while (r0 < nr) {
      int r1 = nr;                                 // the same runs of rows as the graph warps
      if (s_rp[nr] - s_rp[r0] > kEdgeCap) {
        r1 = r0 + 1;
        while (r1 < nr && s_rp[r1 + 1] - s_rp[r0] <= kEdgeCap) ++r1;
      }
      const int eA = s_rp[r0];
      bar_sync_id(kBarPrep, kBarBoth);             // this run's normalised weights are staged
      if (worker) {
        // aggregate (layers.py:90,39) + bias (layers.py:43): four rows per thread in flight and NO data-dependent
        // branches in the edge walk (a finished row keeps loading edge 0 of the run with weight 0) -- with one CTA per
        // SM the shared-memory latency is hidden by instruction-level parallelism, not by occupancy
        for (int rb = r0 + rg; rb < r1; rb += 4 * nrg) {
          int ea[4], ne[4];
          float acc[4][4];
          int kmax = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = min(rb + u * nrg, r1 - 1);
            const bool live = rb + u * nrg < r1;
            ea[u] = s_rp[r] - eA;
            ne[u] = live ? min(s_rp[r + 1] - eA, kEdgeCap) - ea[u] : 0;
            const float a_self = live ? asf[r] : 0.0f;
            const float4 zt = *reinterpret_cast<const float4*>(zcol + r * ZLD);
            acc[u][0] = a_self * zt.x; acc[u][1] = a_self * zt.y; acc[u][2] = a_self * zt.z; acc[u][3] = a_self * zt.w;
            kmax = max(kmax, ne[u]);
          }
          for (int k = 0; k < kmax; ++k) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const bool on = k < ne[u];
              const int e = on ? ea[u] + k : 0;
              const float a = on ? wrow[e] : 0.0f;
              const float4 zj = *reinterpret_cast<const float4*>(zcol + s_col[e] * ZLD);
              acc[u][0] = fmaf(a, zj.x, acc[u][0]); acc[u][1] = fmaf(a, zj.y, acc[u][1]);
              acc[u][2] = fmaf(a, zj.z, acc[u][2]); acc[u][3] = fmaf(a, zj.w, acc[u][3]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = rb + u * nrg;
            if (r < r1) {
              reinterpret_cast<float4*>(g.Y + (size_t)(row0 + r) * g.C + n0)[c4] =
                  make_float4(acc[u][0] + b4.x, acc[u][1] + b4.y, acc[u][2] + b4.z, acc[u][3] + b4.w);
              if (g.want_stats) {
#pragma unroll
                for (int x = 0; x < 4; ++x) { const double a = (double)acc[u][x]; d1[x] += a; d2[x] = fma(a, a, d2[x]); }
              }
            }
          }
        }
      }
      if (r1 < nr) bar_arrive_id(kBarUsed, kBarBoth);              // (dense graphs only) the graph warps may stage the next run
      r0 = r1;
    }
    if (g.want_stats) {
      // BatchNorm partial sums of this tile (layers.py:408-412) in DOUBLE: sum and sum of squares of (Y - b).  The
      // variance is later formed as S2/M - (S1/M)^2; fp32 partials would leave it good to ~6e-8 * mean^2 only, which
      // near-constant channels (var << mean^2, e.g. fully connected graphs) amplify beyond the parity bar.
      double* s_red = reinterpret_cast<double*>(s_redf);              // [row groups][2][ncols]
      if (worker) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          s_red[(rg * 2 + 0) * ncols + 4 * c4 + x] = d1[x];
          s_red[(rg * 2 + 1) * ncols + 4 * c4 + x] = d2[x];
        }
      }
      epi_sync();
      double* part = reinterpret_cast<double*>(g.partial);
      for (int c = et; c < ncols; c += kEpiThreads) {
        double a = 0.0, b = 0.0;
        for (int q = 0; q < nrg; ++q) { a += s_red[(q * 2 + 0) * ncols + c]; b += s_red[(q * 2 + 1) * ncols + c]; }
        part[((size_t)tile * 2 + 0) * g.C + n0 + c] = a;
        part[((size_t)tile * 2 + 1) * g.C + n0 + c] = b;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(g.tmem_cols) : "memory");
  }
}

inline int& fwd_fused_mode() { static int m = 1; return m; }

// column chunking: as few chunks as the 256-column TMEM limit allows when the row tiles alone fill the GPU, more
// (narrower) chunks when they do not -- the epilogue of a chunk is serial work of one CTA
static int pick_chunk(int C, int t_cap) {
  const int tiles = t_cap / 118 + 1;
  int nch = 148 / tiles;
  const int lo = (C + 255) / 256, hi = (C + 63) / 64;
  nch = nch < lo ? lo : (nch > hi ? hi : nch);
  int bn = (((C + nch - 1) / nch) + 15) & ~15;
  return bn > 256 ? 256 : bn;
}

static int max_views_per_chunk(const LayerDev& L, int BN) {
  int best = 1;
  for (int n0 = 0; n0 < L.fo_tot; n0 += BN) {
    int n = 0;
    for (int v = 0; v < L.V; ++v)
      if (L.off[v] < n0 + BN && L.off[v] + L.fo[v] > n0) ++n;
    best = n > best ? n : best;
  }
  return best;
}

static size_t graph_bytes(int nv_max) {
  return (size_t)kEdgeCap * 4 + (size_t)nv_max * (kEdgeCap + EAGCN_ROW_TILE + EAGCN_SIG_STRIDE) * 4 + 16;
}
static size_t front_bytes(int BN, int stages) {                    // pipeline stages, re-used by the Z tile + statistics
  const size_t pipe = (size_t)stages * (tc_passes() == 1 ? 1 : 2) * (BM * BK * 4 + BN * BK * 4);
  const size_t epi = (size_t)EAGCN_ROW_TILE * (BN + 4) * 4 + (size_t)kRedFloats * 4;
  return ((pipe > epi ? pipe : epi) + 15) & ~(size_t)15;
}

// eligibility: pure function of the call's arguments and the process-wide engine switches (forward_b asks again to
// know how the statistics partials are indexed)
bool fwd_fused_ok(const eagcn_plan_t* plan, const eagcn_layer_t* l, const eagcn_work_t* w, int gemm_mode_, int agg_mode_) {
  if (!fwd_fused_mode() || gemm_mode_ == 1 || agg_mode_ != 0) return false;
  if (plan->N > EAGCN_ROW_TILE || !plan->tile_row) return false;    // a molecule could exceed one row tile
  if (l->fo_tot % 4 || l->fin % 4) return false;
  for (int v = 0; v < l->V; ++v) if ((l->fo[v] % 4) || (l->off[v] % 4)) return false;
  if (!w->wallT || !w->H || !w->Z || !w->Y) return false;
  if (!aligned16(w->H) || !aligned16(w->wallT) || !aligned16(w->Z) || !aligned16(w->Y) || !aligned16(w->ball)) return false;
  // shared memory: at least two pipeline stages beside the graph warps' staging region
  LayerDev L = to_dev(l, plan);
  const int BN = pick_chunk((int)l->fo_tot, (int)plan->t_cap);
  if (front_bytes(BN, 2) + graph_bytes(max_views_per_chunk(L, BN)) + 1024 > (size_t)kSmemBudget) return false;
  return encode_fn() != nullptr;
}

int layer_fwd_fused(const eagcn_plan_t* plan, const PlanDev& p, const LayerDev& L, const eagcn_work_t* w, cudaStream_t st) {
  const int C = L.fo_tot, K = L.fin;
  const int BN = pick_chunk(C, p.t_cap);
  const int num_kb = (K + BK - 1) / BK;
  const int nv_max = max_views_per_chunk(L, BN);
  int stages = pick_stages(BN, num_kb, BK, tc_passes());
  while (stages > 2 && front_bytes(BN, stages) + graph_bytes(nv_max) + 1024 > (size_t)kSmemBudget) --stages;
  const float* wT = (const float*)w->wallT;
  CUtensorMap mA, mB, mB2;
  if (!make_map(&mA, (const float*)w->H, p.t_cap, K, K, BM) || !make_map(&mB, wT, C, K, K, BN) ||
      !make_map(&mB2, wT + (size_t)K * C, C, K, K, BN))
    return EAGCN_E_UNSUPPORTED;
  FusedArgs g;
  g.C = C; g.K = K; g.BN = BN; g.stages = stages;
  g.tmem_cols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
  g.n_pad = (int)(w->n_pad > 0 ? w->n_pad : plan->N);
  g.want_stats = (w->training & 1) ? 1 : 0;
  g.save_z = (w->training & 16) ? 0 : 1;
  g.nv_max = nv_max;
  g.passes = tc_passes();
  g.graph_off = (uint32_t)front_bytes(BN, stages);
  g.ball = (const float*)w->ball; g.sig = (const float*)w->sig;
  g.Z = (float*)w->Z; g.Y = (float*)w->Y; g.invR = (float*)w->invR; g.partial = (float*)w->partial;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(layer_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const size_t smem = front_bytes(BN, stages) + graph_bytes(nv_max) + 1024;
  if (smem > (size_t)kSmemBudget) return EAGCN_E_UNSUPPORTED;
  const int tile_cap = 2 * (p.t_cap / EAGCN_ROW_TILE) + 1;          // any two consecutive greedy tiles hold > 128 rows
  dim3 grid((C + BN - 1) / BN, tile_cap, 1);
  EAGCN_PROF("layer_fwd_fused", st);
  EAGCN_LAUNCH(layer_fwd_fused_kernel, grid, kFusedThreads, smem, st)(mA, mB, mB2, p, L, g);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fz
}  // namespace eagcn

extern "C" int eagcn_set_fwd_fused(int on) { eagcn::fz::fwd_fused_mode() = on ? 1 : 0; return 0; }
extern "C" int eagcn_get_fwd_fused(void) { return eagcn::fz::fwd_fused_mode(); }
