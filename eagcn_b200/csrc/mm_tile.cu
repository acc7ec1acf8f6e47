// Small-matrix FP32 GEMM for the dense layers of the read-out head (reference layers.py:382-388, Dense.forward:
// torch.mm(input, weight)) and their two autograd products.
//
// The head multiplies B = 256 molecule rows by 700x256 / 256x64 / 64x12 weights: at most 46 MFLOP per product, i.e.
// ~1 us of FFMA work for the whole GPU.  What matters is latency and parallelism, not tile efficiency: the library picks
// an un-split 32x32x16 kernel for these shapes (64 CTAs walking K = 700 in 44 synchronised steps: 17 us), the 128x64
// projection kernel of gemm_simt.cu leaves 2-8 CTAs.  Here:
//   * 32x32 output tiles; 256 threads = 4 groups of 64 (4x4 register tile each) that split every K slab between them
//     and are summed in fixed order through shared memory; long K with few tiles is also split over CTAs;
//   * K is walked in slabs of 128 staged by cp.async (zero-filling at the matrix edges) in the operands' NATURAL layouts
//     -- no registers, no transposing stores, every load of a slab (both operands) in flight at once: one memory round
//     trip per slab -- double-buffered, so the next slab loads while this one is multiplied; the 4 x 4 register tile
//     reads either layout with 16-byte shared-memory loads (4 k-steps at a time);
//   * split-K partials are combined inside the same launch: each CTA writes its partial tile, the LAST CTA to arrive at
//     a tile (atomic ticket) sums all partials in z order -- the result does not depend on which CTA is last, so the
//     product is bit-reproducible -- and resets the ticket for the next call.
// Operands may be stored transposed (element strides), edges of any size are handled (scalar fallback when a
// 4-element group is cut or misaligned).
#include "common.cuh"

namespace eagcn {

constexpr int kMmTile = 32;            // output tile edge
constexpr int kMmThreads = 256;        // 4 k-groups x (8 x 8 threads, 4 x 4 outputs each)
constexpr int kMmKGroups = kMmThreads / 64;
constexpr int kMmKc = 128;             // K slab staged per pipeline stage
constexpr int kMmLdK = kMmKc + 4;      // staged row when k is the contiguous dimension   (rows = tile index)
constexpr int kMmLdX = kMmTile + 4;    // staged row when the tile index is contiguous    (rows = k)
constexpr int kMmOpFloats = (kMmTile * kMmLdK > kMmKc * kMmLdX) ? kMmTile * kMmLdK : kMmKc * kMmLdX;
constexpr int kMmSmemBytes = 4 * kMmOpFloats * (int)sizeof(float);     // 2 operands x 2 stages

struct MmTileArgs {
  const float* A; const float* B; float* C; float* ws; int* ticket;
  long long sAm, sAk, sBk, sBn;      // element strides: A(m,k) = A[m*sAm + k*sAk], B(k,n) = B[k*sBk + n*sBn]
  int M, N, K, ns, kchunk;
};

// cp.async with zero fill: copies `bytes` (<= size) from src and zero-fills the rest of the `size`-byte destination
__device__ __forceinline__ void mm_cp16(float* dst, const float* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mm_cp4(float* dst, const float* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes)
               : "memory");
}

// k-contiguous staging swizzles the 16-byte chunks of a row with the row's 4-group index: the 8 lanes of a quarter warp
// read rows 4*tx + j -- 4 rows apart, i.e. only 2 distinct bank groups for any 16-byte-aligned row length -- and the XOR
// spreads them over all 8 (chunk index <= 31, row group <= 7: stays inside the row)
__device__ __forceinline__ int mm_swz(int x, int k) { return ((((k >> 2) ^ ((x >> 2) & 7)) << 2) | (k & 3)); }

// Stage one operand slab asynchronously in its NATURAL layout (no transposition, no registers):
//   X_CONTIG: S[k][x]  (rows of kMmLdX floats) for k < roundup4(kc) -- rows beyond kc are zero-filled;
//   else    : S[x][swz(k)] (rows of kMmLdK floats) for k < roundup4(kc) -- the k tail is zero-filled;
// x < 32; everything outside the matrix is zero-filled (src-size 0 with a valid dummy address).  sx / sk: element
// strides of the tile and k dimensions (the contiguous one is 1).
template <bool X_CONTIG>
__device__ __forceinline__ void mm_stage_async(float* S, const float* __restrict__ src, long long sx, long long sk, int x0,
                                               int X, int k_lo, int kc, int tid) {
  const bool al = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  const int kc4 = (kc + 3) & ~3;
  if (X_CONTIG) {
    if (al && sx == 1 && (sk & 3) == 0) {
      for (int idx = tid; idx < kc4 * 8; idx += kMmThreads) {
        const int k = idx >> 3, x4 = (idx & 7) * 4;
        int valid = k < kc ? X - (x0 + x4) : 0;
        valid = valid < 0 ? 0 : (valid > 4 ? 4 : valid);
        mm_cp16(S + k * kMmLdX + x4, valid ? src + (long long)(k_lo + k) * sk + x0 + x4 : src, valid * 4);
      }
    } else {
      for (int idx = tid; idx < kc4 * kMmTile; idx += kMmThreads) {
        const int k = idx >> 5, x = idx & 31;
        const bool ok = k < kc && x0 + x < X;
        mm_cp4(S + k * kMmLdX + x, ok ? src + (long long)(k_lo + k) * sk + (long long)(x0 + x) * sx : src, ok ? 4 : 0);
      }
    }
  } else {
    if (al && sk == 1 && (sx & 3) == 0 && (k_lo & 3) == 0) {
      const int nk4 = kc4 >> 2;
      for (int idx = tid; idx < nk4 * kMmTile; idx += kMmThreads) {
        const int x = idx / nk4, k = (idx - x * nk4) * 4;       // consecutive lanes walk k: coalesced rows
        int valid = x0 + x < X ? kc - k : 0;
        valid = valid < 0 ? 0 : (valid > 4 ? 4 : valid);
        mm_cp16(S + x * kMmLdK + mm_swz(x, k), valid ? src + (long long)(x0 + x) * sx + k_lo + k : src, valid * 4);
      }
    } else {
      for (int idx = tid; idx < kc4 * kMmTile; idx += kMmThreads) {
        const int x = idx / kc4, k = idx - x * kc4;
        const bool ok = k < kc && x0 + x < X;
        mm_cp4(S + x * kMmLdK + mm_swz(x, k), ok ? src + (long long)(x0 + x) * sx + (long long)(k_lo + k) * sk : src, ok ? 4 : 0);
      }
    }
  }
}

// A_MK: A staged as [m][k] (k contiguous in memory), else [k][m].  B_NK: B staged as [n][k], else [k][n].
template <bool A_MK, bool B_NK>
__global__ void __launch_bounds__(kMmThreads) mm_tile_kernel(MmTileArgs g) {
  pdl_prologue();
  extern __shared__ __align__(16) float mm_smem[];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  // a 64-thread group covers the 32 x 32 tile; the 4 groups take every 4th 4-k step of a slab (a single group would
  // leave one warp per scheduler walking 128 k-steps alone: latency-bound) and are summed in group order at the end
  const int kgrp = tid >> 6, t64 = tid & 63;
  const int tx = t64 & 7, ty = t64 >> 3;
  const int m0 = blockIdx.y * kMmTile, n0 = blockIdx.x * kMmTile, z = blockIdx.z;
  const int k_begin = z * g.kchunk, k_end = min(g.K, k_begin + g.kchunk);
  const int nslab = (k_end - k_begin + kMmKc - 1) / kMmKc;

  auto issue = [&](int sl) {
    const int k_lo = k_begin + sl * kMmKc, kc = min(kMmKc, k_end - k_lo);
    float* As = mm_smem + (sl & 1) * 2 * kMmOpFloats;
    mm_stage_async<!A_MK>(As, g.A, g.sAm, g.sAk, m0, g.M, k_lo, kc, tid);
    mm_stage_async<!B_NK>(As + kMmOpFloats, g.B, g.sBn, g.sBk, n0, g.N, k_lo, kc, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  if (nslab > 0) issue(0);
  for (int sl = 0; sl < nslab; ++sl) {
    if (sl + 1 < nslab) {
      issue(sl + 1);                                     // next slab in flight while this one is consumed
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* As = mm_smem + (sl & 1) * 2 * kMmOpFloats;
    const float* Bs = As + kMmOpFloats;
    const int kc4 = (min(kMmKc, k_end - (k_begin + sl * kMmKc)) + 3) & ~3;
#pragma unroll 2
    for (int k = kgrp * 4; k < kc4; k += 4 * kMmKGroups) {
      float av[4][4], bv[4][4];                          // av[i][kk], bv[kk][j]
      if (A_MK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = *reinterpret_cast<const float4*>(As + (ty * 4 + i) * kMmLdK + mm_swz(ty * 4, k));
          av[i][0] = t.x; av[i][1] = t.y; av[i][2] = t.z; av[i][3] = t.w;
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 t = *reinterpret_cast<const float4*>(As + (k + kk) * kMmLdX + ty * 4);
          av[0][kk] = t.x; av[1][kk] = t.y; av[2][kk] = t.z; av[3][kk] = t.w;
        }
      }
      if (B_NK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = *reinterpret_cast<const float4*>(Bs + (tx * 4 + j) * kMmLdK + mm_swz(tx * 4, k));
          bv[0][j] = t.x; bv[1][j] = t.y; bv[2][j] = t.z; bv[3][j] = t.w;
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 t = *reinterpret_cast<const float4*>(Bs + (k + kk) * kMmLdX + tx * 4);
          bv[kk][0] = t.x; bv[kk][1] = t.y; bv[kk][2] = t.z; bv[kk][3] = t.w;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)                     // k ascending in every layout: one summation order
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i][kk], bv[kk][j], acc[i][j]);
    }
    __syncthreads();                                     // this stage may be refilled by the next iteration's issue
  }
  // ---- combine the k-groups through shared memory (the staging buffers are idle now), fixed order 0 + 1 + 2 + 3 ----
  {
    float4* red = reinterpret_cast<float4*>(mm_smem);
    if (kgrp > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        red[(i * (kMmKGroups - 1) + (kgrp - 1)) * 64 + t64] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    if (kgrp == 0) {
#pragma unroll
      for (int q = 0; q < kMmKGroups - 1; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = red[(i * (kMmKGroups - 1) + q) * 64 + t64];
          acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
        }
    }
  }
  const bool owner = kgrp == 0;                          // holds the CTA's tile from here on

  const int gn = n0 + tx * 4;
  const bool vecC = ((g.N & 3) == 0) && gn + 3 < g.N;
  if (g.ns == 1) {
    if (!owner) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + ty * 4 + i;
      if (gm >= g.M) continue;
      float* c = g.C + (long long)gm * g.N + gn;
      if (vecC) *reinterpret_cast<float4*>(c) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      else
#pragma unroll
        for (int j = 0; j < 4; ++j) if (gn + j < g.N) c[j] = acc[i][j];
    }
    return;
  }
  // ---- split-K: publish the partial tile, take a ticket; the last arrival sums all partials in z order ----
  const long long mn = (long long)g.M * g.N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M || !owner) continue;
    float* c = g.ws + (long long)z * mn + (long long)gm * g.N + gn;
    if (vecC) __stcg(reinterpret_cast<float4*>(c), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    else
#pragma unroll
      for (int j = 0; j < 4; ++j) if (gn + j < g.N) __stcg(c + j, acc[i][j]);
  }
  __threadfence();
  __syncthreads();
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (tid == 0) s_last = (atomicAdd(g.ticket + tile, 1) == g.ns - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last || !owner) return;
  __threadfence();
  if (vecC && m0 + ty * 4 + 3 < g.M) {
    // the 4 rows of a thread together, z unrolled: 16 independent loads in flight
    const float* p = g.ws + (long long)(m0 + ty * 4) * g.N + gn;
    float4 s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int zz = 0; zz < g.ns; ++zz) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(p + (long long)zz * mn + (long long)i * g.N));
        s[i].x += v.x; s[i].y += v.y; s[i].z += v.z; s[i].w += v.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(g.C + (long long)(m0 + ty * 4 + i) * g.N + gn) = s[i];
  } else {
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + ty * 4 + i;
      if (gm >= g.M) continue;
      for (int j = 0; j < 4; ++j)
        if (gn + j < g.N) {
          float s = 0.0f;
          for (int zz = 0; zz < g.ns; ++zz) s += __ldcg(g.ws + (long long)zz * mn + (long long)gm * g.N + gn + j);
          g.C[(long long)gm * g.N + gn + j] = s;
        }
    }
  }
  if (tid == 0) g.ticket[tile] = 0;                 // ready for the next call on this ticket array
}

// split plan.  A CTA walks its K range in slabs of kMmKc; K <= 2 slabs or enough tiles to fill the GPU: no split (no
// ticket, no partials).  Longer K with few tiles (den1 forward: 64 tiles x K = 700 / 1400): one slab per CTA.
static void mm_tile_plan(int M, int N, int K, int* ns_out, int* kchunk_out) {
  const int tiles = ((M + kMmTile - 1) / kMmTile) * ((N + kMmTile - 1) / kMmTile);
  int ns = 1;
  if (K > 2 * kMmKc && tiles < 2 * 148) {
    ns = (K + kMmKc - 1) / kMmKc;
    const int cap = (4 * 148 + tiles - 1) / tiles;       // at most ~4 CTAs per SM
    if (ns > cap) ns = cap;
  }
  int kchunk = (K + ns - 1) / ns;
  kchunk = (kchunk + 3) & ~3;
  ns = (K + kchunk - 1) / kchunk;
  *ns_out = ns; *kchunk_out = kchunk;
}

}  // namespace eagcn

extern "C" int64_t eagcn_mm_tile_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0 || M > (1 << 24) || N > (1 << 24) || K > (1 << 24)) return 0;
  int ns, kc;
  eagcn::mm_tile_plan((int)M, (int)N, (int)K, &ns, &kc);
  return ns > 1 ? (int64_t)ns * M * N * (int64_t)sizeof(float) : 0;
}
extern "C" int64_t eagcn_mm_tile_tickets(int64_t M, int64_t N) {
  if (M <= 0 || N <= 0) return 0;
  return ((M + eagcn::kMmTile - 1) / eagcn::kMmTile) * ((N + eagcn::kMmTile - 1) / eagcn::kMmTile);
}
extern "C" int eagcn_mm_tile(const void* A, int64_t lda, int transA, const void* B, int64_t ldb, int transB, void* C,
                             int64_t M, int64_t N, int64_t K, void* ws, int64_t ws_bytes, void* tickets, void* stream) {
  using namespace eagcn;
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || M > (1 << 24) || N > (1 << 24) || K > (1 << 24)) return EAGCN_E_ARG;
  if (lda < (transA ? M : K) || ldb < (transB ? K : N)) return EAGCN_E_ARG;
  int ns, kchunk;
  mm_tile_plan((int)M, (int)N, (int)K, &ns, &kchunk);
  if (ns > 1 && (!ws || !tickets || ws_bytes < (int64_t)ns * M * N * (int64_t)sizeof(float))) return EAGCN_E_ARG;
  const unsigned gx = (unsigned)((N + kMmTile - 1) / kMmTile), gy = (unsigned)((M + kMmTile - 1) / kMmTile);
  if (gy > 65535u || ns > 65535) return EAGCN_E_UNSUPPORTED;
  MmTileArgs g{(const float*)A, (const float*)B, (float*)C, (float*)ws, (int*)tickets,
               transA ? 1 : lda, transA ? lda : 1, transB ? 1 : ldb, transB ? ldb : 1,
               (int)M, (int)N, (int)K, ns, kchunk};
  void (*kern)(MmTileArgs) = transA ? (transB ? mm_tile_kernel<false, true> : mm_tile_kernel<false, false>)
                                    : (transB ? mm_tile_kernel<true, true> : mm_tile_kernel<true, false>);
  static bool attr_set = false;
  if (!attr_set) {
    void (*all[4])(MmTileArgs) = {mm_tile_kernel<false, false>, mm_tile_kernel<false, true>, mm_tile_kernel<true, false>,
                                  mm_tile_kernel<true, true>};
    for (auto k : all) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmSmemBytes);
      if (e != cudaSuccess) return (int)e;
    }
    attr_set = true;
  }
  // four template instantiations = four kernels (as the profilers list them): tag them apart
  EAGCN_PROF(transA ? (transB ? "mm_tile_tt" : "mm_tile_tn") : (transB ? "mm_tile_nt" : "mm_tile_nn"), stream);
  EAGCN_LAUNCH(kern, dim3(gx, gy, (unsigned)ns), kMmThreads, kMmSmemBytes, (cudaStream_t)stream)(g);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
