// Small-matrix FP32 GEMM for the dense layers of the read-out head (reference layers.py:382-388, Dense.forward:
// torch.mm(input, weight)) and their two autograd products.
//
// The head multiplies B = 256 molecule rows by 700x256 / 256x64 / 64x12 weights: at most 46 MFLOP per product, i.e.
// ~1 us of FFMA work for the whole GPU.  What matters is latency and parallelism, not tile efficiency: the library picks
// an un-split 32x32x16 kernel for these shapes (64 CTAs walking K = 700 in 44 synchronised steps: 17 us), the 128x64
// projection kernel of gemm_simt.cu leaves 2-8 CTAs.  Here:
//   * 32x32 output tiles, 64 threads (4x4 register tile each), split-K so that ~2 CTAs per SM exist;
//   * a CTA's WHOLE K range (<= 128) is staged in shared memory in one shot -- every global load of the CTA is in
//     flight at once (one memory round trip), then one barrier, then pure FFMA;
//   * split-K partials are combined inside the same launch: each CTA writes its partial tile, the LAST CTA to arrive at
//     a tile (atomic ticket) sums all partials in z order -- the result does not depend on which CTA is last, so the
//     product is bit-reproducible -- and resets the ticket for the next call.
// Operands may be stored transposed (element strides), edges of any size are handled (scalar fallback when a
// 4-element group is cut or misaligned).
#include "common.cuh"

namespace eagcn {

constexpr int kMmTile = 32;          // output tile edge
constexpr int kMmThreads = 64;       // 8 x 8 threads, 4 x 4 outputs each
constexpr int kMmKc = 128;           // K range staged per CTA
constexpr int kMmLd = kMmTile + 4;   // padded row of the staged tiles (keeps float4 alignment)

struct MmTileArgs {
  const float* A; const float* B; float* C; float* ws; int* ticket;
  long long sAm, sAk, sBk, sBn;      // element strides: A(m,k) = A[m*sAm + k*sAk], B(k,n) = B[k*sBk + n*sBn]
  int M, N, K, ns, kchunk;
};

// S[k][x] = src(x0 + x, k_lo + k) for k < kc, x < 32; zero outside the matrix.  `sx` / `sk` are the element strides of
// the x (tile) and k dimensions; exactly one of them is 1.
__device__ __forceinline__ void mm_stage(float (*S)[kMmLd], const float* __restrict__ src, long long sx, long long sk,
                                         int x0, int X, int k_lo, int kc, int tid) {
  if (sx == 1) {
    // rows of S are contiguous in memory: (k, x4) items, 8 lanes cover one 128-byte row segment
    const bool vec = ((sk & 3) == 0) && ((x0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int idx = tid; idx < kc * 8; idx += kMmThreads) {
      const int k = idx >> 3, x4 = (idx & 7) * 4;
      const float* g = src + (long long)(k_lo + k) * sk + x0 + x4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec && x0 + x4 + 3 < X) v = __ldg(reinterpret_cast<const float4*>(g));
      else {
        if (x0 + x4 + 0 < X) v.x = __ldg(g + 0);
        if (x0 + x4 + 1 < X) v.y = __ldg(g + 1);
        if (x0 + x4 + 2 < X) v.z = __ldg(g + 2);
        if (x0 + x4 + 3 < X) v.w = __ldg(g + 3);
      }
      *reinterpret_cast<float4*>(&S[k][x4]) = v;
    }
  } else {
    // k is the contiguous dimension: (x, k4) items with x fastest over the lanes -- 16-byte loads from 32 rows, stores
    // transposed without bank conflicts
    const bool vec = ((sx & 3) == 0) && ((k_lo & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int nk4 = (kc + 3) >> 2;
    for (int idx = tid; idx < nk4 * kMmTile; idx += kMmThreads) {
      const int x = idx & (kMmTile - 1), k = (idx >> 5) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (x0 + x < X) {
        const float* g = src + (long long)(x0 + x) * sx + k_lo + k;
        if (vec && k + 3 < kc) v = __ldg(reinterpret_cast<const float4*>(g));
        else {
          if (k + 0 < kc) v.x = __ldg(g + 0);
          if (k + 1 < kc) v.y = __ldg(g + 1);
          if (k + 2 < kc) v.z = __ldg(g + 2);
          if (k + 3 < kc) v.w = __ldg(g + 3);
        }
      }
      S[k][x] = v.x;
      if (k + 1 < kMmKc) S[k + 1][x] = v.y;
      if (k + 2 < kMmKc) S[k + 2][x] = v.z;
      if (k + 3 < kMmKc) S[k + 3][x] = v.w;
    }
  }
}

__global__ void __launch_bounds__(kMmThreads) mm_tile_kernel(MmTileArgs g) {
  pdl_prologue();
  __shared__ __align__(16) float As[kMmKc][kMmLd];
  __shared__ __align__(16) float Bs[kMmKc][kMmLd];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;
  const int m0 = blockIdx.y * kMmTile, n0 = blockIdx.x * kMmTile, z = blockIdx.z;
  const int k_lo = z * g.kchunk;
  const int kc = max(0, min(g.K, k_lo + g.kchunk) - k_lo);

  mm_stage(As, g.A, g.sAm, g.sAk, m0, g.M, k_lo, kc, tid);
  mm_stage(Bs, g.B, g.sBn, g.sBk, n0, g.N, k_lo, kc, tid);
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 8
  for (int k = 0; k < kc; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }

  const int gn = n0 + tx * 4;
  const bool vecC = ((g.N & 3) == 0) && gn + 3 < g.N;
  if (g.ns == 1) {
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
    if (gm >= g.M) continue;
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
  if (!s_last) return;
  __threadfence();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
    const float* p = g.ws + (long long)gm * g.N + gn;
    float* c = g.C + (long long)gm * g.N + gn;
    if (vecC) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int zz = 0; zz < g.ns; ++zz) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(p + (long long)zz * mn));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      *reinterpret_cast<float4*>(c) = s;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < g.N) {
          float s = 0.0f;
          for (int zz = 0; zz < g.ns; ++zz) s += __ldcg(p + (long long)zz * mn + j);
          c[j] = s;
        }
    }
  }
  if (tid == 0) g.ticket[tile] = 0;                 // ready for the next call on this ticket array
}

// split plan: K ranges of at most kMmKc (multiple of 4), enough of them for ~2 CTAs per SM, none shorter than 32
static void mm_tile_plan(int M, int N, int K, int* ns_out, int* kchunk_out) {
  const int tiles = ((M + kMmTile - 1) / kMmTile) * ((N + kMmTile - 1) / kMmTile);
  int ns = (2 * 148 + tiles - 1) / tiles;
  const int ns_max = K / 32 > 1 ? K / 32 : 1;
  if (ns > ns_max) ns = ns_max;
  const int ns_min = (K + kMmKc - 1) / kMmKc;
  if (ns < ns_min) ns = ns_min;
  int kchunk = (K + ns - 1) / ns;
  kchunk = (kchunk + 3) & ~3;
  if (kchunk > kMmKc) kchunk = kMmKc;
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
  EAGCN_PROF("mm_tile_kernel", stream);
  EAGCN_LAUNCH(mm_tile_kernel, dim3(gx, gy, (unsigned)ns), kMmThreads, 0, (cudaStream_t)stream)(g);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
