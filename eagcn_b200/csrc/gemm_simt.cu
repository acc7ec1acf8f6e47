// FP32 FFMA GEMM over packed atom rows (strict-fp32 projection path).
//
// The dense contraction of the layer (reference layers.py:40, torch.mm(support, W)) and its two
// backward products.  Row counts come from DEVICE memory (plan.counts[T]) so the whole step stays
// free of host synchronisation and is CUDA-graph capturable:
//   NN : C[T,N]  = A[T,K] . B[K,N]              (forward:  Z  = H . W_all)
//   NT : C[T,N]  = A[T,K] . B[N,K]^T            (backward: dH = Q . W_all^T)
//   TN : C[M,N]  = A[T,M]^T . B[T,N]            (backward: dW = H^T . Q, split-K over T, fixed-order
//                                                reduction -> deterministic)
// 128x64x16 tiles, 256 threads, 8x4 register tile per thread, padded shared-memory operand tiles.
// This is the FFMA (CUDA-core) path: bit-faithful fp32 products; the tcgen05 path (gemm_tc.cu)
// replaces it for the tensor-core modes.
#include "common.cuh"

namespace eagcn {

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;
constexpr int GAS = GBM + 4;   // padded strides (floats); keep 16-byte alignment of float4 reads
constexpr int GBS = GBN + 4;

struct GemmArgs {
  const float* A; const float* B; float* C;
  long long sAm, sAk, sBk, sBn;   // element strides
  int ldc;
  int Mcap, N, Kcap;              // static extents (capacity)
  const int* Mdev;                // if non-null: valid rows of A/C = min(*Mdev, Mcap) (NN / NT)
  const int* Kdev;                // if non-null: valid K = min(*Kdev, Kcap)            (TN)
  int kchunk;                     // K range per blockIdx.z (split-K); C then points at partials
  long long split_stride;         // elements between split partials
};

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GTHREADS) gemm_simt_kernel(GemmArgs g) {
  pdl_prologue();
  __shared__ __align__(16) float As[GBK][GAS];
  __shared__ __align__(16) float Bs[GBK][GBS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;        // 16 x 16 threads: ty -> 8 rows, tx -> 4 cols
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int M = g.Mdev ? min(*g.Mdev, g.Mcap) : g.Mcap;
  const int K = g.Kdev ? min(*g.Kdev, g.Kcap) : g.Kcap;
  const int k_lo = blockIdx.z * g.kchunk;
  const int k_hi = min(K, k_lo + g.kchunk);
  float* C = g.C + (long long)blockIdx.z * g.split_stride;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  if (m0 < M) {
    for (int k0 = k_lo; k0 < k_hi; k0 += GBK) {
      // ---- stage A tile (128 x 16) ----
#pragma unroll
      for (int i = 0; i < (GBM * GBK) / GTHREADS; ++i) {
        const int idx = tid + i * GTHREADS;
        int m, k;
        if (A_KC) { k = idx % GBK; m = idx / GBK; } else { m = idx % GBM; k = idx / GBM; }
        const int gm = m0 + m, gk = k0 + k;
        float v = 0.0f;
        if (gm < M && gk < k_hi) v = __ldg(g.A + gm * g.sAm + gk * g.sAk);
        As[k][m] = v;
      }
      // ---- stage B tile (16 x 64) ----
#pragma unroll
      for (int i = 0; i < (GBN * GBK) / GTHREADS; ++i) {
        const int idx = tid + i * GTHREADS;
        int n, k;
        if (B_KC) { k = idx % GBK; n = idx / GBK; } else { n = idx % GBN; k = idx / GBN; }
        const int gn = n0 + n, gk = k0 + k;
        float v = 0.0f;
        if (gn < g.N && gk < k_hi) v = __ldg(g.B + gk * g.sBk + gn * g.sBn);
        Bs[k][n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < GBK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- epilogue: rows in [M, Mcap) are written as zeros so slack rows stay defined ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= g.Mcap) continue;
    const bool live = gm < M;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < g.N) C[(long long)gm * g.ldc + gn] = live ? acc[i][j] : 0.0f;
    }
  }
}

// C[i] = sum_z part[z][i], z ascending (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ C,
                                                            long long n, int nsplit) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int z = 0; z < nsplit; ++z) s += part[(long long)z * n + i];
  C[i] = s;
}

int gemm_nn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int Mcap, int N, int K,
            const int* Mdev, cudaStream_t st) {
  GemmArgs g{A, B, C, lda, 1, ldb, 1, ldc, Mcap, N, K, Mdev, nullptr, K + GBK, 0};
  dim3 grid((N + GBN - 1) / GBN, (Mcap + GBM - 1) / GBM, 1);
  EAGCN_PROF("gemm_simt_nn", st);
  EAGCN_LAUNCH((gemm_simt_kernel<true, false>), grid, GTHREADS, 0, st)(g);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

int gemm_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int Mcap, int N, int K,
            const int* Mdev, cudaStream_t st) {
  // B given as [N, K] row-major (K contiguous)
  GemmArgs g{A, B, C, lda, 1, 1, ldb, ldc, Mcap, N, K, Mdev, nullptr, K + GBK, 0};
  dim3 grid((N + GBN - 1) / GBN, (Mcap + GBM - 1) / GBM, 1);
  EAGCN_PROF("gemm_simt_nt", st);
  EAGCN_LAUNCH((gemm_simt_kernel<true, true>), grid, GTHREADS, 0, st)(g);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

namespace tc { int tn_splits(int M, int N, int Kcap); }

long long gemm_tn_workspace_floats(int M, int N, int Kcap) { return (long long)tc::tn_splits(M, N, Kcap) * M * N; }

int splitk_reduce(const float* ws, float* C, long long n, int ns, cudaStream_t st) {
  EAGCN_PROF("splitk_reduce_kernel", st);
  EAGCN_LAUNCH(splitk_reduce_kernel, (unsigned)((n + 255) / 256), 256, 0, st)(ws, C, n, ns);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

// nsplit_out != nullptr: leave the split-K partials in ws (the caller reduces them) and report their number
int gemm_tn(const float* A, int lda, const float* B, int ldb, float* C, int M, int N, int Kcap, const int* Kdev,
            float* ws, long long ws_floats, cudaStream_t st, int* nsplit_out = nullptr) {
  // A given as [K, M] row-major, B as [K, N] row-major; C [M, N] (ldc = N)
  const int ns = tc::tn_splits(M, N, Kcap);
  if (ws_floats < (long long)ns * M * N) return EAGCN_E_ARG;
  int kchunk = (Kcap + ns - 1) / ns;
  kchunk = ((kchunk + GBK - 1) / GBK) * GBK;
  GemmArgs g{A, B, ws, 1, lda, ldb, 1, N, M, N, Kcap, nullptr, Kdev, kchunk, (long long)M * N};
  dim3 grid((N + GBN - 1) / GBN, (M + GBM - 1) / GBM, ns);
  EAGCN_PROF("gemm_simt_tn", st);
  EAGCN_LAUNCH((gemm_simt_kernel<false, false>), grid, GTHREADS, 0, st)(g);
  EAGCN_LAUNCH_CHECK();
  if (nsplit_out) { *nsplit_out = ns; return 0; }
  return splitk_reduce(ws, C, (long long)M * N, ns, st);
}


}  // namespace eagcn
