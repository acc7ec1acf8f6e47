// tcgen05 (5th-gen tensor core) projection GEMM with fp32-faithful 3xTF32 error compensation.
//
//   mode NT :  C[T, N] = A[T, K] . B[N, K]^T     A, B fp32, K contiguous ("K-major" UMMA operands)
//              forward Z = H . W_all (B = W_all^T) and backward dH = Q . W_all^T
//   mode TN :  C[M, N] = A[T, M]^T . B[T, N]      A, B fp32, M/N contiguous ("MN-major" UMMA operands),
//              split-K over the T rows (one partial tile per blockIdx.z, reduced in fixed order afterwards):
//              backward dW_all = H^T . Q
//
// The dense contraction of the layer (reference layers.py:40) must match the fp32 reference to 1e-5,
// which a single TF32 (10-bit mantissa) or BF16 pass cannot.  Each fp32 operand x is split exactly into
//   x_hi = x with the 13 low mantissa bits cleared (exactly representable in TF32)
//   x_lo = x - x_hi                                  (exact in fp32; TF32 keeps its top 11 bits)
// and D = A_hi.B_hi (main accumulator) + [A_lo.B_hi + A_hi.B_lo] (separate correction accumulator, summed in
// the epilogue) is accumulated in fp32 in tensor memory: relative product error ~2^-21, i.e. fp32-class,
// at 1/3 of the TF32 tensor rate (still >4x the FFMA peak).  The two accumulators keep the number of
// (truncating) tensor-core accumulation steps into the large-magnitude sum at K/8; measured error vs a
// float64 product: a few 1e-6 of max|C| (tests/test_gpu_gemm.py).
//
// Structure (one 128 x BN output tile per CTA, BN <= 256 chosen on the host to fit N):
//   warp 0     : TMA producer   -- cp.async.bulk.tensor 2D loads of the raw fp32 tiles, 128B-swizzled
//   warps 2..5 : transform      -- split raw tiles into hi (in place) / lo (second buffer) in shared
//                                  memory; the split is element-wise so it is swizzle-agnostic
//   warp 1     : MMA issuer     -- one elected lane issues tcgen05.mma.kind::tf32 (UMMA 128 x BN x 8),
//                                  accumulators live in TMEM; tcgen05.commit releases smem stages
//   warps 2..5 : epilogue       -- tcgen05.ld 32x32b -> registers -> global (rows >= T masked)
// mbarrier pipeline: full (TMA -> transform), ready (transform -> MMA), empty (MMA -> TMA), acc (MMA -> epilogue).
// The number of valid rows T is read from device memory (plan.counts) -- no host synchronisation.
#include <cuda.h>
#include "common.cuh"

namespace eagcn {
namespace tc {

constexpr int BM = 128;            // UMMA M
constexpr int BK = 32;             // fp32 elements per k-block = 128 bytes = one SWIZZLE_128B row (TN mode: always)
// NT mode can instead run k-blocks of 16 elements (64-byte rows, SWIZZLE_64B): half the bytes per pipeline stage, so
// twice the stages fit (the 2-stage BK = 32 pipeline of the wide tiles exposes TMA latency + transform: 2 090 cycles
// per 32 k against 1 290 of tensor-pipe work).  Parity-tested, but measured slower end to end; see pick_bn.
constexpr int UK = 8;              // UMMA K for kind::tf32 (32 bytes)
constexpr int MAX_STAGES = 6;      // smem ring depth is chosen on the host (as many stages as fit in 200 KB)
constexpr int kThreads = 320;      // warp0 TMA, warp1 MMA, warps 2..9 transform + epilogue
constexpr int kXformThreads = 256;
constexpr int kSmemBudget = 224 * 1024;   // dynamic shared memory per CTA (227 KB opt-in limit minus the static part)
constexpr int kStgLd = 36;         // floats per row of an epilogue staging tile (32 + 4: conflict-free 128-bit access)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (1: unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (1024 B: 8 rows x 128 B) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
// bk == 16: SWIZZLE_64B (layout 4), 8 rows x 64 B = 512 B per group.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int bk) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((bk == 16 ? 512 : 1024) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(bk == 16 ? 4 : 2) << 61;
  return d;
}

// MN-major 32-bit operands: the only shared-memory layout UMMA accepts is SWIZZLE_128B with 32-byte atoms
// (cute UMMA::Layout_MN_SW128_32B_Atom, LayoutType::SWIZZLE_128B_BASE32B = 1; TMA: SWIZZLE_128B_ATOM_32B).
// The tile is stored as [MN/32 boxes][BK rows of k][32 floats of MN = 128 B]; 4 consecutive k rows form one
// 512-byte swizzle atom.  leading byte offset = distance between 32-element MN groups (one box = BK*128 B),
// stride byte offset = distance between groups of 4 k rows (512 B).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t box_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((box_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
// A operand from TENSOR MEMORY (lane = row of the 128-row tile, one 32-bit column per k element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
// 16 consecutive 32-bit columns of the warp's 32 TMEM lanes <- 16 registers per thread (lane = thread)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// hi/lo split of one staged operand tile by the 256 transform threads: x -> (x with 13 low mantissa bits cleared,
// x - that).  Four 16-byte vectors per thread are in flight at a time (the clock trace showed the one-at-a-time loop
// latency-bound: 1.7k cycles for a 48 KB TN stage).
__device__ __forceinline__ void split_tile(uint4* hi, uint4* lo, int n, int xt) {
  for (int i0 = xt; i0 < n; i0 += 4 * kXformThreads) {
    uint4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * kXformThreads; if (i < n) x[u] = hi[i]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kXformThreads;
      if (i < n) {
        uint4 h, l;
        h.x = x[u].x & 0xFFFFE000u; h.y = x[u].y & 0xFFFFE000u; h.z = x[u].z & 0xFFFFE000u; h.w = x[u].w & 0xFFFFE000u;
        l.x = __float_as_uint(__uint_as_float(x[u].x) - __uint_as_float(h.x));
        l.y = __float_as_uint(__uint_as_float(x[u].y) - __uint_as_float(h.y));
        l.z = __float_as_uint(__uint_as_float(x[u].z) - __uint_as_float(h.z));
        l.w = __float_as_uint(__uint_as_float(x[u].w) - __uint_as_float(h.w));
        hi[i] = h; lo[i] = l;
      }
    }
  }
}

// Tensor-core precision of the projection products.  3 (default): fp32-faithful 3xTF32 (hi/lo split, two accumulators);
// 1: ONE TF32 pass on the raw fp32 operands (the tensor core truncates the mantissa to 10 bits itself): ~1e-3 relative
// error -- outside the 1e-5 parity bar, the analogue of BASELINE.json's "bf16" configuration, reported beside the strict
// mode, never the default.  No hi/lo split, half the shared memory per stage, a third of the MMAs.  Process-wide.
inline int& tc_passes() { static int v = 3; return v; }
// Where the split activation operand (A) lives.  1 (default): the transform warps read the raw fp32 tile from shared
// memory ONCE and write its hi / lo parts straight into TENSOR MEMORY (tcgen05.st); the MMAs take A from TMEM
// (tcgen05.mma [d], [a], b-desc).  Per 128 x BN x 32 k-block that removes 32 KB of shared-memory writes (hi / lo copies)
// and 48 KB of UMMA operand reads (3 x 16 KB of A) from the pipe that bounds this kernel (shared-memory bandwidth, see
// DESIGN.md 4): 262 KB -> 182 KB at BN = 240.  TMEM budget: 2 x BN accumulator columns + 64 columns (hi | lo) per
// stage <= 512, so BN <= 192 with two stages.  0: hi / lo copies in shared memory (the round-1 form).  Process-wide.
// Bit 0: the K-major products (Z = H W, dH = Q W^T); bit 1: the MN-major split-K product dW = H^T Q.
inline int& tc_a_tmem() { static int v = 3; return v; }
constexpr int kATmemCols = 2 * BK;   // TMEM columns of one A stage: 32 (hi) + 32 (lo)

struct TcArgs {
  float* C;
  int ldc, Mcap, N, K, BN;
  const int* Mdev;                 // NT: live rows of A / C = min(*Mdev, Mcap)
  uint32_t tmem_cols;
  int mode;                        // 0 = NT (K-major), 1 = TN (MN-major, split-K)
  int stages;                      // shared-memory pipeline depth (2..MAX_STAGES)
  int b_split;                     // NT: B arrives already split (mapB = hi, mapB2 = lo): no transform of B
  const int* Kdev;                 // TN: live K rows = min(*Kdev, K)
  int kchunk;                      // TN: K rows per blockIdx.z
  int bk;                          // elements per k-block: 32 (SWIZZLE_128B) or, NT only, 16 (SWIZZLE_64B)
  long long split_stride;          // TN: elements between split partials
  long long* trace;                // diagnostic: clock64 stamps of CTA (0,0,0)'s pipeline events (nullptr: off)
  int passes;                      // 3: hi/lo-compensated 3xTF32, 1: single TF32 pass on the raw operands
  int a_tmem;                      // 1: A hi / lo in tensor memory (needs passes == 3, bk == 32, 2*BN + stages*64 <= 512)
};
constexpr int kTraceKb = 64, kTracePer = 8, kTraceStride = 8 + kTracePer * kTraceKb;

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapB2, TcArgs g) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_ready[MAX_STAGES], bar_empty[MAX_STAGES], bar_acc;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * g.BN;
  const bool tn = g.mode == 1;
  const int M = tn ? g.Mcap : min(*g.Mdev, g.Mcap);
  const int bk = g.bk;
  int k_lo = 0, num_kb = (g.K + bk - 1) / bk;
  if (tn) {
    const int Kt = min(*g.Kdev, g.K);
    k_lo = blockIdx.z * g.kchunk;
    const int k_hi = min(Kt, k_lo + g.kchunk);
    num_kb = k_hi > k_lo ? (k_hi - k_lo + bk - 1) / bk : 0;
    g.C += (long long)blockIdx.z * g.split_stride;
  }

  const bool tr = g.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (tr && threadIdx.x == 0) {
    g.trace[0] = num_kb; g.trace[1] = g.BN; g.trace[2] = g.stages; g.trace[3] = g.mode; g.trace[4] = clock64(); g.trace[7] = bk;
  }
  if (m0 >= M || num_kb == 0) {        // nothing to accumulate: keep the tile defined (zeros)
    for (int i = threadIdx.x; i < BM * g.BN; i += kThreads) {
      const int r = m0 + i / g.BN, c = n0 + i % g.BN;
      if (r < g.Mcap && c < g.N) g.C[(size_t)r * g.ldc + c] = 0.0f;
    }
    return;
  }

  // 1024-byte aligned carve-up: per stage [A raw/hi | A lo | B raw/hi | B lo]
  // 1024-byte aligned start (SWIZZLE_128B atoms) as an OFFSET into the __shared__ array: an integer round trip of the
  // pointer would lose the address space and turn every access below into a generic LD.E / ST.E instead of LDS / STS
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t bytesA = BM * bk * 4, bytesB = (uint32_t)g.BN * bk * 4;
  const bool one = g.passes == 1;                       // single TF32 pass: [A | B] per stage, no lo copies
  const bool atm = g.a_tmem != 0;                       // A hi / lo in tensor memory: [A raw | B hi | B lo] per stage
  const uint32_t stage_bytes = one ? bytesA + bytesB : (atm ? bytesA + 2 * bytesB : 2 * bytesA + 2 * bytesB);
  const uint32_t offB = (one || atm) ? bytesA : 2 * bytesA;
  auto sA_hi = [&](int s) { return base + (size_t)s * stage_bytes; };
  auto sA_lo = [&](int s) { return base + (size_t)s * stage_bytes + bytesA; };
  auto sB_hi = [&](int s) { return base + (size_t)s * stage_bytes + offB; };
  auto sB_lo = [&](int s) { return base + (size_t)s * stage_bytes + offB + bytesB; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_ready[s], kXformThreads / 32);     // one arrival per transform warp
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                     // TMEM allocation by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_smem;
  const uint32_t tmem_c = tmem_d + (uint32_t)g.BN;                 // correction accumulator: columns [BN, 2*BN)
  const uint32_t tmem_a0 = tmem_d + 2u * (uint32_t)g.BN;           // a_tmem: stage s -> columns [+64 s, +64 s + 32) hi, then lo
  // TN: 32-wide MN boxes actually inside the tensors (the others are zero-filled by the transform warps)
  const uint32_t box_bytes = bk * 128;
  const int nboxA = tn ? min(BM / 32, (g.Mcap - m0 + 31) / 32) : 0;
  const int nboxB = tn ? min(g.BN / 32, (g.N - n0 + 31) / 32) : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0, ph = 0;                               // running stage / phase: no integer division per k-block
      for (int kb = 0; kb < num_kb; ++kb) {
        if (kb >= g.stages) mbar_wait(&bar_empty[s], ph ^ 1);
        if (tr && kb < kTraceKb) g.trace[8 + kb * kTracePer + 0] = clock64();
        if (!tn) {
          mbar_expect_tx(&bar_full[s], bytesA + ((g.b_split && !one) ? 2 : 1) * bytesB);
          tma_load_2d(&mapA, &bar_full[s], sA_hi(s), kb * bk, m0);
          tma_load_2d(&mapB, &bar_full[s], sB_hi(s), kb * bk, n0);
          if (g.b_split && !one) tma_load_2d(&mapB2, &bar_full[s], sB_lo(s), kb * bk, n0);
        } else {
          const int krow = k_lo + kb * bk;
          mbar_expect_tx(&bar_full[s], (uint32_t)(nboxA + nboxB) * box_bytes);
          for (int i = 0; i < nboxA; ++i) tma_load_2d(&mapA, &bar_full[s], sA_hi(s) + i * box_bytes, m0 + 32 * i, krow);
          for (int i = 0; i < nboxB; ++i) tma_load_2d(&mapB, &bar_full[s], sB_hi(s) + i * box_bytes, n0 + 32 * i, krow);
        }
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (bit4), a=b=TF32 (2<<7, 2<<10), K-major A/B,
    // N>>3 at [17,23), M>>4 at [24,29)
    // TN: a_major (bit 15) = b_major (bit 16) = 1 (MN-major)
    // a_tmem: A comes from tensor memory, which is always "K-major" (lane = row): only b_major is set for TN
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24) | (tn ? ((atm ? 0u : (1u << 15)) | (1u << 16)) : 0u);
    // a single thread runs the whole issue loop: nothing of it needs the other 31 lanes, and re-converging the warp
    // every k-block (__syncwarp + 32 lanes polling the barrier) sat on the critical path of the tensor pipe
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (tr && kb < kTraceKb) g.trace[8 + kb * kTracePer + 5] = clock64();
        mbar_wait(&bar_ready[s], ph);
        if (tr && kb < kTraceKb) g.trace[8 + kb * kTracePer + 6] = clock64();
        tc_fence_after();
        if (tr && kb < kTraceKb) g.trace[8 + kb * kTracePer + 3] = clock64();
        const uint64_t dAh = tn ? make_desc_mn(smem_u32(sA_hi(s)), box_bytes) : make_desc(smem_u32(sA_hi(s)), bk);
        const uint64_t dAl = tn ? make_desc_mn(smem_u32(sA_lo(s)), box_bytes) : make_desc(smem_u32(sA_lo(s)), bk);
        const uint64_t dBh = tn ? make_desc_mn(smem_u32(sB_hi(s)), box_bytes) : make_desc(smem_u32(sB_hi(s)), bk);
        const uint64_t dBl = tn ? make_desc_mn(smem_u32(sB_lo(s)), box_bytes) : make_desc(smem_u32(sB_lo(s)), bk);
        const int nk = bk / UK;
#pragma unroll 4
        for (int k = 0; k < nk; ++k) {
          // NT: +32 bytes per UMMA_K inside the swizzled (128 B or 64 B) row;  TN: +1024 bytes (next group of 8 k rows)
          const uint64_t adv = tn ? (uint64_t)((k * 1024) >> 4) : (uint64_t)((k * UK * 4) >> 4);
          const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
          if (atm) {
            const uint32_t ah = tmem_a0 + (uint32_t)(s * kATmemCols + k * UK), al = ah + (uint32_t)BK;
            umma_tf32_ts(tmem_c, al, dBh + adv, idesc, acc);
            umma_tf32_ts(tmem_c, ah, dBl + adv, idesc, 1u);
            umma_tf32_ts(tmem_d, ah, dBh + adv, idesc, acc);
            continue;
          }
          if (!one) {
            umma_tf32(tmem_c, dAl + adv, dBh + adv, idesc, acc);  // correction terms -> second accumulator
            umma_tf32(tmem_c, dAh + adv, dBl + adv, idesc, 1u);
          }
          umma_tf32(tmem_d, dAh + adv, dBh + adv, idesc, acc);    // main term
        }
        if (tr && kb < kTraceKb) g.trace[8 + kb * kTracePer + 4] = clock64();
        umma_commit(&bar_empty[s]);                               // smem stage free once these MMAs retire
        if (kb == num_kb - 1) umma_commit(&bar_acc);              // accumulator complete
        if (tr && kb < kTraceKb) g.trace[8 + kb * kTracePer + 7] = clock64();
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== transform (hi/lo split), then epilogue =====================
    const int xt = threadIdx.x - 64;                              // 0..255
    int s = 0, ph = 0;                                            // running stage / phase (the kb % stages, kb / stages pair
                                                                  // was 10 % of this kernel's executed instructions)
    for (int kb = 0; kb < num_kb; ++kb, s = (s + 1 == g.stages ? 0 : s + 1), ph ^= (s == 0)) {
      mbar_wait(&bar_full[s], ph);
      if (tr && xt == 0 && kb < kTraceKb) g.trace[8 + kb * kTracePer + 1] = clock64();
      if (tn) {                                                    // boxes outside the tensors were not loaded
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        if (!atm) {
          uint4* ah = reinterpret_cast<uint4*>(sA_hi(s));
          for (int i = nboxA * (int)(box_bytes / 16) + xt; i < (int)(bytesA / 16); i += kXformThreads) ah[i] = z;
        }
        uint4* bh = reinterpret_cast<uint4*>(sB_hi(s));
        for (int i = nboxB * (int)(box_bytes / 16) + xt; i < (int)(bytesB / 16); i += kXformThreads) bh[i] = z;
      }
      if (atm) {
        // A: this thread owns row (TMEM lane) 32*quad + lane of the tile and half of the k-block's 32 k values: 16 fp32
        // from shared memory -> hi / lo in registers -> two tcgen05.st (16 columns each).  The stage's TMEM columns are
        // free: bar_full[s] fired only after the MMAs that read them retired (bar_empty -> TMA -> bar_full).
        const int quad = warp & 3, kh = (warp - 2) >> 2;           // TMEM lane quadrant of this warp, k half
        uint32_t x[16];
        if (!tn) {
          // K-major SWIZZLE_128B: row r at r*128, its 16-byte chunk c at position c ^ (r & 7)
          const int r = quad * 32 + lane;
          const uint8_t* rowp = sA_hi(s) + r * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(rowp + (((kh * 4 + c) ^ (r & 7)) << 4));
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
          }
        } else if (quad < nboxA) {
          // MN-major, 32-byte-atom 128B swizzle: box `quad` holds [32 k][32 m]; element (k, m) at
          // k*128 + (((m >> 3) ^ (k & 3)) << 5) + (m & 7)*4 -- a warp reads one permuted 128-byte line per k
          const uint8_t* boxp = sA_hi(s) + quad * box_bytes + (lane & 7) * 4;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = kh * 16 + i;
            x[i] = *reinterpret_cast<const uint32_t*>(boxp + k * 128 + ((((lane >> 3) ^ (k & 3))) << 5));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = 0u;
        }
        uint32_t h[16], l[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          h[i] = x[i] & 0xFFFFE000u;
          l[i] = __float_as_uint(__uint_as_float(x[i]) - __uint_as_float(h[i]));
        }
        const uint32_t ta = tmem_a0 + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * kATmemCols + kh * 16);
        tmem_st16(ta, h);
        tmem_st16(ta + (uint32_t)BK, l);
        if (!g.b_split)
          split_tile(reinterpret_cast<uint4*>(sB_hi(s)), reinterpret_cast<uint4*>(sB_lo(s)), (int)(bytesB / 16), xt);
        tmem_st_wait();
        tc_fence_before();
      } else if (!one) {
        split_tile(reinterpret_cast<uint4*>(sA_hi(s)), reinterpret_cast<uint4*>(sA_lo(s)), (int)(bytesA / 16), xt);
        if (!g.b_split)
          split_tile(reinterpret_cast<uint4*>(sB_hi(s)), reinterpret_cast<uint4*>(sB_lo(s)), (int)(bytesB / 16), xt);
      }
      fence_proxy_async();                                         // generic-proxy writes -> visible to UMMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ready[s]);                   // 8 arrivals per stage instead of 256
      if (tr && xt == 0 && kb < kTraceKb) g.trace[8 + kb * kTracePer + 2] = clock64();
    }
    // ---- epilogue: TMEM -> registers -> global ----
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    if (tr && xt == 0) g.trace[5] = clock64();
    const int quad = warp & 3;                                     // TMEM lane quadrant this warp may access
    const int row = m0 + quad * 32 + lane;
    const bool live = row < M;
    const bool in_cap = row < g.Mcap;
    float* crow = g.C + (size_t)row * g.ldc;
    const bool vec = ((g.ldc & 3) == 0) && ((n0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
    // 8 epilogue warps: warps w and w+4 share a TMEM lane quadrant and take alternate 32-column chunks
    const int half = (warp - 2) >> 2;
    float* stg = reinterpret_cast<float*>(base) + (warp - 2) * (32 * kStgLd);   // 4.5 KB per warp, stage memory is free now
    for (int c0 = half * 32; c0 < g.BN; c0 += 64) {
      uint32_t r[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      uint32_t q[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
            "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
            "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
            "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
          : "r"(taddr + (uint32_t)g.BN));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!one) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(q[j]));
      }
      const int cbase = n0 + c0;
      if (vec && cbase + 32 <= g.N && c0 + 32 <= g.BN) {
        // lane == row here: a direct store would touch 32 different 128-byte lines per instruction (the trace showed
        // the epilogue store-bound: 6.5k of 36k cycles).  Transpose through this warp's staging tile in the (now idle)
        // pipeline buffers so that every store instruction writes 4 rows x 128 contiguous bytes.
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = live ? __uint_as_float(r[4 * j]) : 0.f; o.y = live ? __uint_as_float(r[4 * j + 1]) : 0.f;
          o.z = live ? __uint_as_float(r[4 * j + 2]) : 0.f; o.w = live ? __uint_as_float(r[4 * j + 3]) : 0.f;
          *reinterpret_cast<float4*>(stg + lane * kStgLd + 4 * j) = o;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3);
          const float4 v = *reinterpret_cast<const float4*>(stg + rr * kStgLd + (lane & 7) * 4);
          const int grow = m0 + quad * 32 + rr;
          if (grow < g.Mcap) *reinterpret_cast<float4*>(g.C + (size_t)grow * g.ldc + cbase + (lane & 7) * 4) = v;
        }
        __syncwarp();
      } else if (in_cap) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = cbase + j;
          if (c < g.N && c0 + j < g.BN) crow[c] = live ? __uint_as_float(r[j]) : 0.f;
        }
      }
    }
    tc_fence_before();
    if (tr && xt == 0) g.trace[6] = clock64();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(g.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D fp32 row-major [rows, cols] (cols contiguous), box = [box_rows, 32 floats], 128B swizzle, OOB -> 0
// (NT: rows = M or N index, cols = K;  TN: rows = K index, cols = M or N index -- same encoding)
static bool make_map(CUtensorMap* m, const float* ptr, long long rows, long long cols, long long ld, int box_rows,
                     bool atom32 = false, int bk = BK) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)(atom32 ? BK : bk), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE,
            atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (bk == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int stage_bytes_of(int BN, int bk, int passes, bool atm) {
  if (passes == 1) return BM * bk * 4 + BN * bk * 4;
  return (atm ? 1 : 2) * BM * bk * 4 + 2 * BN * bk * 4;          // a_tmem: no hi / lo copies of A in shared memory
}
static int pick_stages(int BN, int num_kb, int bk = BK, int passes = 3, bool atm = false) {
  int st = (kSmemBudget - 1024) / stage_bytes_of(BN, bk, passes, atm);
  if (atm && st > (512 - 2 * BN) / kATmemCols) st = (512 - 2 * BN) / kATmemCols;   // TMEM columns left for A stages
  if (st > MAX_STAGES) st = MAX_STAGES;
  if (st > num_kb) st = num_kb;
  return st < 2 ? 2 : st;
}
static uint32_t tmem_cols_for(int BN, int stages, bool atm) {      // power of two >= 32
  const int need = 2 * BN + (atm ? stages * kATmemCols : 0);
  return need <= 128 ? 128u : (need <= 256 ? 256u : 512u);
}
// A-in-TMEM is available for the 3-pass mode with 128-byte k-blocks
static bool use_a_tmem(int bk, int passes, int mode_bit) { return (tc_a_tmem() & mode_bit) != 0 && passes == 3 && bk == BK; }

// N-tile width of the K-major products (multiple of 16, 64..256) from the same pipeline model as pick_tn_shape: one CTA
// per SM, so first avoid a nearly empty second wave; then a k-block costs max(MMA, (TMA round trip + transform +
// MMA) / stages) -- a tile narrow enough for a third shared-memory stage (144 columns for N = 400, with the full
// 224 KB of dynamic shared memory) hides most of the TMA round trip, which the clock trace shows exposed with two.
inline int& nt_bk_override() { static int v = 0; return v; }   // 0: model picks; 16 / 32: forced (A/B measurements)

// a_tmem variant of the model below: shared-memory BYTES per k-block bound the main loop (128 B/clk for
// TMA writes + transform reads/writes + UMMA operand reads together), the MMAs need 12 * max(75, bn/2) cycles.
static int pick_bn_atm(int N, int mtiles, int K, bool b_presplit) {
  int best = 128;
  long long best_cost = 1LL << 60;
  const int kb = (K + BK - 1) / BK;
  for (int bn = 192; bn >= 64; bn -= 16) {                     // 2*bn + 2 stages * 64 TMEM columns <= 512
    const int ntiles = (N + bn - 1) / bn;
    const long long rounds = ((long long)ntiles * mtiles + 147) / 148;
    const int stages = pick_stages(bn, kb, BK, 3, true);
    const long long bytesA = BM * BK * 4, bytesB = (long long)bn * BK * 4;
    // TMA writes (A, B hi, B lo or raw B) + A read + [B split: read + 2 writes] + 3 UMMA reads of B
    const long long bytes = 2 * bytesA + (b_presplit ? 2 : 1) * bytesB + (b_presplit ? 0 : 3 * bytesB) + 3 * bytesB;
    const long long mma = 12 * (bn / 2 > 75 ? bn / 2 : 75);
    long long per_kb = bytes / 128 + 60;
    per_kb = per_kb < mma + 60 ? mma + 60 : per_kb;
    const long long lat = (2500 + mma) / stages;               // TMA round trip + transform + MMA, hidden `stages` deep
    per_kb = per_kb < lat ? lat : per_kb;
    const long long cost = rounds * (kb * per_kb + 4000 + 25 * bn) * 100 + (long long)(ntiles * bn - N);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

static int pick_bn(int N, int mtiles, int K, bool b_presplit, int* bk_out) {
  int best = 128, best_bk = BK;
  long long best_cost = 1LL << 60;
  for (int bk = 32; bk >= 16; bk -= 16) {
    // measured (r02 A/B, bench --tc-bk): the 64-byte k-blocks run 4-6 stages deep but pay the per-k-block hand-offs
    // twice as often -- 704 k vs 723 k molecules/s -- so they are opt-in (eagcn_set_tc_bk(16)) until that cost is gone
    if (bk != (nt_bk_override() ? nt_bk_override() : BK)) continue;
    const int kb = (K + bk - 1) / bk;
    for (int bn = 256; bn >= 64; bn -= 16) {                   // UMMA M = 128 needs N % 16 == 0
      const int ntiles = (N + bn - 1) / bn;
      const long long tiles = (long long)ntiles * mtiles;
      const long long rounds = (tiles + 147) / 148;
      const int stage_bytes = 2 * BM * bk * 4 + 2 * bn * bk * 4;
      int stages = (kSmemBudget - 1024) / stage_bytes;
      stages = stages > MAX_STAGES ? MAX_STAGES : stages;
      stages = stages > kb ? kb : stages;
      stages = stages < 1 ? 1 : stages;
      const int raw_kb = (BM * bk * 4 + (b_presplit ? 2 : 1) * bn * bk * 4) / 1024;   // KB landed per k-block
      const int split_kb = (BM * bk * 4 + (b_presplit ? 0 : bn * bk * 4)) / 1024;     // KB the transform warps split
      const long long xform = 300 + 30 * split_kb;
      const long long mma = 3 * (bk / UK) * (bn / 2 > 75 ? bn / 2 : 75);
      const long long tma = 1500 + 8 * raw_kb;                         // fit: 76 KB ~ 2100 cycles
      long long per_kb = (400 + tma + xform + 100 + mma) / stages;
      per_kb = per_kb < mma + 60 ? mma + 60 : per_kb;                  // + barrier hand-off per k-block
      per_kb = per_kb < xform + 300 ? xform + 300 : per_kb;
      const long long cost = rounds * (kb * per_kb + 4000 + 25 * bn) * 100 + (long long)(ntiles * bn - N);
      if (cost < best_cost) { best_cost = cost; best = bn; best_bk = bk; }
    }
  }
  *bk_out = best_bk;
  return best;
}

bool tc_supported(const float* A, int lda, const float* B, int ldb, int K) {
  return (lda % 4 == 0) && (ldb % 4 == 0) && K >= 1 && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
         ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && encode_fn() != nullptr;
}

struct TraceState { long long* buf = nullptr; int max_launches = 0, n = 0; };
inline TraceState& trace_state() { static TraceState t; return t; }
static long long* next_trace() {
  TraceState& t = trace_state();
  if (!t.buf || t.n >= t.max_launches) return nullptr;
  return t.buf + (size_t)(t.n++) * kTraceStride;
}

// C[Mcap(T live), N] = A[Mcap, K] . B[N, K]^T   (both K-contiguous)
// B_lo != nullptr: B is pre-split (B = hi part, B_lo = lo part, same leading dimension)
int gemm_tc_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int Mcap, int N, int K,
               const int* Mdev, cudaStream_t st, const char* tag, const float* B_lo = nullptr) {
  int bk = BK;
  int BN = pick_bn(N, (Mcap + BM - 1) / BM, K, B_lo != nullptr, &bk);
  const bool atm = use_a_tmem(bk, tc_passes(), 1);
  if (atm) BN = pick_bn_atm(N, (Mcap + BM - 1) / BM, K, B_lo != nullptr);
  CUtensorMap mA, mB, mB2;
  if (!make_map(&mA, A, Mcap, K, lda, BM, false, bk) || !make_map(&mB, B, N, K, ldb, BN, false, bk)) return EAGCN_E_UNSUPPORTED;
  if (!make_map(&mB2, B_lo ? B_lo : B, N, K, ldb, BN, false, bk)) return EAGCN_E_UNSUPPORTED;
  TcArgs g{C, ldc, Mcap, N, K, BN, Mdev, 0, 0, 2, B_lo ? 1 : 0, nullptr, 0, bk, 0, next_trace(), tc_passes(), atm ? 1 : 0};
  g.stages = pick_stages(BN, (K + bk - 1) / bk, bk, g.passes, atm);
  g.tmem_cols = tmem_cols_for(BN, g.stages, atm);                  // main + correction accumulators (+ A stages)
  const size_t smem = (size_t)g.stages * stage_bytes_of(BN, bk, g.passes, atm) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((N + BN - 1) / BN, (Mcap + BM - 1) / BM, 1);
  EAGCN_PROF(tag, st);
  EAGCN_LAUNCH(gemm_tc_kernel, grid, kThreads, smem, st)(mA, mB, mB2, g);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

// split-K factor shared by both engines for dW = H^T Q  (M = fin, N = fo_tot, K rows up to Kcap)
int tn_splits(int M, int N, int Kcap) {
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  int s = (296 + tiles - 1) / tiles;
  const int kmax = (Kcap + 127) / 128;
  if (s > kmax) s = kmax;
  if (s > 32) s = 32;
  return s < 1 ? 1 : s;
}

// TN tile width (multiple of 32: whole MN boxes) and split count from a pipeline model fitted to clock traces of the
// kernel (profiles/r01_gemm_trace.json): per k-block the TMA round trip is ~2100 cycles, the hi/lo transform
// ~300 + 30/KB, twelve UMMAs cost 12 * max(75, BN/2) cycles (a narrow UMMA is not cheaper than ~75 cycles), and the
// three overlap only as far as the stage count allows.  Narrow tiles looked attractive on padding alone (N = 700 ->
// BN = 64) but run 55 k-blocks at the UMMA floor; wide tiles with more K splits are ~2.4x faster.
static void pick_tn_shape(int M, int N, int Kcap, long long ws_floats, int* bn_out, int* ns_out, bool atm) {
  const int mt = (M + BM - 1) / BM, kmax = (Kcap + 127) / 128;
  long long best = 1LL << 60;
  *bn_out = 128; *ns_out = 1;
  if (atm) {                                                    // shared-memory-bytes model, see pick_bn_atm
    for (int bn = 192; bn >= 32; bn -= 32) {
      const int tiles = mt * ((N + bn - 1) / bn);
      int ns = 148 / tiles;
      if (ns < 1) ns = 1;
      if (ns > kmax) ns = kmax;
      if (ns > 32) ns = 32;
      while (ns > 1 && (long long)ns * M * N > ws_floats) --ns;
      const int kb = (((Kcap + ns - 1) / ns) + BK - 1) / BK;
      const int stages = pick_stages(bn, kb, BK, 3, true);
      const long long bytesA = BM * BK * 4, bytesB = (long long)bn * BK * 4;
      const long long bytes = 2 * bytesA + 7 * bytesB;          // TMA A, B + A read + B split (1 + 2) + 3 UMMA reads of B
      const long long mma = 12 * (bn / 2 > 75 ? bn / 2 : 75);
      long long per_kb = bytes / 128 + 60;
      per_kb = per_kb < mma + 60 ? mma + 60 : per_kb;
      const long long lat = (2500 + 30 * (bn / 8) + mma) / stages;
      per_kb = per_kb < lat ? lat : per_kb;
      const long long waves = (tiles * (long long)ns + 147) / 148;
      const long long cost = waves * (kb * per_kb + 4000 + 25 * bn);
      if (cost < best) { best = cost; *bn_out = bn; *ns_out = ns; }
    }
    return;
  }
  for (int bn = 256; bn >= 32; bn -= 32) {
    const int tiles = mt * ((N + bn - 1) / bn);
    int ns = 148 / tiles;
    if (ns < 1) ns = 1;
    if (ns > kmax) ns = kmax;
    if (ns > 32) ns = 32;
    while (ns > 1 && (long long)ns * M * N > ws_floats) --ns;
    const int kb = (((Kcap + ns - 1) / ns) + BK - 1) / BK;
    const int stage_kb = (2 * BM * BK * 4 + 2 * bn * BK * 4) / 1024;
    int stages = (kSmemBudget / 1024 - 1) / stage_kb;
    stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    stages = stages > kb ? kb : stages;
    stages = stages < 1 ? 1 : stages;
    const long long xform = 300 + 30 * (16 + bn / 8);
    const long long mma = 12 * (bn / 2 > 75 ? bn / 2 : 75);
    long long per_kb = (2100 + xform + mma) / stages;
    per_kb = per_kb < mma ? mma : per_kb;
    per_kb = per_kb < xform + 300 ? xform + 300 : per_kb;       // the transform warps are a serial resource
    const long long waves = (tiles * (long long)ns + 147) / 148;
    const long long cost = waves * (kb * per_kb + 4000 + 25 * bn);     // + prologue and epilogue
    if (cost < best) { best = cost; *bn_out = bn; *ns_out = ns; }
  }
}

// C[M, N] = A[Kcap(T live), M]^T . B[Kcap, N]; rows >= T of A and B must be zero (they are: rows_gather,
// bn_apply and agg_bwd zero-fill the slack rows).  ws: split-K partials, reduced by the caller.
int gemm_tc_tn(const float* A, int lda, const float* B, int ldb, float* ws, long long ws_floats, int M, int N, int Kcap,
               const int* Kdev, int* nsplit_out, cudaStream_t st) {
  int BN, ns;
  const bool atm = use_a_tmem(BK, tc_passes(), 2);
  pick_tn_shape(M, N, Kcap, ws_floats, &BN, &ns, atm);
  if (ws_floats < (long long)ns * M * N) return EAGCN_E_ARG;
  CUtensorMap mA, mB;
  if (!make_map(&mA, A, Kcap, M, lda, BK, true) || !make_map(&mB, B, Kcap, N, ldb, BK, true)) return EAGCN_E_UNSUPPORTED;
  int kchunk = (Kcap + ns - 1) / ns;
  kchunk = ((kchunk + BK - 1) / BK) * BK;
  TcArgs g{ws, N, M, N, Kcap, BN, nullptr, 0, 1, 2, 0, Kdev, kchunk, BK, (long long)M * N, next_trace(), tc_passes(), atm ? 1 : 0};
  g.stages = pick_stages(BN, kchunk / BK, BK, g.passes, atm);
  g.tmem_cols = tmem_cols_for(BN, g.stages, atm);
  const size_t smem = (size_t)g.stages * stage_bytes_of(BN, BK, g.passes, atm) + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, ns);
  EAGCN_PROF("gemm_tc_tn", st);
  EAGCN_LAUNCH(gemm_tc_kernel, grid, kThreads, smem, st)(mA, mB, mB, g);
  EAGCN_LAUNCH_CHECK();
  *nsplit_out = ns;
  return 0;
}

}  // namespace tc
}  // namespace eagcn
