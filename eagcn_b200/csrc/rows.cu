// Dense [B,N,F] <-> packed active rows [T,F], and the sum read-out over atoms.
//
// rows_scatter is the reference's "cat(...) * mask3" (layers.py:313): padded rows and atoms without
// bonds are written as exact zeros.  rows_gather is the inverse at the layer input.  The read-out
// is models.py:108 (torch.sum(x2, 1)) evaluated on packed rows: rows of molecule b are contiguous
// [mol_ptr[b], mol_ptr[b+1]) and are summed in ascending order (deterministic).
#include "common.cuh"

namespace eagcn {

// one warp per row, lanes stride the feature dimension (float4 when F % 4 == 0)
__global__ void __launch_bounds__(256) rows_gather_kernel(PlanDev p, const float* __restrict__ dense,
                                                          float* __restrict__ packed, int F) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  if (t >= p.t_cap) return;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  float* dst = packed + (size_t)t * F;
  if (t >= T) {                                   // keep the slack rows of the last tile defined
    for (int c = lane; c < F; c += 32) dst[c] = 0.0f;
    return;
  }
  const float* src = dense + (size_t)p.row_pos[t] * F;
  if ((F & 3) == 0) {
    for (int c = lane; c < F / 4; c += 32) reinterpret_cast<float4*>(dst)[c] = __ldg(reinterpret_cast<const float4*>(src) + c);
  } else {
    for (int c = lane; c < F; c += 32) dst[c] = __ldg(src + c);
  }
}

__global__ void __launch_bounds__(256) rows_scatter_kernel(PlanDev p, const float* __restrict__ packed,
                                                           float* __restrict__ dense, int F) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pos = blockIdx.x * 8 + warp;
  if (pos >= p.B * p.N) return;
  const int t = p.pos_row[pos];
  float* dst = dense + (size_t)pos * F;
  if ((F & 3) == 0) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = reinterpret_cast<const float4*>(packed + (size_t)max(t, 0) * F);
    for (int c = lane; c < F / 4; c += 32) reinterpret_cast<float4*>(dst)[c] = t >= 0 ? __ldg(src + c) : z;
  } else {
    const float* src = packed + (size_t)max(t, 0) * F;
    for (int c = lane; c < F; c += 32) dst[c] = t >= 0 ? __ldg(src + c) : 0.0f;
  }
}

// grid (B, ceil(F/64)); 64 channels x 8 row groups: a group sums a contiguous eighth of the molecule's rows (the
// largest molecule would otherwise be a 132-deep serial chain), the groups are combined in fixed order
__global__ void __launch_bounds__(512) readout_sum_kernel(PlanDev p, const float* __restrict__ packed,
                                                          float* __restrict__ out, int F) {
  pdl_prologue();
  __shared__ float s[8][64];
  const int b = blockIdx.x;
  const int cx = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cx;
  const int t0 = p.mol_ptr[b], t1 = p.mol_ptr[b + 1];
  const int per = (t1 - t0 + 7) >> 3;
  const int a0 = t0 + g * per, a1 = min(t1, a0 + per);
  float acc = 0.0f;
  if (c < F) {
#pragma unroll 4
    for (int t = a0; t < a1; ++t) acc += __ldg(packed + (size_t)t * F + c);
  }
  s[g][cx] = acc;
  __syncthreads();
  if (g == 0 && c < F) {
    float r = s[0][cx];
#pragma unroll
    for (int k = 1; k < 8; ++k) r += s[k][cx];
    out[(size_t)b * F + c] = r;
  }
}

// float4 form (F % 4 == 0): grid (B, ceil(F/256)); 64 float4 channel lanes x 4 row groups.  The 512-thread scalar kernel
// above needs B * F/64 = 2 816 CTAs (five waves of tiny CTAs: 8.9 us for 13.6 MB); this one 768 CTAs with 16-byte loads.
__global__ void __launch_bounds__(256) readout_sum_vec_kernel(PlanDev p, const float* __restrict__ packed,
                                                              float* __restrict__ out, int F) {
  pdl_prologue();
  __shared__ float4 s[4][64];
  const int b = blockIdx.x;
  const int cx = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int nc4 = F >> 2, c4 = blockIdx.y * 64 + cx;
  const int t0 = p.mol_ptr[b], t1 = p.mol_ptr[b + 1];
  const int per = (t1 - t0 + 3) >> 2;
  const int a0 = t0 + g * per, a1 = min(t1, a0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < nc4) {
    const float4* src = reinterpret_cast<const float4*>(packed) + c4;
#pragma unroll 4
    for (int t = a0; t < a1; ++t) {
      const float4 v = __ldg(src + (size_t)t * nc4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  s[g][cx] = acc;
  __syncthreads();
  if (g == 0 && c4 < nc4) {
    float4 r = s[0][cx];
#pragma unroll
    for (int k = 1; k < 4; ++k) { const float4 v = s[k][cx]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
    reinterpret_cast<float4*>(out)[(size_t)b * nc4 + c4] = r;
  }
}

__global__ void __launch_bounds__(256) readout_sum_bwd_kernel(PlanDev p, const float* __restrict__ dout,
                                                              float* __restrict__ dpacked, int F) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  if (t >= p.t_cap) return;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  float* dst = dpacked + (size_t)t * F;
  if (t >= T) { for (int c = lane; c < F; c += 32) dst[c] = 0.0f; return; }
  const int b = p.row_pos[t] / p.N;
  const float* src = dout + (size_t)b * F;
  for (int c = lane; c < F; c += 32) dst[c] = __ldg(src + c);
}

// snapshot the philox (seed, offset) for one dropout call site and advance the live state: one launch instead of a
// clone plus an add (both graph-replay safe, but two kernels per layer call)
__global__ void rng_fork_kernel(unsigned long long* __restrict__ state, unsigned long long* __restrict__ snap,
                                unsigned long long inc, int n) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long seed = state[0], off = state[1];
    for (int i = 0; i < n; ++i) { snap[2 * i] = seed; snap[2 * i + 1] = off + (unsigned long long)i * inc; }
    state[1] = off + (unsigned long long)n * inc;
  }
}

}  // namespace eagcn
using namespace eagcn;

extern "C" int eagcn_rng_fork(void* state, void* snapshot, int64_t increment, void* stream) {
  if (!state || !snapshot || increment <= 0) return EAGCN_E_ARG;
  EAGCN_PROF("rng_fork_kernel", stream);
  EAGCN_LAUNCH(rng_fork_kernel, 1, 32, 0, stream)((unsigned long long*)state, (unsigned long long*)snapshot,
                                                  (unsigned long long)increment, 1);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int eagcn_rng_fork_n(void* state, void* snapshots, int64_t n, int64_t increment, void* stream) {
  if (!state || !snapshots || increment <= 0 || n <= 0 || n > 4096) return EAGCN_E_ARG;
  EAGCN_PROF("rng_fork_kernel", stream);
  EAGCN_LAUNCH(rng_fork_kernel, 1, 32, 0, stream)((unsigned long long*)state, (unsigned long long*)snapshots,
                                                  (unsigned long long)increment, (int)n);
  EAGCN_LAUNCH_CHECK();
  return 0;
}


extern "C" int eagcn_rows_gather(const eagcn_plan_t* plan, const void* dense, void* packed, int64_t F, void* stream) {
  if (!plan_ok(plan) || !dense || !packed || F <= 0) return EAGCN_E_ARG;
  PlanDev p = to_dev(plan);
  EAGCN_PROF("rows_gather_kernel", (cudaStream_t)stream);
  EAGCN_LAUNCH(rows_gather_kernel, (p.t_cap + 7) / 8, 256, 0, (cudaStream_t)stream)(p, (const float*)dense, (float*)packed, (int)F);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
extern "C" int eagcn_rows_scatter(const eagcn_plan_t* plan, const void* packed, void* dense, int64_t F, void* stream) {
  if (!plan_ok(plan) || !dense || !packed || F <= 0) return EAGCN_E_ARG;
  PlanDev p = to_dev(plan);
  EAGCN_PROF("rows_scatter_kernel", (cudaStream_t)stream);
  EAGCN_LAUNCH(rows_scatter_kernel, (p.B * p.N + 7) / 8, 256, 0, (cudaStream_t)stream)(p, (const float*)packed, (float*)dense, (int)F);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
extern "C" int eagcn_readout_sum(const eagcn_plan_t* plan, const void* packed, void* out, int64_t F, void* stream) {
  if (!plan_ok(plan) || !out || !packed || F <= 0) return EAGCN_E_ARG;
  PlanDev p = to_dev(plan);
  EAGCN_PROF("readout_sum_kernel", (cudaStream_t)stream);
  if ((F & 3) == 0 && aligned16(packed) && aligned16(out)) {
    dim3 grid(p.B, (unsigned)((F / 4 + 63) / 64));
    EAGCN_LAUNCH(readout_sum_vec_kernel, grid, 256, 0, (cudaStream_t)stream)(p, (const float*)packed, (float*)out, (int)F);
  } else {
    dim3 grid(p.B, (unsigned)((F + 63) / 64));
    EAGCN_LAUNCH(readout_sum_kernel, grid, 512, 0, (cudaStream_t)stream)(p, (const float*)packed, (float*)out, (int)F);
  }
  EAGCN_LAUNCH_CHECK();
  return 0;
}
extern "C" int eagcn_readout_sum_bwd(const eagcn_plan_t* plan, const void* dout, void* dpacked, int64_t F, void* stream) {
  if (!plan_ok(plan) || !dout || !dpacked || F <= 0) return EAGCN_E_ARG;
  PlanDev p = to_dev(plan);
  EAGCN_PROF("readout_sum_bwd_kernel", (cudaStream_t)stream);
  EAGCN_LAUNCH(readout_sum_bwd_kernel, (p.t_cap + 7) / 8, 256, 0, (cudaStream_t)stream)(p, (const float*)dout, (float*)dpacked, (int)F);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
