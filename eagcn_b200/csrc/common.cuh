// Shared device/host helpers for the eagcn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/eagcn_b200.h"

// Every kernel launch is bracketed by EAGCN_PROF(name, stream) ... EAGCN_LAUNCH_CHECK().  Besides error
// propagation this feeds two diagnostics (not thread-safe, off the product path's critical section):
// a launch counter (bench.py's gpu_launches) and an opt-in per-kernel CUDA-event profiler
// (eagcn_profile / eagcn_profile_report) used for the roofline line of bench.py.
#include <vector>
namespace eagcn {
struct ProfRec { const char* name; cudaEvent_t a, b; };
struct ProfState {
  long long launches = 0;
  int on = 0;
  bool open = false;
  ProfRec cur{};
  cudaStream_t cur_st = nullptr;
  std::vector<ProfRec> recs;
};
inline ProfState& prof() { static ProfState s; return s; }
inline void prof_begin(const char* name, cudaStream_t st) {
  ProfState& s = prof();
  ++s.launches;
  if (s.on) {
    s.cur.name = name;
    cudaEventCreate(&s.cur.a); cudaEventCreate(&s.cur.b);
    cudaEventRecord(s.cur.a, st);
    s.cur_st = st; s.open = true;
  }
}
inline void prof_end() {
  ProfState& s = prof();
  if (s.on && s.open) { cudaEventRecord(s.cur.b, s.cur_st); s.recs.push_back(s.cur); s.open = false; }
}
}  // namespace eagcn
#define EAGCN_PROF(name, st) ::eagcn::prof_begin(name, (cudaStream_t)(st))
#define EAGCN_LAUNCH_CHECK()                                  \
  do {                                                        \
    ::eagcn::prof_end();                                      \
    cudaError_t e__ = cudaPeekAtLastError();                  \
    if (e__ != cudaSuccess) return (int)e__;                  \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// A step is ~50 small dependent kernels; on the dependent chain of a CUDA graph each launch costs a few microseconds
// of scheduling latency.  Every kernel here starts with pdl_prologue(): it lets the NEXT kernel of the stream be
// scheduled as soon as all CTAs of this one have started (griddepcontrol.launch_dependents) and then blocks until the
// PREVIOUS kernel has completed and flushed (griddepcontrol.wait) -- nothing is read or written before that, so the
// semantics are those of ordinary stream order.  Kernels are launched through EAGCN_LAUNCH, which attaches
// cudaLaunchAttributeProgrammaticStreamSerialization (captured into CUDA graphs as programmatic edges).
namespace eagcn {
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline int& pdl_mode() { static int m = 1; return m; }
template <typename K>
struct PdlLaunch {
  K kern; dim3 grid, block; size_t smem; cudaStream_t st;
  template <typename... A>
  void operator()(A&&... a) const {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_mode() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, static_cast<A&&>(a)...);    // errors surface through EAGCN_LAUNCH_CHECK
  }
};
template <typename K>
inline PdlLaunch<K> make_launch(K k, dim3 g, dim3 b, size_t smem, cudaStream_t st) { return PdlLaunch<K>{k, g, b, smem, st}; }
}  // namespace eagcn
#define EAGCN_LAUNCH(kernel, grid, block, smem, st) ::eagcn::make_launch(kernel, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(st))

#define EAGCN_TINY 1e-9f            // reference layers.py:294
#define EAGCN_SIG_STRIDE 257        // sigma table row: [0..255] codes, [256] = sigmoid(self_r)
#define EAGCN_NO_EDGE 255

namespace eagcn {

constexpr int kWarp = 32;
constexpr int kStatRows = 32;       // rows per statistics tile (agg / bn kernels)
constexpr int kAggWarps = 8;        // warps per CTA of the aggregation kernels (one row at a time each)
constexpr int kAggRows = kStatRows / kAggWarps;   // rows per warp
constexpr int kAggThreads = kAggWarps * 32;
constexpr int kEltRows = 4;         // rows per thread of the float4 element-wise BatchNorm kernels

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Philox4x32-10 (Salmon et al. 2011), counter-based: the dropout keep mask is a pure function of
// (seed, offset, element index) so backward regenerates it instead of storing it.
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// keep decision for element idx (one philox call serves 4 consecutive elements).
// Counter layout: low 64 bits = element group idx/4, high 64 bits = call-site stream id + generator offset.  The offset
// advances by 2^20 per dropout call site and forward pass (eagcn_rng_fork) and stream ids are < 2^20, so the high word
// is unique per (call site, pass) and the low word never overlaps it: masks of different passes / sites are disjoint
// Philox streams for ANY tensor size (the round-1 layout added the offset to the low word, which made the masks of
// consecutive passes shifted copies of each other beyond 2^22 elements per site).
__device__ __forceinline__ bool dropout_keep(const Philox& ph, uint64_t offset, uint64_t stream, uint64_t idx, float p) {
  uint4 r = ph(idx >> 2, stream + offset);
  uint32_t x = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  // uniform in [0,1): keep iff u >= p
  return (float)(x >> 8) * (1.0f / 16777216.0f) >= p;
}

// vector form: keep flags of the 4 consecutive elements idx4 .. idx4+3 (idx4 % 4 == 0), one philox call
__device__ __forceinline__ void dropout_keep4(const Philox& ph, uint64_t offset, uint64_t stream, uint64_t idx4, float p,
                                              bool (&keep)[4]) {
  const uint4 r = ph(idx4 >> 2, stream + offset);
  keep[0] = (float)(r.x >> 8) * (1.0f / 16777216.0f) >= p;
  keep[1] = (float)(r.y >> 8) * (1.0f / 16777216.0f) >= p;
  keep[2] = (float)(r.z >> 8) * (1.0f / 16777216.0f) >= p;
  keep[3] = (float)(r.w >> 8) * (1.0f / 16777216.0f) >= p;
}

struct LayerDev {                 // by-value kernel argument: per-view parameter pointers
  int V, fin, fo_tot;
  int fo[EAGCN_MAX_VIEWS];
  int off[EAGCN_MAX_VIEWS + 1];
  int chan[EAGCN_MAX_VIEWS];
  const float* att_w[EAGCN_MAX_VIEWS];
  const float* self_r[EAGCN_MAX_VIEWS];
  const float* W[EAGCN_MAX_VIEWS];
  const float* bias[EAGCN_MAX_VIEWS];
  const float* gamma[EAGCN_MAX_VIEWS];
  const float* beta[EAGCN_MAX_VIEWS];
  float* run_mean[EAGCN_MAX_VIEWS];
  float* run_var[EAGCN_MAX_VIEWS];
  long long* nbt[EAGCN_MAX_VIEWS];
};

struct PlanDev {
  int B, N, V, t_cap, e_cap;
  int chan[EAGCN_MAX_VIEWS];
  int* counts; int* deg; int* blk; int* pos_row; int* row_pos; int* row_ptr; int* mol_ptr;
  int* col; int* colpos; int* rev;
  uint8_t* code; uint8_t* rcode;
  int* tile_row;
};

inline PlanDev to_dev(const eagcn_plan_t* p) {
  PlanDev d;
  d.B = (int)p->B; d.N = (int)p->N; d.V = (int)p->V; d.t_cap = (int)p->t_cap; d.e_cap = (int)p->e_cap;
  for (int v = 0; v < EAGCN_MAX_VIEWS; ++v) d.chan[v] = (int)p->chan[v];
  d.counts = (int*)p->counts; d.deg = (int*)p->deg; d.blk = (int*)p->blk; d.pos_row = (int*)p->pos_row;
  d.row_pos = (int*)p->row_pos; d.row_ptr = (int*)p->row_ptr; d.mol_ptr = (int*)p->mol_ptr;
  d.col = (int*)p->col; d.colpos = (int*)p->colpos; d.rev = (int*)p->rev;
  d.code = (uint8_t*)p->code; d.rcode = (uint8_t*)p->rcode;
  d.tile_row = (int*)p->tile_row;
  return d;
}

inline LayerDev to_dev(const eagcn_layer_t* l, const eagcn_plan_t* p) {
  LayerDev d;
  d.V = (int)l->V; d.fin = (int)l->fin; d.fo_tot = (int)l->fo_tot;
  for (int v = 0; v < EAGCN_MAX_VIEWS; ++v) {
    d.fo[v] = (int)l->fo[v]; d.off[v] = (int)l->off[v]; d.chan[v] = p ? (int)p->chan[v] : 0;
    d.att_w[v] = (const float*)l->att_w[v]; d.self_r[v] = (const float*)l->self_r[v];
    d.W[v] = (const float*)l->W[v]; d.bias[v] = (const float*)l->bias[v];
    d.gamma[v] = (const float*)l->gamma[v]; d.beta[v] = (const float*)l->beta[v];
    d.run_mean[v] = (float*)l->run_mean[v]; d.run_var[v] = (float*)l->run_var[v];
    d.nbt[v] = (long long*)l->nbt[v];
  }
  d.off[EAGCN_MAX_VIEWS] = (int)l->off[EAGCN_MAX_VIEWS];
  return d;
}

inline bool plan_ok(const eagcn_plan_t* p) {
  return p && p->B > 0 && p->N > 0 && p->V > 0 && p->V <= EAGCN_MAX_VIEWS && p->t_cap > 0 &&
         (p->t_cap % EAGCN_ROW_TILE) == 0 && p->e_cap > 0 && p->counts && p->deg && p->blk && p->pos_row &&
         p->row_pos && p->row_ptr && p->mol_ptr && p->col && p->colpos && p->rev && p->code && p->rcode &&
         p->tile_row && p->B * p->N < (int64_t)2147483000;
}

inline bool aligned16(const void* a) { return (reinterpret_cast<uintptr_t>(a) & 15) == 0; }

inline bool plan_ok_count(const eagcn_plan_t* p) {
  return p && p->B > 0 && p->N > 0 && p->V > 0 && p->V <= EAGCN_MAX_VIEWS && p->counts && p->deg && p->blk &&
         p->B * p->N < (int64_t)2147483000;
}

}  // namespace eagcn
