// The attention matrices a GraphConv_Layer returns to its caller (reference layers.py:83,318):
//   A_weight[v,b,i,j] = sigmoid(a_v[type_v(i,j)]) * adj[b,i,j]        dense [V,B,N,N]
// Dead work for the 'sum' / 'ave' read-outs (models.py:104-108 only consumes it for 'pool'), so the
// host side materialises it only when asked.  Backward: d a_v[c] = sum_{edges of type c} dA * s(1-s).
#include "common.cuh"

namespace eagcn {

__global__ void __launch_bounds__(256) att_dense_kernel(PlanDev p, LayerDev L, float* __restrict__ A) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  if (t >= T) return;
  const int e0 = p.row_ptr[t], e1 = p.row_ptr[t + 1];
  const int pos = p.row_pos[t];
  const size_t plane = (size_t)p.B * p.N * p.N;
  for (int e = e0 + lane; e < e1; e += 32) {
    const int j = p.colpos[e] % p.N;
    for (int v = 0; v < p.V; ++v) {
      const int c = p.code[(size_t)v * p.e_cap + e];
      const float s = c < L.chan[v] ? sigmoidf_(L.att_w[v][c]) : 0.5f;
      A[(size_t)v * plane + (size_t)pos * p.N + j] = s;
    }
  }
}

// deterministic: one CTA per view, edges walked in order by 256 threads, per-thread private bins are too
// big -> two-level: each warp serialises its lanes into a shared histogram slot, warps summed in order.
__global__ void __launch_bounds__(256) att_dense_bwd_kernel(PlanDev p, LayerDev L, const float* __restrict__ dA,
                                                            float* __restrict__ datt) {
  pdl_prologue();
  __shared__ float s_hist[8][256];
  const int v = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0.0f;
  __syncthreads();
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  const size_t plane = (size_t)p.B * p.N * p.N;
  for (int t = warp; t < T; t += 8) {
    const int e0 = p.row_ptr[t], e1 = p.row_ptr[t + 1];
    const int pos = p.row_pos[t];
    for (int eb = e0; eb < e1; eb += 32) {
      const int e = eb + lane;
      float d = 0.f; int c = 0;
      if (e < e1) {
        const int j = p.colpos[e] % p.N;
        c = p.code[(size_t)v * p.e_cap + e];
        const float s = c < L.chan[v] ? sigmoidf_(L.att_w[v][c]) : 0.5f;
        d = dA[(size_t)v * plane + (size_t)pos * p.N + j] * s * (1.0f - s);
      }
      const int cnt = min(32, e1 - eb);
      for (int k = 0; k < cnt; ++k) {
        const float dk = __shfl_sync(0xffffffffu, d, k);
        const int ck = __shfl_sync(0xffffffffu, c, k);
        if (lane == 0) s_hist[warp][ck] += dk;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += s_hist[w][i];
    datt[v * EAGCN_SIG_STRIDE + i] = a;
  }
  if (threadIdx.x == 0) datt[v * EAGCN_SIG_STRIDE + 256] = 0.0f;
}

}  // namespace eagcn
using namespace eagcn;

static bool att_layer_ok(const eagcn_plan_t* plan, const eagcn_layer_t* l) {
  if (!l || l->V != plan->V) return false;
  for (int v = 0; v < l->V; ++v) if (!l->att_w[v]) return false;
  return true;
}

extern "C" int eagcn_attention_dense(const eagcn_plan_t* plan, const eagcn_layer_t* layer, void* A_out, void* stream) {
  if (!plan_ok(plan) || !att_layer_ok(plan, layer) || !A_out) return EAGCN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PlanDev p = to_dev(plan);
  LayerDev L = to_dev(layer, plan);
  cudaError_t e = cudaMemsetAsync(A_out, 0, (size_t)p.V * p.B * p.N * p.N * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  EAGCN_PROF("att_dense_kernel", st);
  EAGCN_LAUNCH(att_dense_kernel, (p.t_cap + 7) / 8, 256, 0, st)(p, L, (float*)A_out);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int eagcn_attention_dense_bwd(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const void* dA, void* datt,
                                         void* stream) {
  if (!plan_ok(plan) || !att_layer_ok(plan, layer) || !dA || !datt) return EAGCN_E_ARG;
  PlanDev p = to_dev(plan);
  LayerDev L = to_dev(layer, plan);
  EAGCN_PROF("att_dense_bwd_kernel", (cudaStream_t)stream);
  EAGCN_LAUNCH(att_dense_bwd_kernel, p.V, 256, 0, (cudaStream_t)stream)(p, L, (const float*)dA, (float*)datt);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
