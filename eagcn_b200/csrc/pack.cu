// Graph-plan construction: padded dense batch -> packed active rows + CSR edges + uint8 edge codes.
//
// Replaces, once per batch and for all layers (forward and backward), what the reference rebuilds
// inside every GraphConv_Layer.forward call: mask_blank / identity (layers.py:294-304) and the
// 1x1-conv attention-score input (layers.py:82), whose one-hot planes [B,C_v,N,N] collapse to one
// uint8 code per directed edge and view (the conv over a one-hot vector is a table lookup).
// The one-hot planes are only *gathered at bonded pairs* (adj != 0); the off-graph zeros of the
// 4*(Kb+10)*N^2 bytes per molecule are never read -- they are multiplied by adj == 0 in the
// reference (layers.py:83) and cannot influence the result.
//
// Four small kernels, no host synchronisation, deterministic output order (rows ascending by
// flat position, neighbours ascending by column):
//   count : one warp per padded row -> degree; per-CTA (active rows, edges) totals
//   scan  : single CTA exclusive scan of the per-CTA totals -> T, E
//   fill  : row numbering, CSR offsets, neighbour positions, per-view codes (validated one-hot)
//   link  : neighbour row ids, reverse-edge index (validates symmetry), reverse codes
#include "common.cuh"

namespace eagcn {

constexpr int kPackRows = 32;      // padded rows per CTA
constexpr int kPackThreads = 256;  // 8 warps x 4 rows
constexpr int kRowsPerWarp = kPackRows / 8;

struct RelPtrs { const float* p[EAGCN_MAX_VIEWS]; };

// adjacency value at (b,i,j): from the dense fp32 adjacency or from the view-0 uint8 code plane
template <bool kCodes>
__device__ __forceinline__ float adj_at(const float* __restrict__ adj, const uint8_t* __restrict__ codes, const PlanDev& p,
                                        size_t row_base_adj, size_t row_base_code, int j) {
  if (kCodes) return __ldg(codes + row_base_code + j) != EAGCN_NO_EDGE ? 1.0f : 0.0f;
  return __ldg(adj + row_base_adj + j);
}

template <bool kCodes>
__global__ void __launch_bounds__(kPackThreads) pack_count_kernel(PlanDev p, const float* __restrict__ adj,
                                                                  const uint8_t* __restrict__ codes) {
  pdl_prologue();
  __shared__ int s_deg[kPackRows];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.B * p.N;
  bool bad = false;
  int cnt[kRowsPerWarp];
  size_t ba[kRowsPerWarp], bc[kRowsPerWarp];
  bool ok[kRowsPerWarp];
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    const int row = blockIdx.x * kPackRows + warp * kRowsPerWarp + r;
    ok[r] = row < P;
    const int b = ok[r] ? row / p.N : 0, i = ok[r] ? row - b * p.N : 0;
    ba[r] = ((size_t)b * p.N + i) * p.N;
    bc[r] = (((size_t)b * p.V) * p.N + i) * p.N;
    cnt[r] = 0;
  }
#pragma unroll 4
  for (int j0 = 0; j0 < p.N; j0 += 32) {
    const int j = j0 + lane;
    float a[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r)                       // independent loads first (latency overlap)
      a[r] = (ok[r] && j < p.N) ? adj_at<kCodes>(adj, codes, p, ba[r], bc[r], j) : 0.0f;
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      if (a[r] != 0.0f && a[r] != 1.0f) bad = true;
      cnt[r] += __popc(__ballot_sync(0xffffffffu, a[r] != 0.0f));
    }
  }
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    const int row = blockIdx.x * kPackRows + warp * kRowsPerWarp + r;
    if (lane == 0) { s_deg[warp * kRowsPerWarp + r] = cnt[r]; if (ok[r]) p.deg[row] = cnt[r]; }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&p.counts[EAGCN_CNT_STATUS], EAGCN_ST_ADJ_NOT_01);
  __syncthreads();
  if (threadIdx.x == 0) {
    int na = 0, ne = 0;
    for (int r = 0; r < kPackRows; ++r) { na += s_deg[r] > 0; ne += s_deg[r]; }
    p.blk[2 * blockIdx.x] = na;
    p.blk[2 * blockIdx.x + 1] = ne;
  }
}

// single CTA: exclusive scan (in place) of blk[2*i], blk[2*i+1]; totals -> counts
__global__ void __launch_bounds__(1024) pack_scan_kernel(PlanDev p, int nblk) {
  pdl_prologue();
  __shared__ int s_a[1024], s_e[1024];
  const int tid = threadIdx.x;
  const int per = (nblk + 1023) / 1024;
  const int lo = tid * per, hi = min(nblk, lo + per);
  int sa = 0, se = 0;
  for (int i = lo; i < hi; ++i) { sa += p.blk[2 * i]; se += p.blk[2 * i + 1]; }
  s_a[tid] = sa; s_e[tid] = se;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {           // Hillis-Steele inclusive scan
    int va = 0, ve = 0;
    if (tid >= o) { va = s_a[tid - o]; ve = s_e[tid - o]; }
    __syncthreads();
    s_a[tid] += va; s_e[tid] += ve;
    __syncthreads();
  }
  int ra = s_a[tid] - sa, re = s_e[tid] - se;     // exclusive prefix of this thread's chunk
  for (int i = lo; i < hi; ++i) {
    int a = p.blk[2 * i], e = p.blk[2 * i + 1];
    p.blk[2 * i] = ra; p.blk[2 * i + 1] = re;
    ra += a; re += e;
  }
  if (tid == 1023) {
    const int T = s_a[1023], E = s_e[1023];
    p.counts[EAGCN_CNT_T] = T;
    p.counts[EAGCN_CNT_E] = E;
    int st = 0;
    if (T > p.t_cap) st |= EAGCN_ST_ROW_CAP;
    if (E > p.e_cap) st |= EAGCN_ST_EDGE_CAP;
    if (st) atomicOr(&p.counts[EAGCN_CNT_STATUS], st);
  }
}

// plane_lanes != 0: plane-per-lane gather (fewest instructions; one 32-byte sector per (pair, plane) -- right for planes
// in device memory, where the kernel is bound by the DRAM random-access rate: ~0.5 M 64-byte atoms per Tox21 batch);
// 0: column-coalesced gather (neighbouring bonded columns of one plane share a load: fewest memory REQUESTS -- right for
// planes left in pinned host memory, where the PCIe read-request rate is the bound).
template <bool kCodes>
__global__ void __launch_bounds__(kPackThreads) pack_fill_kernel(PlanDev p, const float* __restrict__ adj,
                                                                 const uint8_t* __restrict__ codes, RelPtrs rel,
                                                                 int plane_lanes) {
  pdl_prologue();
  __shared__ int s_deg[kPackRows], s_t[kPackRows], s_e[kPackRows];
  __shared__ unsigned s_mask[8][kRowsPerWarp][8];
  __shared__ int s_run[8][kRowsPerWarp];             // edges of a row already emitted (rows wider than 256 columns)
  __shared__ int s_hit[8][EAGCN_MAX_VIEWS][32];      // per (view, edge of the chunk): (#nonzero planes << 16) + plane index
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.B * p.N;
  const int row0 = blockIdx.x * kPackRows;
  if (threadIdx.x < kPackRows) {
    const int row = row0 + threadIdx.x;
    s_deg[threadIdx.x] = row < P ? p.deg[row] : 0;
    s_run[threadIdx.x / kRowsPerWarp][threadIdx.x % kRowsPerWarp] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = p.blk[2 * blockIdx.x], e = p.blk[2 * blockIdx.x + 1];
    for (int r = 0; r < kPackRows; ++r) {
      s_t[r] = t; s_e[r] = e;
      t += s_deg[r] > 0; e += s_deg[r];
    }
  }
  __syncthreads();
  const int T = p.counts[EAGCN_CNT_T], E = p.counts[EAGCN_CNT_E];
  if (threadIdx.x < kPackRows) {
    const int row = row0 + threadIdx.x;
    if (row < P) {
      const bool act = s_deg[threadIdx.x] > 0;
      const int t = s_t[threadIdx.x];
      p.pos_row[row] = (act && t < p.t_cap) ? t : -1;
      if (act && t < p.t_cap) { p.row_pos[t] = row; p.row_ptr[t] = s_e[threadIdx.x]; }
      if (row % p.N == 0) p.mol_ptr[row / p.N] = min(t, p.t_cap);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.mol_ptr[p.B] = min(T, p.t_cap);
    if (T <= p.t_cap) p.row_ptr[T] = E;
  }
  bool bad = false;
  // ---- per-lane constants of the plane-per-lane gather (sum_v C_v <= 64: every shipped configuration) ----
  // lane l owns one-hot planes l and l + 32 of the concatenated (view, channel) list; lane v < V decodes view v
  int sumC = 0;
  for (int v = 0; v < p.V; ++v) sumC += p.chan[v];
  const bool fast = !kCodes && sumC <= 64 && plane_lanes;
  const float* pl_ptr[2] = {nullptr, nullptr};   // plane base: rel[v] + c * N * N
  long long pl_bstride[2] = {0, 0};              // C_v * N * N: offset of the next molecule
  int my_off = 0, my_C = 0;                      // lane v < V: first plane and channel count of view v
  if (fast) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int c = lane + 32 * h, v = 0;
      if (c < sumC) {
        while (c >= p.chan[v]) { c -= p.chan[v]; ++v; }
        pl_ptr[h] = rel.p[v] + (size_t)c * p.N * p.N;
        pl_bstride[h] = (long long)p.chan[v] * p.N * p.N;
      }
    }
    if (lane < p.V) {
      for (int v = 0; v < lane; ++v) my_off += p.chan[v];
      my_C = p.chan[lane];
    }
  }
  for (int jb = 0; jb < p.N; jb += 256) {
    // ---- phase A: the adjacency rows of all of this warp's active rows, 256 columns at a time, in ONE memory round
    // trip (rows are interleaved over the warps -- w, w+8, ... -- because the active rows of a molecule are
    // consecutive); the per-chunk edge masks go through shared memory (dynamic indexing)
    {
      float a8[kRowsPerWarp][8];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const int lr = r * (kPackRows / kRowsPerWarp) + warp;
        const int row = row0 + lr;
        const bool live = row < P && s_deg[lr] > 0 && s_t[lr] < p.t_cap && s_e[lr] + s_deg[lr] <= p.e_cap;
        const int b = live ? row / p.N : 0, i = live ? row - b * p.N : 0;
        const size_t ba = ((size_t)b * p.N + i) * p.N, bc = (((size_t)b * p.V) * p.N + i) * p.N;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int jj = jb + c * 32 + lane;
          a8[r][c] = (live && jj < p.N) ? adj_at<kCodes>(adj, codes, p, ba, bc, jj) : 0.0f;
        }
      }
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const unsigned mk = __ballot_sync(0xffffffffu, a8[r][c] != 0.0f);
          if (lane == 0) s_mask[warp][r][c] = mk;
        }
      __syncwarp();
    }
    // ---- phase B: per row, per 32-column chunk with bonds ----
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int lr = r * (kPackRows / kRowsPerWarp) + warp;
      const int row = row0 + lr;
      if (row >= P) continue;
      const int deg = s_deg[lr];
      if (deg == 0 || s_t[lr] >= p.t_cap) continue;
      const int e0 = s_e[lr];
      if (e0 + deg > p.e_cap) continue;            // status already flagged by the scan
      const int b = row / p.N, i = row - b * p.N;
      const size_t bc = (((size_t)b * p.V) * p.N + i) * p.N;
      const size_t rowoff = (size_t)i * p.N;
      int run = s_run[warp][r];
      for (int c = 0; c < 8 && jb + c * 32 < p.N; ++c) {
        const unsigned m = s_mask[warp][r][c];
        if (m == 0) continue;
        const int j0 = jb + c * 32, j = j0 + lane;
        const bool nz = (m >> lane) & 1u;
        const int ne = __popc(m);
        const int slot = __popc(m & ((1u << lane) - 1u));          // edge index of this lane inside the chunk
        if (nz) p.colpos[e0 + run + slot] = b * p.N + j;
        if (kCodes) {
          for (int v = 0; v < p.V; ++v) {
            const int C = p.chan[v];
            if (nz) {
              const int cc = __ldg(codes + bc + (size_t)v * p.N * p.N + j);
              if (cc > C) bad = true;
              p.code[(size_t)v * p.e_cap + e0 + run + slot] = (uint8_t)(cc > C ? C : cc);
            }
          }
        } else if (fast) {
          // plane-per-lane gather: for each bonded pair, lane l reads planes l and l + 32 at (i, j); two ballots give the
          // 64-bit set of non-zero planes, lane v cuts out view v's bits (exactly one -> its index, none -> C_v).  Four
          // pairs (8 loads per lane) are in flight before the first ballot.
          const float* q0 = pl_ptr[0] ? pl_ptr[0] + (size_t)b * pl_bstride[0] + rowoff + j0 : nullptr;
          const float* q1 = pl_ptr[1] ? pl_ptr[1] + (size_t)b * pl_bstride[1] + rowoff + j0 : nullptr;
          unsigned mm = m;
          for (int k0 = 0; k0 < ne; k0 += 4) {
            float x0[4], x1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              x0[u] = 0.0f; x1[u] = 0.0f;
              if (k0 + u < ne) {
                const int jj = __ffs(mm) - 1;
                mm &= mm - 1;
                if (q0) x0[u] = __ldg(q0 + jj);
                if (q1) x1[u] = __ldg(q1 + jj);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if ((x0[u] != 0.0f && x0[u] != 1.0f) || (x1[u] != 0.0f && x1[u] != 1.0f)) bad = true;
              const unsigned h0 = __ballot_sync(0xffffffffu, x0[u] != 0.0f);
              const unsigned h1 = __ballot_sync(0xffffffffu, x1[u] != 0.0f);
              if (lane < p.V && k0 + u < ne) {
                const unsigned long long hits = ((unsigned long long)h1 << 32) | h0;
                const unsigned long long sub = (hits >> my_off) & (my_C >= 64 ? ~0ull : ((1ull << my_C) - 1ull));
                const int cnt = __popcll(sub);
                if (cnt > 1) bad = true;
                p.code[(size_t)lane * p.e_cap + e0 + run + k0 + u] = (uint8_t)(cnt == 1 ? __ffsll((long long)sub) - 1 : my_C);
              }
            }
          }
        } else {
          // generic gather (sum_v C_v > 64): the (plane, edge) pairs of every view are spread over the lanes
          // (consecutive lanes -> neighbouring columns of one plane), up to 4 loads per lane in flight
          for (int v = 0; v < p.V; ++v) s_hit[warp][v][lane] = 0;
          __syncwarp();
          const int total = ne * sumC;
          for (int base = 0; base < total; base += 128) {
            float x[4]; int vv[4], cc[4], kk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int idx = base + u * 32 + lane;
              x[u] = 0.0f; vv[u] = 0; cc[u] = 0; kk[u] = 0;
              if (idx < total) {
                int gc = idx / ne, v = 0;
                kk[u] = idx - gc * ne;
                while (gc >= p.chan[v]) { gc -= p.chan[v]; ++v; }
                vv[u] = v; cc[u] = gc;
                const int jj = __fns(m, 0, kk[u] + 1);               // lane position of the k-th edge
                x[u] = __ldg(rel.p[v] + ((((size_t)b * p.chan[v]) + gc) * p.N + i) * p.N + j0 + jj);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (x[u] != 0.0f) {
                if (x[u] != 1.0f) bad = true;
                atomicAdd(&s_hit[warp][vv[u]][kk[u]], (1 << 16) + cc[u]);
              }
          }
          __syncwarp();
          if (nz) {
            for (int v = 0; v < p.V; ++v) {
              const int h = s_hit[warp][v][slot];
              if ((h >> 16) > 1) bad = true;
              p.code[(size_t)v * p.e_cap + e0 + run + slot] = (uint8_t)((h >> 16) == 1 ? (h & 0xFFFF) : p.chan[v]);
            }
          }
          __syncwarp();
        }
        run += ne;
      }
      if (lane == 0) s_run[warp][r] = run;
      __syncwarp();
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&p.counts[EAGCN_CNT_STATUS], EAGCN_ST_NOT_ONEHOT);
}

// Molecule-aligned row tiles for the fused layer kernel (layer_fused.cu): greedy runs of WHOLE molecules with at most
// EAGCN_ROW_TILE active rows, so that every neighbour of a row lies in the tile that holds its projection on-chip.
// One warp: the 32 lanes test the next 32 molecule boundaries at once (a tile holds ~7 molecules at Tox21 sizes).
// A molecule with more than EAGCN_ROW_TILE active rows is cut into plain 128-row tiles and flagged (counts[4]): the
// fused kernel is not used for such a batch (the host knows from N > EAGCN_ROW_TILE that this can happen).
constexpr int kTileMolCap = 8191;   // molecule pointers staged in shared memory (beyond: read through L2)
__device__ void build_row_tiles(const PlanDev& p, int T, const int* __restrict__ s_mp) {
  const int lane = threadIdx.x & 31;
  const int cap = p.t_cap / 32 + 1;
  const bool sm = p.B <= kTileMolCap;
  auto mp = [&](int i) { return min(sm ? s_mp[i] : p.mol_ptr[i], T); };
  int k = 0, b = 0, big = 0;
  while (b < p.B && k < cap) {
    const int s = mp(b);
    if (s >= T) break;
    int last = b;                                      // last boundary index with mol_ptr[last] - s <= 128 rows
    for (int base = b + 1; base <= p.B; base += 32) {
      const int i = base + lane;
      const bool fits = i <= p.B && mp(i) - s <= EAGCN_ROW_TILE;
      const unsigned m = __ballot_sync(0xffffffffu, fits);
      const int run = m == 0xffffffffu ? 32 : __ffs(~m) - 1;   // leading run of fitting boundaries
      last = base + run - 1;
      if (run < 32) break;
    }
    if (last == b) {                                   // the molecule alone exceeds one tile
      big = 1;
      const int e = mp(b + 1);
      for (int r = s; r < e && k < cap; r += EAGCN_ROW_TILE) { if (lane == 0) p.tile_row[k] = r; ++k; }
      b = b + 1;
      continue;
    }
    if (mp(last) > s) { if (lane == 0) p.tile_row[k] = s; ++k; }   // skip runs of bond-less molecules
    b = last;
  }
  if (lane == 0) {
    p.tile_row[k] = T;
    p.counts[EAGCN_CNT_TILES] = k;
    p.counts[EAGCN_CNT_BIGMOL] = big;
  }
}

// one warp per active row: neighbour row ids, reverse edge (binary search), reverse codes; the last CTA builds the
// molecule-aligned row tiles
__global__ void __launch_bounds__(256) pack_link_kernel(PlanDev p) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  if (blockIdx.x == gridDim.x - 1) {
    extern __shared__ int s_mp[];                      // [B + 1] molecule pointers (this CTA only)
    if (p.B <= kTileMolCap) {
      for (int i = threadIdx.x; i <= p.B; i += 256) s_mp[i] = p.mol_ptr[i];
      __syncthreads();
    }
    if (warp == 0) build_row_tiles(p, T, s_mp);
    return;
  }
  const int t = blockIdx.x * 8 + warp;
  if (t >= T) return;
  const int e0 = p.row_ptr[t], e1 = p.row_ptr[t + 1];
  if (e1 > p.e_cap) return;
  const int mypos = p.row_pos[t];
  bool asym = false;
  for (int e = e0 + lane; e < e1; e += 32) {
    const int jpos = p.colpos[e];
    const int jr = p.pos_row[jpos];
    int re = e, jcol = t;
    if (jr < 0) {
      asym = true;
    } else {
      jcol = jr;
      int lo = p.row_ptr[jr], hi = p.row_ptr[jr + 1];
      if (hi > p.e_cap) { asym = true; hi = lo; }
      int found = -1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int c = p.colpos[mid];
        if (c == mypos) { found = mid; break; }
        if (c < mypos) lo = mid + 1; else hi = mid;
      }
      if (found < 0) asym = true; else re = found;
    }
    p.col[e] = jcol;
    p.rev[e] = re;
    for (int v = 0; v < p.V; ++v) p.rcode[(size_t)v * p.e_cap + e] = p.code[(size_t)v * p.e_cap + re];
  }
  if (__any_sync(0xffffffffu, asym) && lane == 0) atomicOr(&p.counts[EAGCN_CNT_STATUS], EAGCN_ST_ASYMMETRIC);
}

// inverse of the packing for one view (parity check: indexing must round-trip bit-exactly)
__global__ void __launch_bounds__(256) unpack_view_kernel(PlanDev p, int v, float* __restrict__ rel_out,
                                                          float* __restrict__ adj_out) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  const int T = min(p.counts[EAGCN_CNT_T], p.t_cap);
  if (t >= T) return;
  const int e0 = p.row_ptr[t], e1 = min(p.row_ptr[t + 1], p.e_cap);
  const int pos = p.row_pos[t];
  const int b = pos / p.N, i = pos - b * p.N;
  const int C = p.chan[v];
  for (int e = e0 + lane; e < e1; e += 32) {
    const int j = p.colpos[e] - b * p.N;
    const int c = p.code[(size_t)v * p.e_cap + e];
    if (adj_out) adj_out[((size_t)b * p.N + i) * p.N + j] = 1.0f;
    if (rel_out && c < C) rel_out[(((size_t)b * C + c) * p.N + i) * p.N + j] = 1.0f;
  }
}

template <bool kCodes>
static int pack_count_impl(const eagcn_plan_t* plan, const void* src, cudaStream_t st) {
  if (!plan_ok_count(plan) || !src) return EAGCN_E_ARG;
  PlanDev p = to_dev(plan);
  const int P = p.B * p.N;
  const int nblk = (P + kPackRows - 1) / kPackRows;
  cudaError_t e = cudaMemsetAsync(p.counts, 0, 8 * sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  EAGCN_PROF("pack_count_kernel", st);
  EAGCN_LAUNCH((pack_count_kernel<kCodes>), nblk, kPackThreads, 0, st)(p, kCodes ? nullptr : (const float*)src,
                                                           kCodes ? (const uint8_t*)src : nullptr);
  EAGCN_LAUNCH_CHECK();
  EAGCN_PROF("pack_scan_kernel", st);
  EAGCN_LAUNCH(pack_scan_kernel, 1, 1024, 0, st)(p, nblk);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

template <bool kCodes>
static int pack_fill_impl(const eagcn_plan_t* plan, const void* src, const void* const* rel, cudaStream_t st) {
  if (!plan_ok(plan) || !src) return EAGCN_E_ARG;
  PlanDev p = to_dev(plan);
  RelPtrs rp;
  for (int v = 0; v < EAGCN_MAX_VIEWS; ++v) rp.p[v] = nullptr;
  if (!kCodes) {
    if (!rel) return EAGCN_E_ARG;
    for (int v = 0; v < p.V; ++v) {
      if (!rel[v] || p.chan[v] <= 0 || p.chan[v] > 254) return EAGCN_E_ARG;
      rp.p[v] = (const float*)rel[v];
    }
  }
  const int P = p.B * p.N;
  const int nblk = (P + kPackRows - 1) / kPackRows;
  EAGCN_PROF("pack_fill_kernel", st);
  int plane_lanes = 1;
  if (!kCodes) {                       // where do the one-hot planes live?  (host-side query, no synchronisation)
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, rel[0]) != cudaSuccess) (void)cudaGetLastError();   // unregistered pointer: not ours
    else if (at.type == cudaMemoryTypeHost) plane_lanes = 0;
  }
  EAGCN_LAUNCH((pack_fill_kernel<kCodes>), nblk, kPackThreads, 0, st)(p, kCodes ? nullptr : (const float*)src,
                                                          kCodes ? (const uint8_t*)src : nullptr, rp, plane_lanes);
  EAGCN_LAUNCH_CHECK();
  EAGCN_PROF("pack_link_kernel", st);
  EAGCN_LAUNCH(pack_link_kernel, (p.t_cap + 7) / 8 + 1, 256, p.B <= kTileMolCap ? (size_t)(p.B + 1) * 4 : 0, st)(p);
  EAGCN_LAUNCH_CHECK();
  return 0;
}

}  // namespace eagcn

using namespace eagcn;

extern "C" int eagcn_pack_count(const eagcn_plan_t* plan, const void* adj, void* stream) {
  return pack_count_impl<false>(plan, adj, (cudaStream_t)stream);
}
extern "C" int eagcn_pack_fill(const eagcn_plan_t* plan, const void* adj, const void* const* rel, void* stream) {
  return pack_fill_impl<false>(plan, adj, rel, (cudaStream_t)stream);
}
extern "C" int eagcn_pack_count_codes(const eagcn_plan_t* plan, const void* codes, void* stream) {
  return pack_count_impl<true>(plan, codes, (cudaStream_t)stream);
}
extern "C" int eagcn_pack_fill_codes(const eagcn_plan_t* plan, const void* codes, void* stream) {
  return pack_fill_impl<true>(plan, codes, nullptr, (cudaStream_t)stream);
}

extern "C" int eagcn_unpack_view(const eagcn_plan_t* plan, int64_t v, void* rel_out, void* adj_out, void* stream) {
  if (!plan_ok(plan) || v < 0 || v >= plan->V) return EAGCN_E_ARG;
  PlanDev p = to_dev(plan);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nn = (size_t)p.B * p.N * p.N;
  cudaError_t e;
  if (rel_out) { e = cudaMemsetAsync(rel_out, 0, nn * p.chan[v] * sizeof(float), st); if (e) return (int)e; }
  if (adj_out) { e = cudaMemsetAsync(adj_out, 0, nn * sizeof(float), st); if (e) return (int)e; }
  EAGCN_PROF("unpack_view_kernel", st);
  EAGCN_LAUNCH(unpack_view_kernel, (p.t_cap + 7) / 8, 256, 0, st)(p, (int)v, (float*)rel_out, (float*)adj_out);
  EAGCN_LAUNCH_CHECK();
  return 0;
}
