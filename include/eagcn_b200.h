/*
 * eagcn_b200 -- C ABI of the B200-native EAGCN multi-view edge-attention graph-convolution path.
 *
 * Drop-in boundary.  The reference (Luckick/EAGCN) has no FFI of its own: its
 * "plugin surface" is the nn.Module forward of eagcn_pytorch/layers.py.  Each entry point below
 * replaces a span of that Python/ATen code; the host side (eagcn_b200/layers.py, a mirror of the
 * reference's layers.py classes) calls these through ctypes with raw device pointers taken from
 * torch tensors (tensor.data_ptr()) and torch's current CUDA stream.
 *
 * Conventions: every pointer is a DEVICE pointer to contiguous row-major memory unless it says
 * "host"; sizes are int64_t; functions are asynchronous with respect to the host, never synchronise
 * the device, never throw and never exit.  The compute entry points keep no state between calls:
 * everything they read or write is passed in.  Two exceptions, both documented where declared:
 * the eagcn_set_* engine switches (process-wide, for A/B measurements: set them before the first
 * compute call, not concurrently with one) and the diagnostics block at the end (launch counter,
 * per-kernel profiler: not thread-safe).  One CUDA device per process (torch.distributed's
 * one-process-per-GPU model): the kernels' shared-memory opt-ins are cached per process.
 * Return value: 0 = ok, negative = invalid argument (EAGCN_E_*), positive = cudaError_t of the
 * failing runtime call / launch.
 *
 * All structs are made of 8-byte fields only (void*, int64_t, double) so that the ctypes mirror in
 * eagcn_b200/_lib.py cannot disagree about padding.
 */
#ifndef EAGCN_B200_H_
#define EAGCN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAGCN_ABI_VERSION 18
#define EAGCN_MAX_VIEWS 16
#define EAGCN_ROW_TILE 128          /* packed-row capacity granularity (one MMA tile of rows) */

#define EAGCN_E_ARG (-1)            /* null pointer / bad size */
#define EAGCN_E_UNSUPPORTED (-2)    /* shape outside what the kernels cover */

/* bits of plan.counts[2] (device-side status word, read lazily by the host) */
#define EAGCN_ST_ADJ_NOT_01 1       /* adjacency value other than 0.0 / 1.0             */
#define EAGCN_ST_NOT_ONEHOT 2       /* relation tensor not one-hot 0/1 on a bonded pair */
#define EAGCN_ST_ASYMMETRIC 4       /* adj[b,i,j] != adj[b,j,i]                         */
#define EAGCN_ST_EDGE_CAP   8       /* more directed edges than e_cap                    */
#define EAGCN_ST_ROW_CAP   16       /* more active rows than t_cap                       */

/* indices into plan.counts (int32 device array of 8) */
#define EAGCN_CNT_T 0               /* number of active atom rows (rows with >=1 bond)  */
#define EAGCN_CNT_E 1               /* number of directed edges (nnz of adj)            */
#define EAGCN_CNT_STATUS 2
#define EAGCN_CNT_TILES 3           /* number of molecule-aligned row tiles (plan.tile_row)  */
#define EAGCN_CNT_BIGMOL 4          /* 1: some molecule has more than EAGCN_ROW_TILE active rows (no fused layer kernel) */

/*
 * Graph plan: the packed form of one padded batch (adj [B,N,N] + V one-hot relation tensors).
 * Replaces, for every layer and for forward AND backward, the reference's per-layer mask /
 * identity construction (layers.py:294-304) and the 1x1-conv attention-score input
 * (layers.py:82: the one-hot planes become one uint8 code per edge and view).
 * Active rows (m[b,i] = max_j adj[b,i,j] = 1, layers.py:295) are numbered 0..T-1 in ascending
 * flat position b*N+i; edges are CSR over those rows, neighbours ascending.
 */
typedef struct eagcn_plan {
  int64_t B, N, V;                  /* batch, padded atoms, views                          */
  int64_t t_cap, e_cap;             /* row capacity (multiple of EAGCN_ROW_TILE), edge cap  */
  int64_t chan[EAGCN_MAX_VIEWS];    /* C_v: channels of relation tensor v                  */
  void* counts;                     /* int32 [8]                                           */
  void* deg;                        /* int32 [B*N]        scratch (degree per flat row)    */
  void* blk;                        /* int32 [2*nblk+2]   scratch, nblk = ceil(B*N/32)     */
  void* pos_row;                    /* int32 [B*N]        flat position -> row or -1       */
  void* row_pos;                    /* int32 [t_cap]      row -> flat position             */
  void* row_ptr;                    /* int32 [t_cap+1]    CSR offsets                      */
  void* mol_ptr;                    /* int32 [B+1]        first row of each molecule       */
  void* col;                        /* int32 [e_cap]      neighbour row                    */
  void* colpos;                     /* int32 [e_cap]      neighbour flat position          */
  void* rev;                        /* int32 [e_cap]      index of the reverse edge        */
  void* code;                       /* uint8 [V][e_cap]   type_v(i,j); C_v = all-zero vec  */
  void* rcode;                      /* uint8 [V][e_cap]   type_v(j,i)                      */
  void* tile_row;                   /* int32 [t_cap/32+2] first row of each molecule-aligned row tile: whole molecules,
                                     * <= EAGCN_ROW_TILE rows each (greedy), tile_row[n_tiles] = T.  The fused layer
                                     * kernel runs one tensor-core row tile per entry, so every neighbour of a row is
                                     * inside the tile whose projection it holds on-chip                            */
} eagcn_plan_t;

/*
 * Parameters of one GraphConv_Layer (layers.py:266-288): V GraphConv_blocks, each
 * att.weight [1,C_v,1,1], self_r [1], graph_conv.weight [fin,fo_v] / .bias [fo_v],
 * batch_norm.bn.{weight,bias,running_mean,running_var,num_batches_tracked}.
 */
typedef struct eagcn_layer {
  int64_t fin, fo_tot, V;
  int64_t fo[EAGCN_MAX_VIEWS];
  int64_t off[EAGCN_MAX_VIEWS + 1]; /* prefix sums of fo                                   */
  void* att_w[EAGCN_MAX_VIEWS];     /* f32 [C_v]                                           */
  void* self_r[EAGCN_MAX_VIEWS];    /* f32 [1]                                             */
  void* W[EAGCN_MAX_VIEWS];         /* f32 [fin, fo_v]                                     */
  void* bias[EAGCN_MAX_VIEWS];      /* f32 [fo_v]                                          */
  void* gamma[EAGCN_MAX_VIEWS];     /* f32 [fo_v]  bn.weight                               */
  void* beta[EAGCN_MAX_VIEWS];      /* f32 [fo_v]  bn.bias                                 */
  void* run_mean[EAGCN_MAX_VIEWS];  /* f32 [fo_v]  updated in training                     */
  void* run_var[EAGCN_MAX_VIEWS];   /* f32 [fo_v]                                          */
  void* nbt[EAGCN_MAX_VIEWS];       /* i64 [1]     num_batches_tracked                     */
} eagcn_layer_t;

/* Per-call buffers of one layer invocation; everything is caller-allocated (torch caching
 * allocator).  Saved-for-backward: Z, Y, invR, mean, invstd, sums, sig, wall (+ H). */
typedef struct eagcn_work {
  void* H;        /* f32 [t_cap, fin]      packed input rows                               */
  void* Z;        /* f32 [t_cap, fo_tot]   H @ W_all                                       */
  void* Y;        /* f32 [t_cap, fo_tot]   pre-BatchNorm output                            */
  void* X;        /* f32 [t_cap, fo_tot]   layer output, packed rows                       */
  void* invR;     /* f32 [V, t_cap]        1 / attention row sum                           */
  void* wall;     /* f32 [fin, fo_tot]     concatenated projection weights                 */
  void* ball;     /* f32 [4, fo_tot]       bias, gamma, beta, (spare)                      */
  void* sig;      /* f32 [V, 257]          sigmoid(att_w) table, [256] = sigmoid(self_r)   */
  void* partial;  /* f32 [n_tiles, 2, fo_tot]  per-tile statistics partials                */
  void* sums;     /* f64 [2, fo_tot]       batch sums (all-reduced by the host for global BN) */
  void* mean;     /* f32 [fo_tot]                                                         */
  void* invstd;   /* f32 [fo_tot]                                                         */
  void* rng;      /* u64 [2]               philox seed, offset (device; graph-replay safe) */
  int64_t training;     /* bit4 (16): no backward pass will follow (the fused forward then skips the store of Z);
                         * bit2: eagcn_layer_prepare was already called for wall / wallT / wsplit / ball / sig;
                         * bit0: training mode (batch statistics, dropout); bit1: the host all-reduces `sums`
                         * between forward_a and forward_b and `bsums` between backward_a and backward_b
                         * (global-batch BatchNorm); without bit1 the reductions are fused into fewer kernels */
  int64_t rng_stream;   /* distinguishes layers sharing one rng state                      */
  int64_t m_total;      /* BatchNorm population: B*N of the (global) padded batch           */
  int64_t n_pad;        /* padded width N used for the (N - deg)*1e-9 normaliser term       */
  double p_drop;
  double eps;
  double momentum;
  /* backward only */
  void* dX;       /* f32 [t_cap, fo_tot]   gradient wrt X (packed)                         */
  void* dY;       /* f32 [t_cap, fo_tot]   workspace                                       */
  void* Q;        /* f32 [t_cap, fo_tot]   workspace: sum_v A_v^T dY_v                     */
  void* dH;       /* f32 [t_cap, fin]      out; NULL = the layer input needs no gradient (product skipped) */
  void* dwall;    /* f32 [fin * fo_tot]    out: weight gradients, view-blocked: view v's [fin, fo_v]
                   *                        block is contiguous at offset fin * off[v]            */
  void* dvec;     /* f32 [3, fo_tot]       out: dbias, dgamma, dbeta                       */
  void* datt;     /* f32 [V, 257]          out: d att_w (first C_v), [256] = d self_r      */
  void* bsums;    /* f64 [2, fo_tot]       backward batch sums (sum g, sum g*xhat)         */
  void* gemm_ws;  /* f32 split-K workspace, gemm_ws_bytes bytes                            */
  int64_t gemm_ws_bytes;
  void* wallT;    /* f32 [2, fo_tot, fin]  W_all transposed, split hi / lo (K-major B operand of Z = H W)   */
  void* wsplit;   /* f32 [2, fin, fo_tot]  W_all split hi / lo          (K-major B operand of dH = Q W^T)   */
  int64_t phase;  /* eagcn_layer_backward_b only.  0: everything.  Otherwise a bit set of the parts to launch, so that
                   * the host can put independent parts on different streams: 1 = aggregation backward (Q, attention
                   * partials; needs backward_a), 2 = dH = Q W^T (needs 1), 4 = dW = H^T Q + d att / d self_r sums
                   * (needs 1; independent of 2)                                                          */
  void* tickets;  /* i32 [fo_tot/128 + 1], ZERO when first used (the kernels leave it zero), or NULL.  eagcn_layer_backward_a
                   * then reduces the BatchNorm backward sums in the LAST CTA of each 128-channel block (fixed tile order:
                   * deterministic) instead of a separate reduction launch -- for t_cap <= 6 144 rows (192 tiles); larger
                   * batches keep the separate launch, which measured faster there.  One array per stream of calls      */
} eagcn_work_t;

int eagcn_version(void);
/* sizeof of the ABI structs as compiled (0: eagcn_plan_t, 1: eagcn_layer_t, 2: eagcn_work_t): lets a
 * foreign-function binding verify its mirror of the layouts at load time                                          */
int64_t eagcn_sizeof(int which);
/* number of CTAs/tiles the statistics partial buffer must hold for a given t_cap */
int64_t eagcn_stat_tiles(int64_t t_cap);
/* floats the `partial` buffer must hold (BatchNorm partials, then attention-gradient partials) */
int64_t eagcn_partial_floats(int64_t t_cap, int64_t fo_tot, int64_t V);
/* bytes of split-K workspace eagcn_layer_backward_b wants for dW_all = H^T Q */
int64_t eagcn_gemm_workspace_bytes(int64_t fin, int64_t fo_tot, int64_t t_cap);

/* --- graph plan (replaces layers.py:294-304 + the one-hot inputs of layers.py:82) -------- */
/* phase 1: degrees, active-row / edge counts -> counts[0..1] (host may read them to size the
 * plan exactly; not needed when capacities are known)                                        */
int eagcn_pack_count(const eagcn_plan_t* plan, const void* adj, void* stream);
/* phase 2: fill CSR + codes.  rel[v]: f32 [B, C_v, N, N] device pointers (host array of V)   */
int eagcn_pack_fill(const eagcn_plan_t* plan, const void* adj, const void* const* rel, void* stream);
/* packed uint8 data boundary: same plan from codes u8 [B,V,N,N]
 * (255 = no bond) instead of adj + one-hot planes                                            */
int eagcn_pack_count_codes(const eagcn_plan_t* plan, const void* codes, void* stream);
int eagcn_pack_fill_codes(const eagcn_plan_t* plan, const void* codes, void* stream);
/* inverse of packing, for the bit-exactness check: one-hot f32 [B,C_v,N,N] of view v and adj  */
int eagcn_unpack_view(const eagcn_plan_t* plan, int64_t v, void* rel_out, void* adj_out, void* stream);

/* --- dense <-> packed rows --------------------------------------------------------------- */
int eagcn_rows_gather(const eagcn_plan_t* plan, const void* dense, void* packed, int64_t F, void* stream);
/* dense[b,i,:] = packed[row] for active rows, 0 elsewhere (the "* mask3" of layers.py:313)   */
int eagcn_rows_scatter(const eagcn_plan_t* plan, const void* packed, void* dense, int64_t F, void* stream);
/* sum read-out over atoms (models.py:108): out[b,:] = sum_rows packed; and its backward      */
int eagcn_readout_sum(const eagcn_plan_t* plan, const void* packed, void* out, int64_t F, void* stream);
int eagcn_readout_sum_bwd(const eagcn_plan_t* plan, const void* dout, void* dpacked, int64_t F, void* stream);

/* --- one GraphConv_Layer (layers.py:293-325, structure == 'Concate') ---------------------- */
/* part A: weights concat + sigmoid tables, Z = H @ W_all (layers.py:40), attention score /
 * row-normalise / aggregate / +bias (layers.py:82-92,39,43) -> Y, BatchNorm batch sums -> sums */
/* optional first step of part A, callable on another stream before the plan's arrays exist (only plan->V and
 * plan->chan are read): fills work.wall / wallT / wsplit / ball / sig from the parameters.  eagcn_layer_forward_a skips
 * it when bit2 of work.training is set.                                                                          */
int eagcn_layer_prepare(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w, void* stream);
int eagcn_layer_forward_a(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w, void* stream);
/* Part A is ONE kernel launch (layer_fused.cu) when the layout allows it -- padded width N <= EAGCN_ROW_TILE, every
 * fo_v a multiple of 4, 16-byte aligned rows, tcgen05 engine: per molecule-aligned row tile the projection
 * Z = H W_all runs on the tensor cores (TMA -> tcgen05 -> TMEM), the tile of Z goes TMEM -> shared memory and the
 * attention score lookup, row normalisation, neighbour aggregation, bias and BatchNorm partial sums are the epilogue;
 * Z is written once only as the activation backward needs (bit4 of work.training set: not at all).  Otherwise
 * projection GEMM + aggregation kernel.  eagcn_set_fwd_fused(0) forces the two-kernel form (measurements).        */
int eagcn_set_fwd_fused(int on);
int eagcn_get_fwd_fused(void);
/* part B: BatchNorm finalize (+ running stats, layers.py:408-412), ReLU, dropout
 * (layers.py:93-94), concat (layers.py:313) -> X                                              */
int eagcn_layer_forward_b(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w, void* stream);
/* backward part A: dX -> BatchNorm backward batch sums (bsums)                                */
int eagcn_layer_backward_a(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w, void* stream);
/* backward part B: dY, attention gradients, Q = sum_v A_v^T dY_v, dH = Q @ W_all^T,
 * dW_all = H^T @ Q, dbias / dgamma / dbeta / d att / d self_r                                 */
int eagcn_layer_backward_b(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const eagcn_work_t* w, void* stream);

/* --- returned attention A1 (layers.py:83,318): dense [V,B,N,N] and its backward ------------ */
int eagcn_attention_dense(const eagcn_plan_t* plan, const eagcn_layer_t* layer, void* A_out, void* stream);
int eagcn_attention_dense_bwd(const eagcn_plan_t* plan, const eagcn_layer_t* layer, const void* dA, void* datt, void* stream);

/* --- test hook: the keep mask eagcn_layer_forward_b draws (u8 [t_cap, fo_tot]) ------------- */
int eagcn_dropout_mask(const eagcn_plan_t* plan, const eagcn_work_t* w, int64_t fo_tot, void* keep_out, void* stream);
/* snapshot[0..1] = state[0..1] (philox seed, offset: device u64 [2]); state[1] += increment.  One call per dropout
 * site and forward pass: forward and backward of that site both read the snapshot (F.dropout, layers.py:94).        */
int eagcn_rng_fork(void* state, void* snapshot, int64_t increment, void* stream);
/* n call sites in one launch: snapshots u64 [n][2], snapshot i = (seed, offset + i*increment); offset += n*increment */
int eagcn_rng_fork_n(void* state, void* snapshots, int64_t n, int64_t increment, void* stream);
/* same generator over a flat index range (eagcn_bn_act_forward draws element b*C+c of stream rng_stream):
 * keep_out u8 [total]                                                                          */
int eagcn_dropout_mask_flat(const void* rng, int64_t rng_stream, double p_drop, int64_t total, void* keep_out, void* stream);

/* --- fused BatchNorm1d (+ReLU)(+dropout) of the read-out head ---------------------------------- */
/* y[B,C] = dropout(relu(BatchNorm1d(x)))  -- reference models.py:112 (Graph_BN: relu = 0, p_drop = 0), :114-116
 * (relu(bn_den1(x)) then F.dropout), :119 (relu(bn_den2(x))).  training != 0: batch statistics (two-pass), running
 * statistics / num_batches_tracked updated like nn.BatchNorm1d; training == 0: running statistics.  mean_out /
 * invstd_out [C] receive the statistics used (saved for backward).  Dropout keeps element i iff the library's Philox
 * stream (rng = device int64 {seed, offset}, rng_stream) says so for flat index i (eagcn_dropout_mask_flat).          */
int eagcn_bn_act_forward(const void* x, void* y, const void* gamma, const void* beta, void* run_mean, void* run_var,
                         void* nbt, void* mean_out, void* invstd_out, int64_t B, int64_t C, int training, int relu,
                         double p_drop, const void* rng, int64_t rng_stream, double momentum, double eps, void* stream);
/* autograd replay of the above: dx [B,C], dgamma [C], dbeta [C] from dy and the saved x, mean, invstd.             */
int eagcn_bn_act_backward(const void* x, const void* dy, const void* gamma, const void* beta, const void* mean,
                          const void* invstd, void* dx, void* dgamma, void* dbeta, int64_t B, int64_t C, int training,
                          int relu, double p_drop, const void* rng, int64_t rng_stream, void* stream);

/* 0 (default): float4 kernels (4 or 8 channels per CTA, one Philox call per 4 elements) when C % 4 == 0, the arrays are
 * 16-byte aligned and B <= 1024; 1: always the 32-channel kernels.  Process-wide.                                 */
int eagcn_set_bn_act_mode(int mode);
int eagcn_get_bn_act_mode(void);

/* --- dense layers of the head ------------------------------------------------------------------ */
/* C[M,N] = op(A) . op(B) in strict fp32, row-major; transX != 0: the operand is stored transposed (A as [K,M], B as
 * [N,K]); lda / ldb = row strides in elements; C is dense (ldc = N).  This is `torch.mm(input, weight)` of reference
 * layers.py:382-388 (Dense.forward) and its two autograd products, on the small-matrix tile kernel (mm_tile.cu): 32x32 tiles, split-K combined INSIDE the launch (the last
 * CTA to reach a tile sums the partials in z order: bit-reproducible, no second kernel).  `tickets`: device int32
 * [eagcn_mm_tile_tickets(M,N)], zero when first used; the kernel leaves it zero, so one array can serve every call that
 * is ordered after the previous one (calls that may run concurrently need separate arrays).  ws:
 * eagcn_mm_tile_workspace_bytes(M,N,K) bytes (0: neither ws nor tickets are needed).                                 */
int64_t eagcn_mm_tile_workspace_bytes(int64_t M, int64_t N, int64_t K);
int64_t eagcn_mm_tile_tickets(int64_t M, int64_t N);
int eagcn_mm_tile(const void* A, int64_t lda, int transA, const void* B, int64_t ldb, int transB, void* C, int64_t M,
                  int64_t N, int64_t K, void* ws, int64_t ws_bytes, void* tickets, void* stream);

/* --- projection GEMM engine ------------------------------------------------------------------ */
/* 0 (default): tcgen05 3xTF32 tensor-core kernel where the operand layout allows it (16-byte aligned
 * rows), FFMA kernel otherwise; 1: always the FFMA kernel; 2: tcgen05 for the K-major products
 * (Z = H W, dH = Q W^T), FFMA for the MN-major weight gradient (dW = H^T Q).  Process-wide.    */
int eagcn_set_gemm_mode(int mode);
int eagcn_get_gemm_mode(void);
/* --- aggregation engine ------------------------------------------------------------------------ */
/* 0 (default): shared-memory tile kernels for the neighbour aggregation (layers.py:90,39) and its backward when
 * the layout allows it (every fo_v a multiple of 4 and <= 512, 16-byte aligned buffers); the backward tile kernel
 * also applies the BatchNorm/ReLU/dropout backward on the fly, so eagcn_layer_backward_b then leaves work.dY
 * unwritten.  1: always the generic warp-per-row kernels (dY materialised).  Process-wide.               */
/* k-block of the K-major tcgen05 products: 0 (default) = chosen per shape by the pipeline model, 16 = 64-byte rows
 * (SWIZZLE_64B, up to 6 stages), 32 = 128-byte rows (SWIZZLE_128B).  Process-wide; for measurements.            */
int eagcn_set_tc_bk(int bk);
/* tensor-core precision of the projection products (layers.py:40 and its autograd products).  3 (default): fp32-faithful
 * 3xTF32 -- every fp32 operand split exactly into hi + lo, three TF32 passes, two fp32 accumulators: meets the 1e-5 parity
 * bar.  1: ONE TF32 pass on the raw operands (the tensor core truncates to a 10-bit mantissa): ~1e-3 relative error, the
 * analogue of BASELINE.json's "bf16" configuration -- reported beside the strict mode by bench.py, never the default.
 * Process-wide; for measurements.                                                                                 */
int eagcn_set_tc_passes(int passes);
int eagcn_get_tc_passes(void);
/* where the hi / lo split of the ACTIVATION operand of the 3xTF32 products lives.  Bit 0: K-major products (Z = H W,
 * dH = Q W^T), bit 1: the split-K product dW = H^T Q.  Bit set (default 3): the transform warps write hi / lo straight
 * into TENSOR MEMORY (tcgen05.st) and the MMAs take A from TMEM -- no hi / lo copies and no A operand reads in shared
 * memory, whose bandwidth bounds these kernels.  Bit clear: hi / lo copies in shared memory.  Same arithmetic, bit-
 * identical results.  Process-wide; for measurements.                                                              */
int eagcn_set_tc_a_tmem(int mask);
int eagcn_get_tc_a_tmem(void);
int eagcn_gemm_trace(void* buf, int64_t max_launches);   /* diagnostic: clock stamps of the GEMM pipeline (see .cu) */
int64_t eagcn_gemm_trace_stride(void);
int eagcn_set_agg_mode(int mode);
int eagcn_get_agg_mode(void);
/* --- forward BatchNorm fusion ------------------------------------------------------------------- */
/* 0 (default): in training mode with per-replica statistics eagcn_layer_forward_b is ONE launch (reduction of the tile
 * partials + finalize + normalise / ReLU / dropout; bit-identical statistics); 1: the separate stat_reduce and
 * bn_apply kernels.  Process-wide.                                                                              */
int eagcn_set_fuse_mode(int mode);
int eagcn_get_fuse_mode(void);
/* --- programmatic dependent launch ------------------------------------------------------------- */
/* 1 (default): kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization -- every kernel begins
 * with griddepcontrol.launch_dependents + griddepcontrol.wait, so results are those of plain stream order while the
 * scheduling latency of the next launch overlaps the running kernel; 0: plain launches.  Process-wide.            */
int eagcn_set_pdl(int on);
int eagcn_get_pdl(void);
/* stand-alone projection product  C[m_cap, N] = A[m_cap, K] . B[N, K]^T  (fp32, rows contiguous; lda/ldb/ldc
 * in elements).  Only the first min(*m_dev, m_cap) rows are live (m_dev: device int32); the remaining
 * rows of C are written as zeros.  engine: 0 = tcgen05 3xTF32 (EAGCN_E_UNSUPPORTED if the layout does
 * not allow it), 1 = FFMA.  This is reference layers.py:40 (torch.mm(support, W)) in isolation.    */
int eagcn_gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t m_cap,
                  int64_t N, int64_t K, const void* m_dev, int engine, void* stream);

/* stand-alone weight-gradient product  C[M, N] = A[k_cap, M]^T . B[k_cap, N]  (fp32 row-major, lda >= M,
 * ldb >= N, C dense [M, N]).  Only the first min(*k_dev, k_cap) rows take part; rows beyond that MUST be zero
 * in both operands (the layer kernels keep their slack rows zero).  ws: eagcn_gemm_workspace_bytes(M, N, k_cap)
 * bytes of split-K workspace.  engine as in eagcn_gemm_nt.  This is the autograd product d(weight) of
 * reference layers.py:40.                                                                          */
int eagcn_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t M, int64_t N, int64_t k_cap,
                  const void* k_dev, void* ws, int64_t ws_bytes, int engine, void* stream);

/* --- diagnostics (not on the data path; not thread-safe) ------------------------------------ */
/* kernels launched by this library since load (bench.py: gpu_launches; under CUDA-graph replay the
 * count taken while capturing one step is the number of kernel nodes replayed per step)          */
int64_t eagcn_launch_count(void);
/* keeps `stream` busy for ~ns nanoseconds (<= 0.1 s): lets the host queue a whole eager step behind it, so that the
 * per-kernel events below bracket kernels, not launch latency                                        */
int eagcn_spin(int64_t ns, void* stream);
/* opt-in per-kernel CUDA-event timing: enable(1) / disable(0), both clear what was recorded      */
int eagcn_profile(int enable);
/* writes {"kernel": [launches, total_ms], ...} into buf; returns bytes written, 0 if cap too small */
int64_t eagcn_profile_report(char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* EAGCN_B200_H_ */
