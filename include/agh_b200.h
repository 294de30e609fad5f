/*
 * agh_b200.h -- C ABI of the B200-native AGH rollout library (libagh_b200.so).
 *
 * The reference (RoyalSkye/AGH, Construction_based/, "CB/") has no FFI layer: its hot path is
 * Python calling PyTorch ATen.  This header is the boundary a maintainer binds with ctypes
 * (INTEGRATION.md shows the stub): plain device pointers, sizes and a CUDA stream handle; no
 * torch types; every function returns 0 on success or a negative AGH_E* code, never throws, and
 * owns no memory (the caller allocates every output, e.g. with torch.empty).  All entry points
 * are asynchronous on `stream` and re-entrant per stream; there is no global state besides the
 * last-error string.
 *
 * Shapes: B instances, N flights, n = N+1 nodes (node 0 = depot), D = 128, H = 8 heads of 16,
 * FF = 512, 3 encoder layers, 92 airport locations (91 gates + depot).
 *
 * Each function cites the reference code it replaces (file:line under CB/).
 */
#ifndef AGH_B200_H
#define AGH_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGH_D 128
#define AGH_H 8
#define AGH_FF 512
#define AGH_LAYERS 3
#define AGH_NODE_SIZE 92
#define AGH_KEY_STRIDE 128 /* floats per node row of the staged K'/V/LK' matrices; 16-byte chunks XOR-swizzled by (node & 7) */
#define AGH_MAX_N 320

enum {
  AGH_OK = 0,
  AGH_E_ARG = -1,      /* bad argument (null pointer, size out of range) */
  AGH_E_CUDA = -2,     /* CUDA runtime error; see agh_last_error() */
  AGH_E_UNSUPPORTED = -3
};

/* decode modes (attention_model.py:372-390) */
enum { AGH_GREEDY = 0, AGH_SAMPLING = 1, AGH_REPLAY = 2 };

/* bits of the per-instance status word written by agh_get_costs (problem_agh.py:24-27,44,64) */
enum { AGH_TOUR_INVALID = 1, AGH_CAPACITY_EXCEEDED = 2, AGH_TW_VIOLATED = 4 };
/* bit of the per-instance status word written by agh_decode: the instance still had unvisited nodes when the step
 * budget t_stride ran out (unvisited nodes that are all infeasible: the reference's Python loop would spin forever) */
enum { AGH_DECODE_STUCK = 8 };

const char* agh_last_error(void);
int agh_version(void);
/* number of CUDA kernels this library has launched since it was loaded (bench.py's gpu_launches) */
unsigned long long agh_launch_count(void);

/* ---- packed weights -------------------------------------------------------------------------
 * One f32 blob built by the host from the reference state_dict (agh_b200/nets/weights.py);
 * order and sizes are fixed by the architecture, agh_wblob_floats() returns the total.
 *   loc_emb[92][128] | init_w[3][128] | init_b[128] |
 *   3 x { wqkv[128][384] | wo[128][128] | bn1_scale[128] | bn1_shift[128] | w1t[128][512] | b1[512] |
 *         w2t[512][128] | b2[128] | bn2_scale[128] | bn2_shift[128] } |
 *   wdec[128][512] (K'/4 | V | LK'=W_out^T W_lk/sqrt(128) | P_sc) | wfc_t[128][128] |
 *   fleet_emb[10][128] | w_cap[128] | w_time[128]
 */
size_t agh_wblob_floats(void);

/* ---- a1: fleet-batch assembly (train.py:40-57 == train.py:167-184 == eval.py:97-113) --------
 * From the raw instance arrays and the per-level tw_left table build the node vectors of fleet f.
 *   loc[B][N] i64 (1..91), type[B][N] i64 (0..2), departure[B][N] f32, demand_all[B][10][N] f32,
 *   tw_left_levels[6][B][N] f32, duration3[3] f64 (fleet_info['duration'][f]),
 *   next_duration3[3] f64 (fleet_info['next_duration'][precedence[f]]), nd_is_f32: level-4 quirk.
 * Outputs: coord[B][n] u8, demand[B][N] f32, tw_left[B][n] f32, tw_right[B][n] f64, duration[B][n] f64.
 */
int agh_fleet_task(const int64_t* loc, const int64_t* type, const float* departure, const float* demand_all,
                   const float* tw_left_levels, int level, int fleet /*1..10*/, const double* duration3,
                   const double* next_duration3, int nd_is_f32, int B, int N, uint8_t* coord, float* demand,
                   float* tw_left, double* tw_right, double* duration, void* stream);

/* train.py:67  bat_tw_left[level+1] = max(bat_tw_left[level+1], serve_time[:,1:]) */
int agh_chain_tw_left(float* tw_left_levels, int level, const float* serve /*[B][n]*/, int B, int N, void* stream);

/* ---- a2+a3: init embedding + 3-layer graph attention encoder, eval-mode BatchNorm ------------
 * (attention_model.py:272-294, graph_encoder.py:56-111,135-172,195-207)
 *   h_out[B][n][128] f32.  workspace: agh_encoder_workspace_bytes(B,N) bytes.
 */
size_t agh_encoder_workspace_bytes(int B, int N);
/* GEMM back end of agh_encode / agh_precompute: 1 = tcgen05 tensor cores with the two-term FP16 split (default; fp32-level accuracy),
 * 0 = fp32 CUDA-core SGEMM (kept for A/B measurements; env AGH_GEMM=simt selects it at load time) */
int agh_set_gemm_backend(int backend);
/* Feed-forward sublayer of the encoder: 1 (default) or 2 = one fused tcgen05 kernel per layer with the hidden activations kept on
 * chip, for every row count (the choice must not depend on the batch size: per-instance results are bit-identical across batch
 * sizes); 0 = two GEMMs with the [M][512] hidden activations in HBM (A/B measurements and tests; env AGH_FFN=unfused selects it) */
int agh_set_ffn_fused(int fused);
/* Encoder self-attention: 1 = auto (default): the tcgen05 kernel with S / P / O in TMEM for graphs of 72 .. 112 nodes, the mma.sync
 * kernel otherwise (a function of the graph size only, never of the batch size); 0 = always mma.sync; 2 = tcgen05 for every graph
 * of up to 112 nodes (A/B measurements and tests; env AGH_MHA=mma / tc selects 0 / 2 at load time) */
int agh_set_mha_tc(int on);
/* The self-attention of one encoder layer alone (graph_encoder.py:83-104, all 8 heads, no projection):
 *   qkv [3][B*n][128] f32 (Q, K, V of all instances as three stacked row-major matrices) -> heads [B*n][128] f32 */
int agh_encoder_mha(const float* qkv, int B, int N, float* heads, void* stream);
/* The feed-forward sublayer of encoder layer `layer` alone (graph_encoder.py:159-172, eval-mode BatchNorm):
 *   y[M][128] = BN2(x + W2 relu(W1 x + b1) + b2);  x != y.  workspace: agh_encoder_ffn_workspace_bytes(M) bytes. */
size_t agh_encoder_ffn_workspace_bytes(int M);
int agh_encoder_ffn(const float* wblob, int layer, const float* x, int M, float* y, void* workspace, void* stream);
/* twr_f32: tw_right held f32 values in the caller (fleet 10) -> tw_right/1440 is an f32 division */
int agh_encode(const float* wblob, const uint8_t* coord, const float* demand, const float* tw_left,
               const double* tw_right, int twr_f32, int B, int N, float* h_out, void* workspace, void* stream);

/* init embedding only (attention_model.py:272-294): h0[B][n][128] */
int agh_init_embed(const float* wblob, const uint8_t* coord, const float* demand, const float* tw_left,
                   const double* tw_right, int twr_f32, int B, int N, float* h0, void* stream);

/* encoder layers only (graph_encoder.py:195-207) on caller-provided initial embeddings h0[B][n][128];
 * h0 may alias neither h_out nor the workspace. */
int agh_encode_layers(const float* wblob, const float* h0, int B, int N, float* h_out, void* workspace, void* stream);

/* ---- a4: decoder precompute (attention_model.py:392-414,604-611) ---------------------------
 *   h[B][n][128] -> keys[3][B][n][128] (K', V, LK': three plain row-major matrices), psc[B][n][128], qbase[B][128]
 *   fleet_ids[B] i32 (0..9) or NULL with `fleet` used for every instance.
 */
int agh_precompute(const float* wblob, const float* h, const int32_t* fleet_ids, int fleet, int B, int N,
                   float* keys, float* psc, float* qbase, void* stream);

/* ---- a5-a12, a14: fused persistent decode (attention_model.py:298-356,429-451,510-522,550-584;
 *      state_agh.py:60-174), one CTA per instance for all T steps ----------------------------- */
typedef struct {
  /* inputs */
  const float* keys;       /* [3][KB][n][128], KB = key_mod if key_mod > 0 else B */
  const float* psc;        /* [B][n][128] */
  const float* qbase;      /* [B][128] */
  const uint8_t* coord;    /* [B][n] */
  const float* demand;     /* [B][N] */
  const float* tw_left;    /* [B][n] */
  const double* tw_right;  /* [B][n] */
  const double* duration;  /* [B][n], duration[b][0] = 0 */
  const float* distance;   /* [92*92] */
  const float* travel;     /* [92*92] travel time = distance / 110 as an f32 IEEE division (state_agh.py:118,167),
                              precomputed once so that the mask path loads it instead of dividing every step */
  const float* wblob;      /* for w_cap / w_time */
  /* shared-key replication (sample_many, utils/functions.py:171-188): instance r of B uses the
     keys/psc/qbase/node vectors of instance r % key_mod when key_mod > 0 */
  int key_mod;
  int smem_mats;           /* shared-memory residency: 0 = auto; else 8 | bitmask (1: V, 2: LK', 4: P_sc); the rest is read via L2.
                              Bit 64: never take the shared-key kernel (one warp per replica, used for key_mod > 0 sampling) */
  int mode;                /* AGH_GREEDY | AGH_SAMPLING | AGH_REPLAY */
  float temp;              /* softmax temperature (attention_model.py:447) */
  uint64_t seed;           /* Philox key; counter = (instance id, step, node) */
  int64_t id_offset;       /* global id of instance 0 (sharding-independent sampling) */
  const int16_t* actions;  /* REPLAY: [B][t_stride] */
  int replay_steps;        /* REPLAY: steps to run for every instance */
  const float* inject_logp;/* optional [B][t_stride][n]: select from these log-probs instead */
  const float* inject_z;   /* optional [B][t_stride][n]: use these logits (after the tanh clipping, before the temperature
                              and the mask) instead of the computed ones; everything downstream is the production path */
  int t_stride;            /* row stride (steps) of pi/actions/logp_sel/dumps; >= max(2N-1, 2); also the step budget */
  /* outputs */
  int16_t* pi;             /* [B][t_stride], zero padded */
  int32_t* steps;          /* [B] steps until the instance finished */
  int32_t* status;         /* optional [B]: 0, or AGH_DECODE_STUCK when the step budget ran out first */
  float* ll;               /* [B] sum of selected log-probs */
  float* cost;             /* [B] closed-tour length */
  float* serve;            /* [B][n] */
  float* logp_sel;         /* optional [B][t_stride] */
  float* logp_full;        /* optional [B][t_stride][n] */
  uint32_t* mask_dump;     /* optional [B][t_stride][ceil(n/32)], bit=1 infeasible */
  float* used_dump;        /* optional [B][t_stride] used capacity before the step */
  double* cft_dump;        /* optional [B][t_stride] cur_free_time before the step */
  /* optional per-step tensors the REINFORCE backward differentiates (attention_model.py:510-584 with num_steps = T):
     the step's query, the glimpse attention weights, the glimpse and the log-sum-exp of the masked logits */
  float* q_dump;           /* [B][t_stride][128] */
  float* attn_dump;        /* [B][t_stride][8][n] */
  float* o_dump;           /* [B][t_stride][128] */
  float* lse_dump;         /* [B][t_stride]: log_p_j = logit_j / temp - lse */
  int B, N;
} agh_decode_args;

int agh_decode(const agh_decode_args* args, void* stream);
/* ABI guard for FFI bindings: sizeof(agh_decode_args) as compiled */
size_t agh_sizeof_decode_args(void);

/* ---- f3: instance generation on the device (problem_agh.py:93-116, generate_data.py:8-18): the reference's
 *      distributions from a Philox4x32-10 stream keyed by (seed, instance id, flight) -- not the numpy stream.
 *   arrival_cdf[24]: cumulative arrival-hour probabilities.  demand is [B][fleets][N]. */
int agh_generate(uint64_t seed, long long id_offset, int B, int N, int fleets, float capacity, const float* arrival_cdf,
                 int64_t* loc, float* arrival, float* departure, int64_t* type, float* demand, void* stream);

/* ---- a13: tour validation + cost (problem_agh.py:18-66) ------------------------------------
 *   pi[B][T] i64 or i16.  cost[B] f32, status[B] i32 (0 = valid; AGH_TOUR_INVALID | ... otherwise).
 */
int agh_get_costs(const void* pi, int pi_elem_bytes /* 8: int64 (reference dtype), 2: int16 (agh_decode output) */, int T, const uint8_t* coord, const float* demand, const float* tw_left,
                  const double* tw_right, const double* duration, const float* distance, int B, int N,
                  float* cost, int32_t* status, void* stream);

/* ---- a8 / a12 as single-step batched ops for the StateAGH API (state_agh.py:99-174) --------- */
typedef struct {
  const uint8_t* coord; const float* demand; const float* tw_left; const double* tw_right;
  const double* duration; const float* distance;
  int64_t* prev_a;        /* [B] */
  float* used_capacity;   /* [B] */
  uint8_t* visited;       /* [B][n] */
  float* lengths;         /* [B] */
  double* cur_free_time;  /* [B] */
  float* serve_time;      /* [B][n] */
  int B, N;
} agh_state;

size_t agh_sizeof_state(void);
int agh_state_get_mask(const agh_state* st, uint8_t* mask /*[B][n], 1 = infeasible*/, void* stream);
int agh_state_update(const agh_state* st, const int64_t* selected /*[B]*/, void* stream);

/* ---- a18: REINFORCE training primitives (train.py:159-219; graph_encoder.py:56-172 and attention_model.py:392-451,
 *      510-584 differentiated).  agh_b200/nets/train_fn.py composes them into the forward / backward of one fleet. ---- */
/* generic fp32 GEMM  C = alpha op(A) op(B) [+ bias[n]] [relu] [zeroed where mask <= 0] [+ C]
 *   op(A) is M x K: element (m,k) = transA ? A[k*lda + m] : A[m*lda + k];   op(B) is K x N: (k,n) = transB ? B[n*ldb + k] : B[k*ldb + n]
 *   batched over (b1, b2) with independent element strides; splits > 1 = split-K with atomic accumulation into C
 *   (C must hold the value to accumulate onto; bias / relu / mask are not allowed then) */
typedef struct {
  const float* A; long long lda; int transA;
  const float* B; long long ldb; int transB;
  float* C; long long ldc;
  int M, N, K;
  int batch1, batch2;
  long long sA1, sA2, sB1, sB2, sC1, sC2;
  float alpha;
  int accumulate;
  const float* bias;
  int relu;
  const float* mask; long long ldm;
  int splits;
} agh_gemm_desc;
size_t agh_sizeof_gemm_desc(void);
int agh_gemm(const agh_gemm_desc* desc, void* stream);
/* float64 product for the weight packing (nets/weights.py: project_out folded into the logit keys) */
int agh_gemm_f64(const double* A, int lda, int transA, const double* B, int ldb, int transB, double* C, int ldc,
                 int M, int N, int K, double alpha, void* stream);
/* out[n] += sum_m X[m*ld + n] (bias gradients) */
int agh_colsum(const float* X, long long ld, int M, int N, float* out, void* stream);
/* BatchNorm1d(128) in training mode over M rows (graph_encoder.py:137-138): y, batch mean / rstd (for the backward), running
 * statistics updated in place when non-NULL; scratch: 256 doubles */
int agh_bn_train_fwd(const float* x, const float* gamma, const float* beta, int M, float eps, float momentum, float* y,
                     float* mean, float* rstd, float* running_mean, float* running_var, double* scratch, void* stream);
int agh_bn_train_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int M,
                     float* dx, float* dgamma, float* dbeta, double* scratch, void* stream);
/* encoder self-attention (graph_encoder.py:83-104) on row-major qkv [B*n][384] -> heads [B*n][128], and its backward */
int agh_mha_fwd(const float* qkv, int B, int N, float* heads, void* stream);
int agh_mha_bwd(const float* qkv, const float* heads, const float* dheads, int B, int N, float* dqkv, void* stream);
/* decoder backward, element-wise pieces: logit gradient from the recorded log-probs; dS' of the base-2 glimpse softmax
 * (in place on dattn); fan-out of the query gradient to qbase / P_sc rows / capacity and time columns */
int agh_tf_dlogits(const float* logp_full, const float* lse, const int16_t* pi, const int32_t* steps, const float* grad_ll,
                   float temp, int B, int T, int N, float* dU, void* stream);
int agh_tf_dscore(const float* attn, float* dattn, long long rows, int N, void* stream);
int agh_tf_dquery(const float* dQ, const int16_t* pi, const int32_t* steps, const float* used, const double* cft, int B,
                  int T, int N, float* dqbase, float* dpsc, float* dwcap, float* dwtime, void* stream);

/* ---- a16: best-of-R reduction of sample_many (utils/functions.py:190-223) -------------------
 *   costs[R*B] (replica-major: row r*B+b), pi[R*B][t_stride] i16, serve[R*B][n]
 *   -> best_cost[B], best_pi[B][t_stride], best_serve[B][n], best_rep[B]
 */
int agh_best_of(const float* costs, const int16_t* pi, const float* serve, int R, int B, int N, int t_stride,
                float* best_cost, int16_t* best_pi, float* best_serve, int32_t* best_rep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AGH_B200_H */
