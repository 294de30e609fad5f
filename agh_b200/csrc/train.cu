// REINFORCE training path (train.py:159-219 with the model code it differentiates: graph_encoder.py:56-172,
// attention_model.py:392-451,510-584): the hand-written CUDA primitives behind agh_b200/nets/train_fn.py.
//
// The forward pass of a training batch is the sampling rollout itself (agh_decode on train-mode embeddings); its
// log-likelihood is what REINFORCE differentiates.  Nothing here calls cuBLAS / cuDNN / ATen contractions:
//   * agh_gemm           one generic fp32 GEMM  C = alpha op(A) op(B) [+ bias] [relu] [* (mask > 0)] [+ C], batched over two
//                        indices with independent strides, optional split-K (atomic accumulation) for the long
//                        reductions of the weight gradients (dW = X^T dY with K = B*n rows)
//   * agh_colsum         bias gradients
//   * agh_bn_train_fwd / agh_bn_train_bwd   BatchNorm1d with batch statistics over the B*n rows (graph_encoder.py:137-138),
//                        running-statistics update included (momentum 0.1, unbiased variance)
//   * agh_mha_bwd        backward of the encoder self-attention (forward = the eval kernel, agh_mha_fwd)
//   * agh_tf_dlogits / agh_tf_dscore / agh_tf_dquery   element-wise pieces of the decoder backward; its contractions
//                        ([T,n]x[n,128] per instance and head) are batched agh_gemm calls on the per-step tensors the
//                        rollout kernel recorded (query, attention weights, glimpse, log-probs)
//   * agh_gemm_f64       the one float64 product of the weight packing (project_out folded into the logit keys)
// The batch of a training step is small (64 instances = 6,464 rows at N = 100), so these are CUDA-core kernels sized for
// latency, not the tcgen05 pipeline of the 10^6-row evaluation GEMMs.
#include "common.cuh"
#include "gemm.cuh"
#include <math_constants.h>

namespace agh {

// ------------------------------------------------------------------------------------------
// generic GEMM: 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread
// ------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16;

__global__ void __launch_bounds__(256) gemm_generic_kernel(const agh_gemm_desc g) {
  __shared__ __align__(16) float As[GK][GT + 4];
  __shared__ __align__(16) float Bs[GK][GT + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
  const int splits = g.splits > 1 ? g.splits : 1;
  const int bz = blockIdx.z / splits, split = blockIdx.z % splits;
  const int b1 = bz / g.batch2, b2 = bz % g.batch2;
  const float* A = g.A + (size_t)b1 * g.sA1 + (size_t)b2 * g.sA2;
  const float* B = g.B + (size_t)b1 * g.sB1 + (size_t)b2 * g.sB2;
  float* C = g.C + (size_t)b1 * g.sC1 + (size_t)b2 * g.sC2;
  int kchunk = (g.K + splits - 1) / splits;
  kchunk = (kchunk + GK - 1) / GK * GK;
  const int kbeg = split * kchunk, kend = min(g.K, kbeg + kchunk);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += GK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      {   // A tile: element (m, k) of op(A)
        const int m = g.transA ? (idx & 63) : (idx >> 4), k = g.transA ? (idx >> 6) : (idx & 15);
        float v = 0.f;
        if (m0 + m < g.M && k0 + k < kend)
          v = g.transA ? A[(size_t)(k0 + k) * g.lda + m0 + m] : A[(size_t)(m0 + m) * g.lda + k0 + k];
        As[k][m] = v;
      }
      {   // B tile: element (k, n) of op(B)
        const int n = g.transB ? (idx >> 4) : (idx & 63), k = g.transB ? (idx & 15) : (idx >> 6);
        float v = 0.f;
        if (n0 + n < g.N && k0 + k < kend)
          v = g.transB ? B[(size_t)(n0 + n) * g.ldb + k0 + k] : B[(size_t)(k0 + k) * g.ldb + n0 + n];
        Bs[k][n] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      float* c = C + (size_t)m * g.ldc + n;
      if (splits > 1) {
        atomicAdd(c, v);
      } else {
        if (g.bias != nullptr) v += g.bias[n];
        if (g.relu) v = fmaxf(v, 0.f);
        if (g.mask != nullptr && !(g.mask[(size_t)m * g.ldm + n] > 0.f)) v = 0.f;
        if (g.accumulate) v += *c;
        *c = v;
      }
    }
  }
}

// column sums of X [M][N] (bias gradients): out[n] += sum_m X[m][n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, long long ld, int M, int N, int rows_per_block,
                                                     float* __restrict__ out) {
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int c = threadIdx.x; c < N; c += 256) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += X[(size_t)r * ld + c];
    atomicAdd(out + c, s);
  }
}

// naive float64 product (weight packing): C[M][N] = alpha * A^T?[..] ...; one thread per output element
__global__ void gemm_f64_kernel(const double* __restrict__ A, int lda, int transA, const double* __restrict__ B, int ldb,
                                int transB, double* __restrict__ C, int ldc, int M, int N, int K, double alpha) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  double s = 0.0;
  for (int k = 0; k < K; ++k) {
    const double a = transA ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
    const double b = transB ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n];
    s = fma(a, b, s);
  }
  C[(size_t)m * ldc + n] = alpha * s;
}

// ------------------------------------------------------------------------------------------
// BatchNorm1d, training mode (graph_encoder.py:135-138; torch.nn.BatchNorm1d: biased variance normalises,
// unbiased variance updates running_var, momentum 0.1)
// ------------------------------------------------------------------------------------------
// sums[0..127] += sum_r x[r][c] ; sums[128..255] += sum_r x[r][c]^2     (float64 accumulators)
__global__ void __launch_bounds__(128) bn_stats_kernel(const float* __restrict__ x, int M, int rows_per_block, double* __restrict__ sums) {
  const int c = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  double s = 0.0, q = 0.0;
  for (int r = r0; r < r1; ++r) {
    const double v = (double)x[(size_t)r * D + c];
    s += v;
    q = fma(v, v, q);
  }
  atomicAdd(sums + c, s);
  atomicAdd(sums + D + c, q);
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, int M,
                                                       float eps, float momentum, float* __restrict__ y, float* __restrict__ mean_out,
                                                       float* __restrict__ rstd_out, float* __restrict__ run_mean,
                                                       float* __restrict__ run_var) {
  __shared__ float s_scale[D], s_shift[D];
  if (threadIdx.x < D) {
    const int c = threadIdx.x;
    const double mean = sums[c] / M;
    const double var = fmax(sums[D + c] / M - mean * mean, 0.0);
    const double rstd = 1.0 / sqrt(var + (double)eps);
    s_scale[c] = (float)(rstd * (double)gamma[c]);
    s_shift[c] = (float)((double)beta[c] - mean * rstd * (double)gamma[c]);
    if (blockIdx.x == 0) {
      mean_out[c] = (float)mean;
      rstd_out[c] = (float)rstd;
      if (run_mean != nullptr) {
        const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        run_mean[c] = (float)((1.0 - (double)momentum) * (double)run_mean[c] + (double)momentum * mean);
        run_var[c] = (float)((1.0 - (double)momentum) * (double)run_var[c] + (double)momentum * unbiased);
      }
    }
  }
  __syncthreads();
  const size_t total = (size_t)M * (D / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (D / 4)) * 4;
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x = fmaf(v.x, s_scale[c4], s_shift[c4]); v.y = fmaf(v.y, s_scale[c4 + 1], s_shift[c4 + 1]);
    v.z = fmaf(v.z, s_scale[c4 + 2], s_shift[c4 + 2]); v.w = fmaf(v.w, s_scale[c4 + 3], s_shift[c4 + 3]);
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

// sums[0..127] += sum_r dy ; sums[128..255] += sum_r dy * xhat
__global__ void __launch_bounds__(128) bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd, int M,
                                                           int rows_per_block, double* __restrict__ sums) {
  const int c = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  const float mu = mean[c], rs = rstd[c];
  double s = 0.0, q = 0.0;
  for (int r = r0; r < r1; ++r) {
    const float g = dy[(size_t)r * D + c];
    s += (double)g;
    q += (double)(g * ((x[(size_t)r * D + c] - mu) * rs));
  }
  atomicAdd(sums + c, s);
  atomicAdd(sums + D + c, q);
}

// dx = gamma * rstd * (dy - dbeta / M - xhat * dgamma / M); dgamma / dbeta are written by block 0
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const double* __restrict__ sums, int M,
                                                           float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float s_mu[D], s_rs[D], s_g[D], s_db[D], s_dg[D];
  if (threadIdx.x < D) {
    const int c = threadIdx.x;
    s_mu[c] = mean[c]; s_rs[c] = rstd[c]; s_g[c] = gamma[c] * rstd[c];
    s_db[c] = (float)(sums[c] / M);
    s_dg[c] = (float)(sums[D + c] / M);
    if (blockIdx.x == 0) {
      dbeta[c] = (float)sums[c];
      dgamma[c] = (float)sums[D + c];
    }
  }
  __syncthreads();
  const size_t total = (size_t)M * (D / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (D / 4)) * 4;
    const float4 g = reinterpret_cast<const float4*>(dy)[i];
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 o;
    o.x = s_g[c4] * (g.x - s_db[c4] - (v.x - s_mu[c4]) * s_rs[c4] * s_dg[c4]);
    o.y = s_g[c4 + 1] * (g.y - s_db[c4 + 1] - (v.y - s_mu[c4 + 1]) * s_rs[c4 + 1] * s_dg[c4 + 1]);
    o.z = s_g[c4 + 2] * (g.z - s_db[c4 + 2] - (v.z - s_mu[c4 + 2]) * s_rs[c4 + 2] * s_dg[c4 + 2]);
    o.w = s_g[c4 + 3] * (g.w - s_db[c4 + 3] - (v.w - s_mu[c4 + 3]) * s_rs[c4 + 3] * s_dg[c4 + 3]);
    reinterpret_cast<float4*>(dx)[i] = o;
  }
}

// ------------------------------------------------------------------------------------------
// backward of the encoder self-attention (graph_encoder.py:83-104): one CTA per (instance, head)
//   S = Q K^T / 4, A = softmax(S), O = A V;  given dO:  dV = A^T dO,  dA = dO V^T,
//   dS = A o (dA - rowsum(A o dA)),  dQ = dS K / 4,  dK = dS^T Q / 4;   rowsum(A o dA) = dO . O
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mha_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ heads,
                                                      const float* __restrict__ dheads, int n, float* __restrict__ dqkv) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x;
  float* sQ = sm;                       // [n][16]
  float* sK = sQ + (size_t)n * HD;
  float* sV = sK + (size_t)n * HD;
  float* sdO = sV + (size_t)n * HD;
  float* sMx = sdO + (size_t)n * HD;     // row max
  float* sL = sMx + n;                  // row sum
  float* sDl = sL + n;                  // dO_i . O_i
  const size_t row0 = (size_t)b * n;
  for (int idx = tid; idx < n * 4; idx += 128) {
    const int j = idx >> 2, c = idx & 3;
    const float* q = qkv + (row0 + j) * 3 * D + h * HD + c * 4;
    reinterpret_cast<float4*>(sQ)[idx] = *reinterpret_cast<const float4*>(q);
    reinterpret_cast<float4*>(sK)[idx] = *reinterpret_cast<const float4*>(q + D);
    reinterpret_cast<float4*>(sV)[idx] = *reinterpret_cast<const float4*>(q + 2 * D);
    reinterpret_cast<float4*>(sdO)[idx] = *reinterpret_cast<const float4*>(dheads + (row0 + j) * D + h * HD + c * 4);
  }
  __syncthreads();
  // pass 1: thread = query row i: softmax statistics, D_i, dQ_i
  for (int i = tid; i < n; i += 128) {
    float q[HD], g[HD];
#pragma unroll
    for (int k = 0; k < HD; ++k) { q[k] = sQ[i * HD + k]; g[k] = sdO[i * HD + k]; }
    float mx = -CUDART_INF_F;
    for (int j = 0; j < n; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < HD; ++k) s = fmaf(q[k], sK[j * HD + k], s);
      mx = fmaxf(mx, 0.25f * s);
    }
    float l = 0.f;
    for (int j = 0; j < n; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < HD; ++k) s = fmaf(q[k], sK[j * HD + k], s);
      l += expf(0.25f * s - mx);
    }
    float dl = 0.f;
    {
      const float* o = heads + (row0 + i) * D + h * HD;
#pragma unroll
      for (int k = 0; k < HD; ++k) dl = fmaf(g[k], o[k], dl);
    }
    sMx[i] = mx; sL[i] = l; sDl[i] = dl;
    float dq[HD];
#pragma unroll
    for (int k = 0; k < HD; ++k) dq[k] = 0.f;
    const float inv_l = 1.0f / l;
    for (int j = 0; j < n; ++j) {
      float s = 0.f, da = 0.f;
#pragma unroll
      for (int k = 0; k < HD; ++k) { s = fmaf(q[k], sK[j * HD + k], s); da = fmaf(g[k], sV[j * HD + k], da); }
      const float a = expf(0.25f * s - mx) * inv_l;
      const float ds = 0.25f * a * (da - dl);
#pragma unroll
      for (int k = 0; k < HD; ++k) dq[k] = fmaf(ds, sK[j * HD + k], dq[k]);
    }
    float* out = dqkv + (row0 + i) * 3 * D + h * HD;
#pragma unroll
    for (int k = 0; k < HD; ++k) out[k] = dq[k];
  }
  __syncthreads();
  // pass 2: thread = key row j: dK_j, dV_j
  for (int j = tid; j < n; j += 128) {
    float kk[HD], vv[HD], dk[HD], dv[HD];
#pragma unroll
    for (int k = 0; k < HD; ++k) { kk[k] = sK[j * HD + k]; vv[k] = sV[j * HD + k]; dk[k] = 0.f; dv[k] = 0.f; }
    for (int i = 0; i < n; ++i) {
      float s = 0.f, da = 0.f;
#pragma unroll
      for (int k = 0; k < HD; ++k) { s = fmaf(sQ[i * HD + k], kk[k], s); da = fmaf(sdO[i * HD + k], vv[k], da); }
      const float a = expf(0.25f * s - sMx[i]) / sL[i];
      const float ds = 0.25f * a * (da - sDl[i]);
#pragma unroll
      for (int k = 0; k < HD; ++k) { dv[k] = fmaf(a, sdO[i * HD + k], dv[k]); dk[k] = fmaf(ds, sQ[i * HD + k], dk[k]); }
    }
    float* out = dqkv + (row0 + j) * 3 * D + D + h * HD;
#pragma unroll
    for (int k = 0; k < HD; ++k) { out[k] = dk[k]; out[D + k] = dv[k]; }
  }
}

// ------------------------------------------------------------------------------------------
// decoder backward, element-wise pieces (attention_model.py:446-447,550-584 differentiated)
// ------------------------------------------------------------------------------------------
// dU[b][t][j] = g_b * (1[j == pi_t] - p_j) * (10 / temp) * (1 - tanh(u_j)^2)   for t < steps[b], 0 otherwise
//   p_j = exp(log_p_j), z_j = log_p_j + lse_t = 10 tanh(u_j) / temp
__global__ void tf_dlogits_kernel(const float* __restrict__ logp, const float* __restrict__ lse, const int16_t* __restrict__ pi,
                                  const int32_t* __restrict__ steps, const float* __restrict__ gll, float temp, int B, int T,
                                  int n, float* __restrict__ dU) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * T * n) return;
  const int j = (int)(idx % n);
  const size_t bt = idx / n;
  const int t = (int)(bt % T), b = (int)(bt / T);
  float out = 0.f;
  if (t < steps[b]) {
    const float lp = logp[idx];
    if (lp != -CUDART_INF_F) {
      const float z = lp + lse[bt];
      const float th = z * temp * 0.1f;
      const float dz = gll[b] * ((j == (int)pi[bt] ? 1.0f : 0.0f) - expf(lp));
      out = dz * (10.0f / temp) * (1.0f - th * th);
    }
  }
  dU[idx] = out;
}

// rows (b, t, h) of the attention weights A [rows][n] and of dA (in place -> dS'):
//   the kernel's softmax runs in base 2 on s' = q . K' (K' carries log2(e) / 4), so dS' = ln2 * A o (dA - sum_j A dA)
__global__ void tf_dscore_kernel(const float* __restrict__ A, float* __restrict__ dA, size_t rows, int n) {
  const size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* a = A + row * n;
  float* d = dA + row * n;
  float dot = 0.f;
  for (int j = lane; j < n; j += 32) dot = fmaf(a[j], d[j], dot);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
  for (int j = lane; j < n; j += 32) d[j] = 0.6931471805599453f * a[j] * (d[j] - dot);
}

// query gradient fan-out: q_t = qbase + P_sc[prev_t] + w_cap (1 - used_t) + w_time (float)(cft_t / 1440)
__global__ void __launch_bounds__(128) tf_dquery_kernel(const float* __restrict__ dQ, const int16_t* __restrict__ pi,
                                                        const int32_t* __restrict__ steps, const float* __restrict__ used,
                                                        const double* __restrict__ cft, int T, int n, float* __restrict__ dqbase,
                                                        float* __restrict__ dpsc, float* __restrict__ dwcap,
                                                        float* __restrict__ dwtime) {
  const int b = blockIdx.x, c = threadIdx.x;
  const int Tb = steps[b];
  float aq = 0.f, ac = 0.f, at = 0.f;
  int prev = 0;
  for (int t = 0; t < Tb; ++t) {
    const size_t bt = (size_t)b * T + t;
    const float v = dQ[bt * D + c];
    aq += v;
    ac = fmaf(1.0f - used[bt], v, ac);
    at = fmaf((float)(cft[bt] / 1440.0), v, at);
    dpsc[((size_t)b * n + prev) * D + c] += v;      // thread c owns column c of this instance: no atomics
    prev = (int)pi[bt];
  }
  dqbase[(size_t)b * D + c] = aq;
  atomicAdd(dwcap + c, ac);
  atomicAdd(dwtime + c, at);
}


// ------------------------------------------------------------------------------------------
// Weight gradients: C[I][J] += sum_k A[k][i] B[k][j] with K = the number of rows of the batch (tens of thousands) and I, J <= 512.
// Both operands are read along their rows (coalesced float4), 128 x 128 tile, 8 x 8 outputs per thread, double-buffered shared
// memory, the K range split over enough CTAs for two per SM; partial tiles are added to C with vector atomics (C is zeroed by the
// caller, as for the split-K form of the generic kernel).
// ------------------------------------------------------------------------------------------
constexpr int TN_T = 128, TN_K = 16;
__global__ void __launch_bounds__(256, 2) sgemm_tn_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B,
                                                          long long ldb, float* __restrict__ C, long long ldc, int I, int J, int K,
                                                          int kchunk) {
  __shared__ __align__(16) float As[2][TN_K][TN_T];
  __shared__ __align__(16) float Bs[2][TN_K][TN_T];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * TN_T, j0 = blockIdx.y * TN_T;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  if (kbeg >= kend) return;
  const int lrow = tid >> 5, lcol = (tid & 31) * 4;            // this thread's float4 of a [16][128] tile: rows lrow, lrow + 8
  float4 ra[2], rb[2];
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = k0 + lrow + 8 * r;
      ra[r] = (k < kend && i0 + lcol < I) ? *reinterpret_cast<const float4*>(A + (size_t)k * lda + i0 + lcol) : z4;
      rb[r] = (k < kend && j0 + lcol < J) ? *reinterpret_cast<const float4*>(B + (size_t)k * ldb + j0 + lcol) : z4;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      *reinterpret_cast<float4*>(&As[buf][lrow + 8 * r][lcol]) = ra[r];
      *reinterpret_cast<float4*>(&Bs[buf][lrow + 8 * r][lcol]) = rb[r];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  load_tiles(kbeg);
  store_tiles(0);
  __syncthreads();
  const int nk = (kend - kbeg + TN_K - 1) / TN_K;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kbeg + (kt + 1) * TN_K);
#pragma unroll
    for (int kk = 0; kk < TN_K; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= I) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = j0 + jh * 64 + tx * 4;
      if (col >= J) continue;
      atomicAdd(reinterpret_cast<float4*>(C + (size_t)row * ldc + col),
                make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]));
    }
  }
}

}  // namespace agh

using namespace agh;

#ifndef AGH_TC_TRAIN_MIN_ROWS
#define AGH_TC_TRAIN_MIN_ROWS 4096
#endif

extern "C" int agh_gemm(const agh_gemm_desc* desc, void* stream) {
  AGH_CHECK_ARG(desc != nullptr);
  agh_gemm_desc g = *desc;
  AGH_CHECK_ARG(g.A && g.B && g.C && g.M >= 0 && g.N >= 0 && g.K >= 0);
  if (g.batch1 <= 0) g.batch1 = 1;
  if (g.batch2 <= 0) g.batch2 = 1;
  if (g.splits <= 0) g.splits = 1;
  AGH_CHECK_ARG(g.splits == 1 || (g.bias == nullptr && !g.relu && g.mask == nullptr));
  if (g.M == 0 || g.N == 0) return AGH_OK;
  // fast path: plain row-major A times a row-major [K][N] matrix with tile-aligned sizes (every activation-sized GEMM of
  // the training step, forward and backward) goes through the 128 x 128 x 16 register-tiled SGEMM of the encoder
  if (!g.transA && !g.transB && g.batch1 * g.batch2 == 1 && g.splits == 1 && g.alpha == 1.0f && g.M >= 256 && g.K % 16 == 0 &&
      g.N % 128 == 0 && g.ldb == g.N && g.lda % 4 == 0 && g.ldc % 4 == 0 && (g.mask == nullptr || g.ldm % 4 == 0) &&
      ((uintptr_t)g.A | (uintptr_t)g.B | (uintptr_t)g.C | (uintptr_t)g.bias | (uintptr_t)g.mask) % 16 == 0) {
    GemmArgs a{};
    a.A = g.A; a.lda = (int)g.lda; a.W = g.B; a.C = g.C; a.ldc = (int)g.ldc; a.M = g.M; a.K = g.K; a.Nc = g.N;
    a.bias = g.bias; a.relu = g.relu; a.accumulate = g.accumulate; a.mask = g.mask; a.ldm = (int)g.ldm;
    return launch_sgemm_train(a, (cudaStream_t)stream);
  }
  // Activation-sized products against a K-major weight matrix (op(B) = B^T with B = [N][K], the layout the tcgen05 GEMM reads):
  // on the tensor cores from AGH_TC_TRAIN_MIN_ROWS rows (the two-term FP16 split keeps fp32-level accuracy).  Measured
  // (tools/bench_train_gemm.py): 15 us against 31-39 us on the CUDA cores at 6,464 rows (the reference's batch of 64 at
  // AGH-100; equal at K = N = 128), 23-66 against 46-176 us at 51,712 rows; the persistent kernel has ~13 us of fixed cost,
  // which is the CUDA-core SGEMM's whole run time around 2,000 rows
  if (!g.transA && g.transB && g.batch1 * g.batch2 == 1 && g.splits == 1 && g.alpha == 1.0f && g.M >= AGH_TC_TRAIN_MIN_ROWS && gemm_tc_enabled() &&
      !g.accumulate && g.mask == nullptr &&
      (g.K == 128 || g.K % 64 == 0) && g.N % 128 == 0 && g.ldb == g.K && g.lda % 4 == 0 && g.ldc % 4 == 0 &&
      ((uintptr_t)g.A | (uintptr_t)g.B) % 16 == 0) {
    GemmArgs a{};
    a.A = g.A; a.lda = (int)g.lda; a.Wt = g.B; a.C = g.C; a.ldc = (int)g.ldc; a.M = g.M; a.K = g.K; a.Nc = g.N;
    a.bias = g.bias; a.relu = g.relu;
    return launch_gemm_tc<EPI_TRAIN>(a, (cudaStream_t)stream);
  }
  // weight gradients (op(A) = A^T with a long reduction, C zeroed by the caller and filled by split-K atomics)
  if (g.transA && !g.transB && g.batch1 * g.batch2 == 1 && g.splits > 1 && g.alpha == 1.0f && g.K >= 2048 && g.M >= 64 &&
      g.M % 4 == 0 && g.N % 4 == 0 && g.lda % 4 == 0 && g.ldb % 4 == 0 && g.ldc % 4 == 0 &&
      ((uintptr_t)g.A | (uintptr_t)g.B | (uintptr_t)g.C) % 16 == 0) {
    const int tiles = ceil_div(g.M, TN_T) * ceil_div(g.N, TN_T);
    int sm = 148;
    {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    }
    int S = ceil_div(2 * sm, tiles);                          // two CTAs per SM
    S = S < 1 ? 1 : (S > g.K / 128 ? g.K / 128 : S);
    int kchunk = ceil_div(ceil_div(g.K, S), TN_K) * TN_K;
    S = ceil_div(g.K, kchunk);
    dim3 grid(ceil_div(g.M, TN_T), ceil_div(g.N, TN_T), (unsigned)S);
    sgemm_tn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g.A, g.lda, g.B, g.ldb, g.C, g.ldc, g.M, g.N, g.K, kchunk);
    AGH_LAUNCH_CHECK();
    return AGH_OK;
  }
  const long long z = (long long)g.batch1 * g.batch2 * g.splits;
  AGH_CHECK_ARG(z <= 65535);
  dim3 grid(ceil_div(g.M, GT), ceil_div(g.N, GT), (unsigned)z);
  gemm_generic_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_gemm_f64(const double* A, int lda, int transA, const double* B, int ldb, int transB, double* C, int ldc,
                            int M, int N, int K, double alpha, void* stream) {
  AGH_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0);
  dim3 grid(ceil_div(N, 128), M);
  gemm_f64_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, alpha);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_colsum(const float* X, long long ld, int M, int N, float* out, void* stream) {
  AGH_CHECK_ARG(X && out && M >= 0 && N > 0);
  if (M == 0) return AGH_OK;
  const int rpb = 16;      // short serial loops, several hundred CTAs: the kernel is latency-bound, not bandwidth-bound
  colsum_kernel<<<ceil_div(M, rpb), 256, 0, (cudaStream_t)stream>>>(X, ld, M, N, rpb, out);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_bn_train_fwd(const float* x, const float* gamma, const float* beta, int M, float eps, float momentum,
                                float* y, float* mean, float* rstd, float* running_mean, float* running_var,
                                double* scratch, void* stream) {
  AGH_CHECK_ARG(x && gamma && beta && y && mean && rstd && scratch && M > 0);
  cudaStream_t st = (cudaStream_t)stream;
  AGH_CUDA(cudaMemsetAsync(scratch, 0, 2 * D * sizeof(double), st));
  const int rpb = 32;
  bn_stats_kernel<<<ceil_div(M, rpb), 128, 0, st>>>(x, M, rpb, scratch);
  AGH_LAUNCH_CHECK();
  const int blocks = min(ceil_div(M * (D / 4), 256), 592);
  bn_apply_kernel<<<blocks, 256, 0, st>>>(x, scratch, gamma, beta, M, eps, momentum, y, mean, rstd, running_mean, running_var);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_bn_train_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                int M, float* dx, float* dgamma, float* dbeta, double* scratch, void* stream) {
  AGH_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && scratch && M > 0);
  cudaStream_t st = (cudaStream_t)stream;
  AGH_CUDA(cudaMemsetAsync(scratch, 0, 2 * D * sizeof(double), st));
  const int rpb = 32;
  bn_bwd_stats_kernel<<<ceil_div(M, rpb), 128, 0, st>>>(dy, x, mean, rstd, M, rpb, scratch);
  AGH_LAUNCH_CHECK();
  const int blocks = min(ceil_div(M * (D / 4), 256), 592);
  bn_bwd_apply_kernel<<<blocks, 256, 0, st>>>(dy, x, mean, rstd, gamma, scratch, M, dx, dgamma, dbeta);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_mha_bwd(const float* qkv, const float* heads, const float* dheads, int B, int N, float* dqkv, void* stream) {
  AGH_CHECK_ARG(qkv && heads && dheads && dqkv && B >= 0 && N >= 1 && N < AGH_MAX_N);
  if (B == 0) return AGH_OK;
  const int n = N + 1;
  const size_t smem = ((size_t)n * HD * 4 + 3 * (size_t)n) * sizeof(float);
  if (smem > 48 * 1024) AGH_CUDA(cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mha_bwd_kernel<<<B * H, 128, smem, (cudaStream_t)stream>>>(qkv, heads, dheads, n, dqkv);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_tf_dlogits(const float* logp_full, const float* lse, const int16_t* pi, const int32_t* steps,
                              const float* grad_ll, float temp, int B, int T, int N, float* dU, void* stream) {
  AGH_CHECK_ARG(logp_full && lse && pi && steps && grad_ll && dU && B >= 0 && T > 0 && N >= 1 && temp > 0.f);
  if (B == 0) return AGH_OK;
  const size_t total = (size_t)B * T * (N + 1);
  tf_dlogits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logp_full, lse, pi, steps, grad_ll, temp,
                                                                                     B, T, N + 1, dU);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_tf_dscore(const float* attn, float* dattn, long long rows, int N, void* stream) {
  AGH_CHECK_ARG(attn && dattn && rows >= 0 && N >= 1);
  if (rows == 0) return AGH_OK;
  const long long blocks = (rows * 32 + 255) / 256;
  tf_dscore_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(attn, dattn, (size_t)rows, N + 1);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_tf_dquery(const float* dQ, const int16_t* pi, const int32_t* steps, const float* used, const double* cft,
                             int B, int T, int N, float* dqbase, float* dpsc, float* dwcap, float* dwtime, void* stream) {
  AGH_CHECK_ARG(dQ && pi && steps && used && cft && dqbase && dpsc && dwcap && dwtime && B >= 0 && T > 0 && N >= 1);
  if (B == 0) return AGH_OK;
  tf_dquery_kernel<<<B, D, 0, (cudaStream_t)stream>>>(dQ, pi, steps, used, cft, T, N + 1, dqbase, dpsc, dwcap, dwtime);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}
