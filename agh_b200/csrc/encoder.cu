// Graph-attention encoder (eval-mode BatchNorm) and decoder precompute: the layer driver, and the fp32 CUDA-core
// kernels (SGEMM, attention) that the tensor-core path (gemm_tc.cu, mha_mma.cu) is validated against.
//
//   agh_init_embed : attention_model.py:272-294
//   agh_encode     : graph_encoder.py:56-111 (MHA), :135-143 (BatchNorm1d, running stats),
//                    :146-172 (layer = MHA+skip, BN, FF+skip, BN), :195-207
//   agh_precompute : attention_model.py:392-414, :604-611
//
// CUDA-core back end (agh_set_gemm_backend(0) / AGH_GEMM=simt): all dense contractions go through one tiled SGEMM
// (128x128x16 tiles, 8x8 register micro-tile, register double buffering) whose epilogue fuses bias / ReLU /
// residual + BatchNorm scale-shift / the four [M][128] outputs of the decoder projection (K', V, LK', P_sc);
// self-attention is a flash-style kernel: two heads per CTA, K_h/V_h staged in shared memory, two query rows per
// thread with a base-2 online softmax, so no [n,n] score matrix ever reaches HBM.  The default back end replaces
// both with tensor-core kernels behind the same launch sequence.
#include "common.cuh"
#include "gemm.cuh"
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

namespace agh {

// ------------------------------------------------------------------------------------------
// init embedding
// ------------------------------------------------------------------------------------------
// One warp writes 32 consecutive rows: lane r first computes the three node features of row r (two of them divisions, one in
// f64 -- once per row instead of once per output element), then the warp walks the rows with the features broadcast and every
// lane producing 4 of the 128 columns; the lane's slices of the feature weights and bias stay in registers.
__global__ void __launch_bounds__(256) init_embed_kernel(const float* __restrict__ w, const uint8_t* __restrict__ coord,
                                  const float* __restrict__ demand, const float* __restrict__ tw_left,
                                  const double* __restrict__ tw_right, int twr_f32, int B, int N, float* __restrict__ h0) {
  const int n = N + 1;
  const int lane = threadIdx.x & 31;
  const size_t rows = (size_t)B * n;
  const size_t row0 = (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
  if (row0 >= rows) return;
  const size_t row = row0 + lane;
  int loc = 0, node = 0;
  float f0 = 0.f, f1 = 0.f, f2 = 0.f;
  if (row < rows) {
    const int j = (int)(row % n);
    const size_t b = row / n;
    loc = coord[row];
    node = j >= 1;
    if (node) {
      // features are concatenated in f64 when tw_right is f64 and then cast to f32
      // (attention_model.py:290-292): demand and tw_left/1440 are f32 values, tw_right/1440 is an
      // f64 division rounded to f32 (an f32 division when tw_right is f32: fleet 10).
      f0 = demand[b * N + (j - 1)];
      f1 = __fdiv_rn(tw_left[row], 1440.0f);
      f2 = twr_f32 ? __fdiv_rn((float)tw_right[row], 1440.0f) : (float)(tw_right[row] / 1440.0);
    }
  }
  const float4 w0 = reinterpret_cast<const float4*>(w + W_INIT_W)[lane];
  const float4 w1 = reinterpret_cast<const float4*>(w + W_INIT_W + D)[lane];
  const float4 w2 = reinterpret_cast<const float4*>(w + W_INIT_W + 2 * D)[lane];
  const float4 bb = reinterpret_cast<const float4*>(w + W_INIT_B)[lane];
  const int cnt = (int)(rows - row0 < 32 ? rows - row0 : 32);
  for (int i = 0; i < cnt; ++i) {
    const int loc_i = __shfl_sync(0xffffffffu, loc, i), node_i = __shfl_sync(0xffffffffu, node, i);
    const float g0 = __shfl_sync(0xffffffffu, f0, i), g1 = __shfl_sync(0xffffffffu, f1, i), g2 = __shfl_sync(0xffffffffu, f2, i);
    float4 v = reinterpret_cast<const float4*>(w + W_LOC_EMB + (size_t)loc_i * D)[lane];
    if (node_i) {
      v.x += fmaf(w2.x, g2, fmaf(w1.x, g1, fmaf(w0.x, g0, bb.x)));
      v.y += fmaf(w2.y, g2, fmaf(w1.y, g1, fmaf(w0.y, g0, bb.y)));
      v.z += fmaf(w2.z, g2, fmaf(w1.z, g1, fmaf(w0.z, g0, bb.z)));
      v.w += fmaf(w2.w, g2, fmaf(w1.w, g1, fmaf(w0.w, g0, bb.w)));
    }
    reinterpret_cast<float4*>(h0 + (row0 + i) * D)[lane] = v;
  }
}

// ------------------------------------------------------------------------------------------
// SGEMM  C[M,Nc] = epi(A[M,K] @ W[K,Nc])
// ------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16;

// BMT: rows per CTA tile: 128 (8 x 8 outputs per thread; the 10^6-row evaluation GEMMs) or 64 (4 x 8; the 6,464-row training
// GEMMs with 128 output columns, which would otherwise launch 51 CTAs on 148 SMs)
template <int EPI, int BMT = 128>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(const GemmArgs g) {
  constexpr int RH = BMT / 64;                              // row halves of 64 per tile; thread rows: 4 per half
  constexpr int TR = 4 * RH;
  __shared__ __align__(16) float As[2][BK][BMT + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BMT, n0 = blockIdx.y * BN;

  // global -> register staging
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;          // rows a_row (, a_row+64)
  const int b_k = tid >> 5, b_n = (tid & 31) * 4;           // k rows b_k, b_k+8
  float4 ra[RH], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < RH; ++i) {
      const int r = m0 + a_row + 64 * i;
      ra[i] = (r < g.M) ? *reinterpret_cast<const float4*>(g.A + (size_t)r * g.lda + k0 + a_k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) rb[i] = *reinterpret_cast<const float4*>(g.W + (size_t)(k0 + b_k + 8 * i) * g.Nc + n0 + b_n);
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < RH; ++i) {
      const int r = a_row + 64 * i;
      As[buf][a_k + 0][r] = ra[i].x; As[buf][a_k + 1][r] = ra[i].y; As[buf][a_k + 2][r] = ra[i].z; As[buf][a_k + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&Bs[buf][b_k + 8 * i][b_n]) = rb[i];
  };

  float acc[TR][8];
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  const int nk = g.K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TR];
#pragma unroll
      for (int hh = 0; hh < RH; ++hh) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][64 * hh + ty * 4]);
        av[4 * hh] = a0.x; av[4 * hh + 1] = a0.y; av[4 * hh + 2] = a0.z; av[4 * hh + 3] = a0.w;
      }
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TR; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + jh * 64 + tx * 4;
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (EPI == EPI_BIAS_RELU) {
        const float4 bb = *reinterpret_cast<const float4*>(g.bias + col);
        v.x = fmaxf(v.x + bb.x, 0.f); v.y = fmaxf(v.y + bb.y, 0.f); v.z = fmaxf(v.z + bb.z, 0.f); v.w = fmaxf(v.w + bb.w, 0.f);
      } else if (EPI == EPI_RES_BN) {
        if (g.bias != nullptr) {
          const float4 bb = *reinterpret_cast<const float4*>(g.bias + col);
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        const float4 r = *reinterpret_cast<const float4*>(g.res + (size_t)row * g.ldr + col);
        const float4 s = *reinterpret_cast<const float4*>(g.bn_s + col);
        const float4 t = *reinterpret_cast<const float4*>(g.bn_t + col);
        v.x = fmaf(r.x + v.x, s.x, t.x); v.y = fmaf(r.y + v.y, s.y, t.y);
        v.z = fmaf(r.z + v.z, s.z, t.z); v.w = fmaf(r.w + v.w, s.w, t.w);
      }
      if (EPI == EPI_TRAIN) {
        if (g.bias != nullptr) {
          const float4 bb = *reinterpret_cast<const float4*>(g.bias + col);
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (g.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (g.mask != nullptr) {
          const float4 mk = *reinterpret_cast<const float4*>(g.mask + (size_t)row * g.ldm + col);
          v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
        }
        if (g.accumulate) {
          const float4 c0 = *reinterpret_cast<const float4*>(g.C + (size_t)row * g.ldc + col);
          v.x += c0.x; v.y += c0.y; v.z += c0.z; v.w += c0.w;
        }
      }
      if (EPI == EPI_DEC) {       // four plain [M][128] matrices: K', V, LK' (keys[0..2]) and P_sc
        float* base = col < 3 * D ? g.keys + (size_t)(col / D) * g.M * D : g.psc;
        *reinterpret_cast<float4*>(base + (size_t)row * D + col % D) = v;
      } else {
        *reinterpret_cast<float4*>(g.C + (size_t)row * g.ldc + col) = v;
      }
    }
  }
}

// AGH_GEMM=simt forces the CUDA-core SGEMM (A/B comparisons); default is the tcgen05 path
static int g_gemm_backend = -1;
static bool use_tensor_cores() {
  if (g_gemm_backend < 0) {
    const char* e = getenv("AGH_GEMM");
    g_gemm_backend = (e != nullptr && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return g_gemm_backend == 1;
}

bool gemm_tc_enabled() { return use_tensor_cores(); }

template <int EPI>
static int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  if (use_tensor_cores() && g.Wt != nullptr) return launch_gemm_tc<EPI>(g, st);
  dim3 grid(ceil_div(g.M, BM), g.Nc / BN);
  sgemm_kernel<EPI><<<grid, 256, 0, st>>>(g);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

int launch_sgemm_train(const GemmArgs& g, cudaStream_t st) {
  // fill the GPU: 64-row tiles when 128-row tiles would give fewer CTAs than two waves of SMs
  if (ceil_div(g.M, BM) * (g.Nc / BN) < 296) {
    dim3 grid(ceil_div(g.M, 64), g.Nc / BN);
    sgemm_kernel<EPI_TRAIN, 64><<<grid, 256, 0, st>>>(g);
    AGH_LAUNCH_CHECK();
    return AGH_OK;
  }
  dim3 grid(ceil_div(g.M, BM), g.Nc / BN);
  sgemm_kernel<EPI_TRAIN><<<grid, 256, 0, st>>>(g);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

// ------------------------------------------------------------------------------------------
// self-attention, one CTA per (instance, head): graph_encoder.py:83-104
// ------------------------------------------------------------------------------------------
// One CTA = two heads of one instance, 64 threads per head, TWO query rows per thread: every K_j / V_j row read
// from shared memory (broadcast LDS.128) feeds two rows, which halves the shared-memory instruction stream that
// otherwise ties with the FMAs.  Online softmax in base 2, four keys per iteration.
__global__ void __launch_bounds__(128) mha_kernel(const float* __restrict__ qkv, int n, float* __restrict__ heads) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / (H / 2);
  const int hl = threadIdx.x >> 6;                           // local head 0/1
  const int h = (blockIdx.x % (H / 2)) * 2 + hl;
  const int t = threadIdx.x & 63;
  float4* sKh = reinterpret_cast<float4*>(sm) + (size_t)hl * n * 8;     // [n][4] float4 K, then [n][4] float4 V
  float4* sVh = sKh + (size_t)n * 4;
  const float* base = qkv + (size_t)b * n * (3 * D);
  for (int idx = t; idx < n * 4; idx += 64) {
    const int j = idx >> 2, c = idx & 3;
    sKh[idx] = *reinterpret_cast<const float4*>(base + (size_t)j * 3 * D + D + h * HD + c * 4);
    sVh[idx] = *reinterpret_cast<const float4*>(base + (size_t)j * 3 * D + 2 * D + h * HD + c * 4);
  }
  __syncthreads();
  for (int i0 = t; i0 < n; i0 += 128) {
    const int i1 = i0 + 64;
    const bool has1 = i1 < n;
    float q0[HD], q1[HD], o0[HD], o1[HD];
    // norm_factor = 1/sqrt(16) (graph_encoder.py:40,89) and log2(e) go into q once: the softmax runs in base 2
    const float qs = 0.25f * 1.4426950408889634f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 v0 = *reinterpret_cast<const float4*>(base + (size_t)i0 * 3 * D + h * HD + c * 4);
      const float4 v1 = has1 ? *reinterpret_cast<const float4*>(base + (size_t)i1 * 3 * D + h * HD + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      q0[4 * c] = v0.x * qs; q0[4 * c + 1] = v0.y * qs; q0[4 * c + 2] = v0.z * qs; q0[4 * c + 3] = v0.w * qs;
      q1[4 * c] = v1.x * qs; q1[4 * c + 1] = v1.y * qs; q1[4 * c + 2] = v1.z * qs; q1[4 * c + 3] = v1.w * qs;
    }
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int k = 0; k < HD; ++k) { o0[k] = 0.f; o1[k] = 0.f; }
    auto score2 = [&](int j, float& s0, float& s1) {
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 kv = sKh[j * 4 + c];
        a0 = fmaf(q0[4 * c], kv.x, a0); a1 = fmaf(q0[4 * c + 1], kv.y, a1); a0 = fmaf(q0[4 * c + 2], kv.z, a0); a1 = fmaf(q0[4 * c + 3], kv.w, a1);
        b0 = fmaf(q1[4 * c], kv.x, b0); b1 = fmaf(q1[4 * c + 1], kv.y, b1); b0 = fmaf(q1[4 * c + 2], kv.z, b0); b1 = fmaf(q1[4 * c + 3], kv.w, b1);
      }
      s0 = a0 + a1; s1 = b0 + b1;
    };
    auto accumulate2 = [&](int j, float p0, float p1) {
      l0 += p0; l1 += p1;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 vv = sVh[j * 4 + c];
        o0[4 * c] = fmaf(p0, vv.x, o0[4 * c]); o0[4 * c + 1] = fmaf(p0, vv.y, o0[4 * c + 1]);
        o0[4 * c + 2] = fmaf(p0, vv.z, o0[4 * c + 2]); o0[4 * c + 3] = fmaf(p0, vv.w, o0[4 * c + 3]);
        o1[4 * c] = fmaf(p1, vv.x, o1[4 * c]); o1[4 * c + 1] = fmaf(p1, vv.y, o1[4 * c + 1]);
        o1[4 * c + 2] = fmaf(p1, vv.z, o1[4 * c + 2]); o1[4 * c + 3] = fmaf(p1, vv.w, o1[4 * c + 3]);
      }
    };
    auto raise_max = [&](float cm, float& m, float& l, float (&o)[HD]) {   // rescale only when the running max rises
      if (cm > m) {
        const float c0 = exp2f(m - cm);
        l *= c0;
#pragma unroll
        for (int k = 0; k < HD; ++k) o[k] *= c0;
        m = cm;
      }
    };
    int j = 0;
    for (; j + 4 <= n; j += 4) {
      float s0[4], s1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) score2(j + u, s0[u], s1[u]);
      raise_max(fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s0[2], s0[3])), m0, l0, o0);
      raise_max(fmaxf(fmaxf(s1[0], s1[1]), fmaxf(s1[2], s1[3])), m1, l1, o1);
#pragma unroll
      for (int u = 0; u < 4; ++u) accumulate2(j + u, exp2f(s0[u] - m0), exp2f(s1[u] - m1));
    }
    for (; j < n; ++j) {
      float s0, s1;
      score2(j, s0, s1);
      raise_max(s0, m0, l0, o0);
      raise_max(s1, m1, l1, o1);
      accumulate2(j, exp2f(s0 - m0), exp2f(s1 - m1));
    }
    {
      const float inv = 1.0f / l0;
      float* out = heads + ((size_t)b * n + i0) * D + h * HD;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(out + 4 * c) = make_float4(o0[4 * c] * inv, o0[4 * c + 1] * inv, o0[4 * c + 2] * inv, o0[4 * c + 3] * inv);
    }
    if (has1) {
      const float inv = 1.0f / l1;
      float* out = heads + ((size_t)b * n + i1) * D + h * HD;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(out + 4 * c) = make_float4(o1[4 * c] * inv, o1[4 * c + 1] * inv, o1[4 * c + 2] * inv, o1[4 * c + 3] * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------
// qbase = W_fc mean_n(h) + fleet_emb : attention_model.py:395-402
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) qbase_kernel(const float* __restrict__ w, const float* __restrict__ h,
                                                    const int32_t* __restrict__ fleet_ids, int fleet, int n,
                                                    float* __restrict__ qbase) {
  __shared__ float mean[D];
  const int b = blockIdx.x, c = threadIdx.x;
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += h[((size_t)b * n + j) * D + c];
  mean[c] = s / (float)n;
  __syncthreads();
  float acc = 0.f;
  const float* wfc = w + W_FC_T;
#pragma unroll 8
  for (int e = 0; e < D; ++e) acc = fmaf(wfc[(size_t)e * D + c], mean[e], acc);
  const int f = fleet_ids != nullptr ? fleet_ids[b] : fleet;
  qbase[(size_t)b * D + c] = acc + w[W_FLEET + (size_t)f * D + c];
}

}  // namespace agh

using namespace agh;

extern "C" int agh_set_gemm_backend(int backend) {
  if (backend != 0 && backend != 1) return AGH_E_ARG;
  agh::g_gemm_backend = backend;
  return AGH_OK;
}

extern "C" size_t agh_encoder_workspace_bytes(int B, int N) {
  // qkv [M][384] + heads [M][128] + x1 [M][128] + hidden [M][512] + xa [M][128]
  const size_t M = (size_t)B * (N + 1);
  return M * (3 * D + D + D + FF + D) * sizeof(float) + 256 + ffn_tc_scratch_bytes() + 1024;
}

// AGH_FFN=unfused keeps the FF sublayer as two GEMMs with the hidden activations in HBM (A/B comparisons and tests)
// The fused kernel is used for EVERY row count: results must not depend on how many instances share a launch (level-batched
// fleets, shards of a multi-GPU batch and the tests rely on bit-identical per-instance results), and the two forms differ in the
// last bits.  Measured per launch (tools/sweep_encoder_kernels.py, profiles/r2_encoder_kernel_sweep.md): it wins from about 20,000
// rows (45 vs 48 us at M = 21,000; 788 vs 1,225 us at 1.01 M) and costs a few us below (33 vs 27 us at M = 1,344).
// 1 = fused (default), 0 = two GEMMs, 2 = same as 1 (kept for symmetry with agh_set_mha_tc).
constexpr int FFN_FUSED_MIN_ROWS = 0;
static int g_ffn_fused = -1;
static bool use_fused_ffn(int M) {
  if (g_ffn_fused < 0) {
    const char* e = getenv("AGH_FFN");
    g_ffn_fused = (e != nullptr && strcmp(e, "unfused") == 0) ? 0 : (e != nullptr && strcmp(e, "fused") == 0) ? 2 : 1;
  }
  if (!use_tensor_cores() || !ffn_tc_available()) return false;
  return g_ffn_fused == 2 || (g_ffn_fused == 1 && M >= FFN_FUSED_MIN_ROWS);
}

// AGH_MHA=mma keeps the encoder attention on the mma.sync kernel for every graph size (A/B comparisons and tests)
// The tcgen05 kernel always works on 112 key columns and 128 query rows; the mma.sync kernel's cost grows with n^2.  Measured
// at B = 10,000: n = 101: 676 vs 1,057 us, n = 51: 835 vs 455 us, n = 21: 862 vs 188 us -> "auto" (1) takes the tcgen05 kernel
// from 72 nodes; 2 forces it for every n <= 112 (unit tests).
constexpr int MHA_TC_MIN_NODES = 72;
static int g_mha_tc = -1;
static bool use_mha_tc(int n) {
  if (g_mha_tc < 0) {
    const char* e = getenv("AGH_MHA");
    g_mha_tc = (e != nullptr && strcmp(e, "mma") == 0) ? 0 : (e != nullptr && strcmp(e, "tc") == 0) ? 2 : 1;
  }
  if (!mha_tc_supported(n)) return false;
  return g_mha_tc == 2 || (g_mha_tc == 1 && n >= MHA_TC_MIN_NODES);
}

extern "C" int agh_set_mha_tc(int on) {
  if (on < 0 || on > 2) return AGH_E_ARG;
  g_mha_tc = on;
  return AGH_OK;
}

extern "C" int agh_set_ffn_fused(int fused) {
  if (fused < 0 || fused > 2) return AGH_E_ARG;
  g_ffn_fused = fused;
  return AGH_OK;
}

extern "C" int agh_init_embed(const float* wblob, const uint8_t* coord, const float* demand, const float* tw_left,
                              const double* tw_right, int twr_f32, int B, int N, float* h0, void* stream) {
  AGH_CHECK_ARG(wblob && coord && demand && tw_left && tw_right && h0 && B >= 0 && N >= 1);
  if (B == 0) return AGH_OK;
  const size_t warps = ((size_t)B * (N + 1) + 31) / 32;        // one warp per 32 rows, 8 warps per block
  init_embed_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(wblob, coord, demand, tw_left, tw_right,
                                                                                     twr_f32, B, N, h0);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_mha_fwd(const float* qkv, int B, int N, float* heads, void* stream) {
  AGH_CHECK_ARG(qkv && heads && B >= 0 && N >= 1 && N < AGH_MAX_N);
  if (B == 0) return AGH_OK;
  const int n = N + 1;
  const size_t mha_smem = (size_t)n * 4 * HD * sizeof(float);
  if (mha_smem > 48 * 1024) AGH_CUDA(cudaFuncSetAttribute(mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mha_smem));
  mha_kernel<<<B * (H / 2), 128, mha_smem, (cudaStream_t)stream>>>(qkv, n, heads);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_encode_layers(const float* wblob, const float* h0, int B, int N, float* h_out, void* workspace,
                                 void* stream) {
  AGH_CHECK_ARG(wblob && h0 && h_out && workspace && B >= 0 && N >= 1 && N < AGH_MAX_N);
  if (B == 0) return AGH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = N + 1;
  const int M = B * n;
  float* qkv = reinterpret_cast<float*>(workspace);
  float* heads = qkv + (size_t)M * 3 * D;
  float* x1 = heads + (size_t)M * D;
  float* hid = x1 + (size_t)M * D;
  float* xa = hid + (size_t)M * FF;
  // FP16-split FF weights of the fused feed-forward kernel: behind the activations, 1024-byte aligned
  void* ffn_scratch = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(xa + (size_t)M * D) + 256 + 1023) & ~(uintptr_t)1023);
  const bool ffn_fused = use_fused_ffn(M);
  if (ffn_fused) {
    const int rc = ffn_tc_split_weights(wblob, ffn_scratch, st);
    if (rc) return rc;
  }
  const float* x = h0;
  const int mha_threads = 128;                                     // 2 heads x 64 threads, 2 query rows per thread
  const size_t mha_smem = (size_t)n * 4 * HD * sizeof(float);
  if (mha_smem > 48 * 1024) AGH_CUDA(cudaFuncSetAttribute(mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mha_smem));
  for (int l = 0; l < LAYERS; ++l) {
    const float* lw = wblob + W_LAYER0 + (size_t)l * L_SIZE;
    const float* tw = wblob + W_TC0 + (size_t)l * T_SIZE;
    float* xout = (l == LAYERS - 1) ? h_out : xa;
    // Between two tensor-core kernels the wide intermediates (qkv, FF hidden) are kept SLAB-MAJOR -- Nc/128 separate
    // [M][128] matrices -- so that every CTA of the producing GEMM stores one contiguous 64 KB block per tile
    // (measured: the same K = 128 x Nc = 512 GEMM takes 570 us with that output form and 815 us row-major).
    static const bool mha_simt = getenv("AGH_MHA") != nullptr && strcmp(getenv("AGH_MHA"), "simt") == 0;
    const bool tcg = use_tensor_cores();
    const bool qkv_slab = tcg && !mha_simt;
    GemmArgs g{};
    g.A = x; g.lda = D; g.W = lw + L_WQKV; g.Wt = tw + T_WQKV; g.C = qkv; g.ldc = 3 * D; g.M = M; g.K = D; g.Nc = 3 * D;
    g.c_slab = qkv_slab;
    int rc = launch_gemm<EPI_PLAIN>(g, st);
    if (rc) return rc;
    if (tcg && !mha_simt && qkv_slab && use_mha_tc(n)) {
      rc = launch_mha_tc(qkv, B, n, heads, st);
      if (rc) return rc;
    } else if (tcg && !mha_simt) {
      rc = launch_mha_mma(qkv, qkv_slab, B, n, heads, st);
      if (rc) return rc;
    } else {
      mha_kernel<<<B * (H / 2), mha_threads, mha_smem, st>>>(qkv, n, heads);
      AGH_LAUNCH_CHECK();
    }
    g = GemmArgs{};
    g.A = heads; g.lda = D; g.W = lw + L_WO; g.Wt = tw + T_WO; g.C = x1; g.ldc = D; g.M = M; g.K = D; g.Nc = D;
    g.res = x; g.ldr = D; g.bn_s = lw + L_BN1S; g.bn_t = lw + L_BN1T;
    rc = launch_gemm<EPI_RES_BN>(g, st);
    if (rc) return rc;
    if (ffn_fused) {
      rc = launch_ffn_tc(x1, ffn_scratch, l, lw + L_B1, lw + L_B2, lw + L_BN2S, lw + L_BN2T, xout, M, st);
      if (rc) return rc;
      x = xout;
      continue;
    }
    g = GemmArgs{};
    g.A = x1; g.lda = D; g.W = lw + L_W1T; g.Wt = tw + T_W1; g.C = hid; g.ldc = FF; g.M = M; g.K = D; g.Nc = FF; g.bias = lw + L_B1;
    g.c_slab = tcg;
    rc = launch_gemm<EPI_BIAS_RELU>(g, st);
    if (rc) return rc;
    g = GemmArgs{};
    g.A = hid; g.lda = FF; g.W = lw + L_W2T; g.Wt = tw + T_W2; g.C = xout; g.ldc = D; g.M = M; g.K = FF; g.Nc = D; g.bias = lw + L_B2;
    g.res = x1; g.ldr = D; g.bn_s = lw + L_BN2S; g.bn_t = lw + L_BN2T;
    g.a_slab = tcg;
    rc = launch_gemm<EPI_RES_BN>(g, st);
    if (rc) return rc;
    x = xout;
  }
  return AGH_OK;
}

// Self-attention of all heads alone (graph_encoder.py:83-104): qkv slab-major [3][B*n][128] (Q, K, V stacked) -> heads [B*n][128],
// through the kernel agh_encode_layers would pick for this graph size (tcgen05 up to 112 nodes unless switched off, else mma.sync).
extern "C" int agh_encoder_mha(const float* qkv, int B, int N, float* heads, void* stream) {
  AGH_CHECK_ARG(qkv && heads && B >= 0 && N >= 1 && N < AGH_MAX_N);
  if (B == 0) return AGH_OK;
  const int n = N + 1;
  if (!use_tensor_cores()) {
    set_error("agh_encoder_mha: the tensor-core back end is switched off (use agh_mha_fwd for the CUDA-core kernel)");
    return AGH_E_UNSUPPORTED;
  }
  if (use_mha_tc(n)) return launch_mha_tc(qkv, B, n, heads, (cudaStream_t)stream);
  return launch_mha_mma(qkv, true, B, n, heads, (cudaStream_t)stream);
}

// Feed-forward sublayer of one encoder layer alone (graph_encoder.py:159-172 with eval-mode BatchNorm): y = BN2(x + FF(x)).
// workspace: M * 512 floats (hidden activations of the two-GEMM form) + 256 + ffn scratch + 1024 bytes.
extern "C" size_t agh_encoder_ffn_workspace_bytes(int M) {
  return (size_t)M * FF * sizeof(float) + 256 + ffn_tc_scratch_bytes() + 1024;
}

extern "C" int agh_encoder_ffn(const float* wblob, int layer, const float* x, int M, float* y, void* workspace, void* stream) {
  AGH_CHECK_ARG(wblob && x && y && workspace && M >= 0 && layer >= 0 && layer < LAYERS && x != y);
  if (M == 0) return AGH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float* lw = wblob + W_LAYER0 + (size_t)layer * L_SIZE;
  const float* tw = wblob + W_TC0 + (size_t)layer * T_SIZE;
  float* hid = reinterpret_cast<float*>(workspace);
  if (use_fused_ffn(M)) {
    void* scratch = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(hid + (size_t)M * FF) + 256 + 1023) & ~(uintptr_t)1023);
    int rc = ffn_tc_split_weights(wblob, scratch, st);
    if (rc) return rc;
    return launch_ffn_tc(x, scratch, layer, lw + L_B1, lw + L_B2, lw + L_BN2S, lw + L_BN2T, y, M, st);
  }
  const bool tcg = use_tensor_cores();
  GemmArgs g{};
  g.A = x; g.lda = D; g.W = lw + L_W1T; g.Wt = tw + T_W1; g.C = hid; g.ldc = FF; g.M = M; g.K = D; g.Nc = FF; g.bias = lw + L_B1;
  g.c_slab = tcg;
  int rc = launch_gemm<EPI_BIAS_RELU>(g, st);
  if (rc) return rc;
  g = GemmArgs{};
  g.A = hid; g.lda = FF; g.W = lw + L_W2T; g.Wt = tw + T_W2; g.C = y; g.ldc = D; g.M = M; g.K = FF; g.Nc = D; g.bias = lw + L_B2;
  g.res = x; g.ldr = D; g.bn_s = lw + L_BN2S; g.bn_t = lw + L_BN2T;
  g.a_slab = tcg;
  return launch_gemm<EPI_RES_BN>(g, st);
}

extern "C" int agh_encode(const float* wblob, const uint8_t* coord, const float* demand, const float* tw_left,
                          const double* tw_right, int twr_f32, int B, int N, float* h_out, void* workspace, void* stream) {
  AGH_CHECK_ARG(h_out != nullptr);
  // h_out doubles as the h0 buffer: layer 0 only reads it before the last layer writes it
  // (x -> xa -> xa ... -> h_out; with 3 layers h0 must not alias h_out, so keep it in the workspace tail)
  AGH_CHECK_ARG(workspace != nullptr);
  const size_t M = (size_t)B * (N + 1);
  float* h0 = reinterpret_cast<float*>(workspace) + M * (3 * D + D + D + FF);   // == xa slot; layer 0 reads it, layer 0's FF2 overwrites xa
  (void)h0;
  // layer 0 reads x=h0 for the QKV GEMM and as the residual of the out-projection, and writes
  // xa only in its last GEMM, whose inputs are hid and x1 -- so h0 may live in the xa slot.
  int rc = agh_init_embed(wblob, coord, demand, tw_left, tw_right, twr_f32, B, N, h0, stream);
  if (rc) return rc;
  return agh_encode_layers(wblob, h0, B, N, h_out, workspace, stream);
}

extern "C" int agh_precompute(const float* wblob, const float* h, const int32_t* fleet_ids, int fleet, int B, int N,
                              float* keys, float* psc, float* qbase, void* stream) {
  AGH_CHECK_ARG(wblob && h && keys && psc && qbase && B >= 0 && N >= 1);
  AGH_CHECK_ARG(fleet_ids != nullptr || (fleet >= 0 && fleet < 10));
  if (B == 0) return AGH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = N + 1;
  GemmArgs g{};
  g.A = h; g.lda = D; g.W = wblob + W_DEC; g.Wt = wblob + W_TC_DEC; g.M = B * n; g.K = D; g.Nc = 4 * D; g.keys = keys; g.psc = psc; g.n = n;
  int rc = launch_gemm<EPI_DEC>(g, st);
  if (rc) return rc;
  qbase_kernel<<<B, D, 0, st>>>(wblob, h, fleet_ids, fleet, n, qbase);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}
