// Shared declarations of the two GEMM back ends (CUDA-core SGEMM in encoder.cu, tcgen05 in gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace agh {

enum { EPI_PLAIN = 0, EPI_BIAS_RELU = 1, EPI_RES_BN = 2, EPI_DEC = 3, EPI_TRAIN = 4 };

struct GemmArgs {
  const float* A; int lda;
  const float* W;          // [K][Nc] row-major   (CUDA-core path)
  const float* Wt;         // [Nc][K] row-major   (tensor-core path, K-major)
  float* C; int ldc;
  int M, K, Nc;
  const float* bias;       // [Nc] or null
  const float* res; int ldr;   // residual
  const float* bn_s; const float* bn_t;
  // EPI_DEC: Nc = 512 = four [M][128] outputs: keys[0] = K', keys[1] = V, keys[2] = LK' (keys is [3][M][128]), psc
  float* keys; float* psc; int n;
  // tensor-core path only: C written / A read as Nc/128 (K/128) separate [M][128] matrices ("slab-major": every CTA's
  // 128 x 128 output tile is one contiguous 64 KB block) instead of one row-major [M][Nc] ([M][K]) matrix
  int c_slab, a_slab;
  // EPI_TRAIN (training primitives, CUDA-core path only): v = acc [+ bias] [relu] [0 where mask <= 0] [+ C]
  int relu, accumulate; const float* mask; int ldm;
};

template <int EPI>
int launch_gemm_tc(const GemmArgs& g, cudaStream_t st);

// whether the tensor-core back end is selected (agh_set_gemm_backend / AGH_GEMM)
bool gemm_tc_enabled();

// agh_gemm's fast path: plain row-major A [M][K] (lda) times W [K][N] (ld = N) through the tiled CUDA-core SGEMM
int launch_sgemm_train(const GemmArgs& g, cudaStream_t st);

// Fused FF sublayer on tcgen05 (ffn_tc.cu): y = BN(x + W2 relu(W1 x + b1) + b2) with the hidden activations kept on chip.
// `scratch` (ffn_tc_scratch_bytes()) receives the FP16-split FF weights of all layers (ffn_tc_split_weights, once per call).
bool ffn_tc_available();
size_t ffn_tc_scratch_bytes();
int ffn_tc_split_weights(const float* wblob, void* scratch, cudaStream_t st);
int launch_ffn_tc(const float* x, const void* scratch, int layer, const float* b1, const float* b2, const float* bn_s,
                  const float* bn_t, float* y, int M, cudaStream_t st);

// Self-attention of all heads on mma.sync tensor cores (mha_mma.cu): qkv [B*n][384] (or slab-major [3][B*n][128]) -> heads [B*n][128].
int launch_mha_mma(const float* qkv, bool slab_major, int B, int n, float* heads, cudaStream_t st);

// Self-attention on tcgen05 (mha_tc.cu) for graphs of up to 112 nodes; qkv slab-major [3][B*n][128].
bool mha_tc_supported(int n);
int launch_mha_tc(const float* qkv, int B, int n, float* heads, cudaStream_t st);

}  // namespace agh
