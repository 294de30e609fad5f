// Shared constants, weight-blob offsets and error plumbing for libagh_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/agh_b200.h"

namespace agh {

constexpr int D = AGH_D;            // embedding width
constexpr int H = AGH_H;            // heads
constexpr int HD = AGH_D / AGH_H;   // 16
constexpr int FF = AGH_FF;
constexpr int LAYERS = AGH_LAYERS;
constexpr int KS = AGH_KEY_STRIDE;  // padded key-row stride (floats)
constexpr int NODE_SIZE = AGH_NODE_SIZE;

// ---- weight blob layout (floats); mirrored by agh_b200/nets/weights.py ----
constexpr size_t W_LOC_EMB = 0;                               // [92][128]
constexpr size_t W_INIT_W = W_LOC_EMB + 92 * D;               // [3][128]
constexpr size_t W_INIT_B = W_INIT_W + 3 * D;                 // [128]
constexpr size_t W_LAYER0 = W_INIT_B + D;
constexpr size_t L_WQKV = 0;                                  // [128][384]
constexpr size_t L_WO = L_WQKV + D * 3 * D;                   // [128][128]
constexpr size_t L_BN1S = L_WO + D * D;
constexpr size_t L_BN1T = L_BN1S + D;
constexpr size_t L_W1T = L_BN1T + D;                          // [128][512]
constexpr size_t L_B1 = L_W1T + D * FF;
constexpr size_t L_W2T = L_B1 + FF;                           // [512][128]
constexpr size_t L_B2 = L_W2T + FF * D;
constexpr size_t L_BN2S = L_B2 + D;
constexpr size_t L_BN2T = L_BN2S + D;
constexpr size_t L_SIZE = L_BN2T + D;
constexpr size_t W_DEC = W_LAYER0 + LAYERS * L_SIZE;          // [128][512]
constexpr size_t W_FC_T = W_DEC + D * 4 * D;                  // [128][128]
constexpr size_t W_FLEET = W_FC_T + D * D;                    // [10][128]
constexpr size_t W_CAP = W_FLEET + 10 * D;                    // [128]
constexpr size_t W_TIME = W_CAP + D;                          // [128]
// K-major ([out][in]) copies of the GEMM weights for the tensor-core path (TMA + tcgen05 want K contiguous)
constexpr size_t W_TC0 = W_TIME + D;
constexpr size_t T_WQKV = 0;                                  // [384][128]
constexpr size_t T_WO = T_WQKV + 3 * D * D;                   // [128][128]
constexpr size_t T_W1 = T_WO + D * D;                         // [512][128]
constexpr size_t T_W2 = T_W1 + FF * D;                        // [128][512]
constexpr size_t T_SIZE = T_W2 + D * FF;
constexpr size_t W_TC_DEC = W_TC0 + LAYERS * T_SIZE;          // [512][128]
constexpr size_t W_TOTAL = W_TC_DEC + 4 * D * D;

void set_error(const char* fmt, ...);
void count_launch();   // agh_launch_count(): kernels launched by this library since load

#define AGH_CHECK_ARG(cond)                                                         \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      agh::set_error("%s:%d: argument check failed: %s", __FILE__, __LINE__, #cond); \
      return AGH_E_ARG;                                                             \
    }                                                                               \
  } while (0)

#define AGH_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      agh::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return AGH_E_CUDA;                                                                        \
    }                                                                                           \
  } while (0)

#define AGH_LAUNCH_CHECK()         \
  do {                             \
    agh::count_launch();           \
    AGH_CUDA(cudaGetLastError());  \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace agh
