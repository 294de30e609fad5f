// Device and host helpers shared by the tcgen05 kernels (gemm_tc.cu, ffn_tc.cu): mbarrier / TMA / UMMA descriptor and
// instruction wrappers, TMEM loads, the in-place FP16 (or TF32) operand split, tensor-map encoding.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace agh {
namespace tc {

// AGH_TC_F16 = 1 (shipped): the operands are split into TWO FP16 values, x = hi + lo' / 2048 with hi = rn_f16(x) and
// lo' = rn_f16((x - hi) * 2048): the same 11 + 11 significant bits as the TF32 split, but tcgen05.mma.kind::f16 multiplies
// K = 16 per instruction at twice the tf32 rate (half the MMA count and half the read-modify-write passes over the TMEM
// accumulator per tile) and the split operands take half the shared memory, which buys a deeper TMA ring.  The
// remainder is pre-scaled by 2^11 so that it stays a NORMAL fp16 number for every |x| >= 2^-14 (no precision is lost
// to fp16's narrow exponent range); the cross-term accumulator is scaled back by 2^-11 in the epilogue (exact).
// Domain: |x| < 65504 (activations after BatchNorm / ReLU and weights are orders of magnitude below; the conversion
// saturates instead of producing inf).  AGH_TC_F16 = 0 selects the 3xTF32 form the scheme was validated against.
#ifndef AGH_TC_F16
#define AGH_TC_F16 1
#endif
constexpr bool F16 = AGH_TC_F16 != 0;
constexpr int BM = 128, BN = 128, BK = 32;             // TMA box: BK fp32 = 128 bytes = one swizzle-128B row
constexpr int BLK = BM * BK * 4;                       // one [128 x 32] fp32 block = one [128 x 64] fp16 block = 16 KB
constexpr int KST = F16 ? 2 * BK : BK;                 // K extent of one pipeline stage (F16: two raw boxes -> one hi + one lo' block)
constexpr int STAGES_WRES = F16 ? 4 : 3, STAGES_STREAM = 3;
constexpr float CROSS_SCALE = F16 ? 1.0f / 2048.0f : 1.0f;
constexpr int EPI_WARPS = 8;                          // two per TMEM lane quarter, each takes half of the column chunks
// Splitter warps: the per-role timeline of the FP16 build (profiles/r1h_fp16_split_notes.md, r1j) shows the K = 128 shapes
// bound by the splitters (1.1 us per K = 64 stage with four warps against a 1.75 us epilogue per two-stage tile), so
// the FP16 form runs eight: two threads per row, one per raw box.
constexpr int SPLIT_WARPS = F16 ? 8 : 4;
constexpr int FIRST_EPI = 2 + SPLIT_WARPS;            // warp index of the first epilogue warp (a multiple of 2: quarters by warp % 4)
constexpr int NTHREADS = 32 * (2 + SPLIT_WARPS + EPI_WARPS);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait until the phase with the given parity has completed (a fresh barrier has "completed" parity 1)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "elect.sync %%rx|%%px, %1;\n"
      "@%%px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1):
// 8-row atoms of 128 B rows, atoms 1024 B apart (SBO), LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (>>4), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset (>>4), bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                               // layout type SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, both operands K-major (cute::UMMA::InstrDescriptor)
// (a/b format field: kind::tf32 2 = TF32; kind::f16 0 = F16)
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
#if AGH_TC_F16
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
#else
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
#endif
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));      // the caller issues tcgen05.wait::ld before reading r[]
}

// 16 TMEM lanes x 32 columns in the mma "C fragment" distribution: thread (g = lane / 4, t = lane % 4) receives, for
// every 8-column group k, r[4k], r[4k+1] = (lane base + g, columns 8k + 2t, 8k + 2t + 1) and r[4k+2], r[4k+3] = the
// same columns of lane base + g + 8.  A quad therefore holds 32 contiguous bytes of one row: global accesses made
// straight from this layout are whole 32-byte sectors, with no transpose.
__device__ __forceinline__ void tmem_ld16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));      // the caller issues tcgen05.wait::ld before reading r[]
}

// x = hi + lo with hi = rn_tf32(x) and lo = rn_tf32(x - hi): both exactly representable in tf32 (so it does not
// matter whether the tensor core truncates or rounds its fp32 containers) and the rounding errors are unbiased.
// Round-to-nearest (ties away) with an integer add + mask: cvt.rna.tf32.f32 runs on the slow conversion pipe.
__device__ __forceinline__ float rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// split `nvec` float4 starting at `hi_ptr` in place, writing the remainders to `lo_ptr` (same element order)
__device__ __forceinline__ void split_block(unsigned char* hi_ptr, unsigned char* lo_ptr, int nvec, int t) {
#pragma unroll 4
  for (int i = t; i < nvec; i += 128) {
    float4* ph = reinterpret_cast<float4*>(hi_ptr) + i;
    const float4 v = *ph;
    float4 hi, lo;
    hi.x = rn_tf32(v.x); lo.x = rn_tf32(v.x - hi.x);
    hi.y = rn_tf32(v.y); lo.y = rn_tf32(v.y - hi.y);
    hi.z = rn_tf32(v.z); lo.z = rn_tf32(v.z - hi.z);
    hi.w = rn_tf32(v.w); lo.w = rn_tf32(v.w - hi.w);
    *ph = hi;
    *(reinterpret_cast<float4*>(lo_ptr) + i) = lo;
  }
}

// FP16 split of one stage, in place: `r0` / `r1` hold the raw fp32 boxes of k = 0..31 / 32..63 as TMA wrote them (128-byte
// rows, 16-byte chunk c of row r at chunk c ^ (r & 7): SWIZZLE_128B); afterwards r0 is the K-major SWIZZLE_128B fp16
// block of the hi parts ([128 rows][64 halfs]) and r1 that of the scaled remainders.  Thread t owns row t of both boxes,
// reads all of it before it writes any of it, and the swizzle makes every LDS.128 / STS.128 conflict-free.
__device__ __forceinline__ void split_f16_row(unsigned char* r0, unsigned char* r1, int t) {
  const int x = t & 7;
  unsigned char* p0 = r0 + t * 128;
  unsigned char* p1 = r1 + t * 128;
  float4 v[16];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    v[c] = *reinterpret_cast<const float4*>(p0 + ((c ^ x) << 4));
    v[8 + c] = *reinterpret_cast<const float4*>(p1 + ((c ^ x) << 4));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {                 // fp16 chunk j = k 8j .. 8j+7 = fp32 chunks 2j, 2j+1
    const float f[8] = {v[2 * j].x, v[2 * j].y, v[2 * j].z, v[2 * j].w, v[2 * j + 1].x, v[2 * j + 1].y, v[2 * j + 1].z, v[2 * j + 1].w};
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t h2;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(f[2 * e + 1]), "f"(f[2 * e]));      // {hi half, lo half} = {b, a}
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
      const float d0 = (f[2 * e] - hf.x) * 2048.0f, d1 = (f[2 * e + 1] - hf.y) * 2048.0f;
      uint32_t l2;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(d1), "f"(d0));
      hw[e] = h2; lw[e] = l2;
    }
    *reinterpret_cast<uint4*>(p0 + ((j ^ x) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(p1 + ((j ^ x) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// The same split with TWO threads per row (256 splitter threads): lanes l and l + 16 of a warp share row 16 w + (l & 15);
// the lower one reads the raw box of k = 0..31 and produces fp16 chunks 0..3 of hi and lo', the upper one k = 32..63 and
// chunks 4..7.  Both destinations of a row overlap both sources, hence the __syncwarp between the reads and the writes.
__device__ __forceinline__ void split_f16_half(unsigned char* r0, unsigned char* r1, int t) {
  const int row = (t >> 5) * 16 + (t & 15), half = (t >> 4) & 1, x = row & 7;
  const unsigned char* src = (half ? r1 : r0) + row * 128;
  float4 v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(src + ((c ^ x) << 4));
  __syncwarp();
  unsigned char* p0 = r0 + row * 128;
  unsigned char* p1 = r1 + row * 128;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int j = half * 4 + jj;
    const float f[8] = {v[2 * jj].x, v[2 * jj].y, v[2 * jj].z, v[2 * jj].w, v[2 * jj + 1].x, v[2 * jj + 1].y, v[2 * jj + 1].z, v[2 * jj + 1].w};
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t h2;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(f[2 * e + 1]), "f"(f[2 * e]));
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
      const float d0 = (f[2 * e] - hf.x) * 2048.0f, d1 = (f[2 * e + 1] - hf.y) * 2048.0f;
      uint32_t l2;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(d1), "f"(d0));
      hw[e] = h2; lw[e] = l2;
    }
    *reinterpret_cast<uint4*>(p0 + ((j ^ x) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(p1 + ((j ^ x) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D fp32 row-major [rows][cols] (row pitch ld floats), box = [128 rows][32 floats], 128-byte swizzle, OOB -> 0
inline int make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return AGH_E_CUDA;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)rc, (unsigned long long)rows,
              (unsigned long long)cols, (unsigned long long)ld);
    return AGH_E_CUDA;
  }
  return AGH_OK;
}

inline int sm_count() {      // of the CURRENT device (a process may drive several GPUs)
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

}  // namespace tc
}  // namespace agh
