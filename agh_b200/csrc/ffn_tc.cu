// Fused feed-forward sublayer of the encoder on tcgen05:   y = BN( x + W2 relu(W1 x + b1) + b2 )   (graph_encoder.py:159-172,
// eval-mode BatchNorm folded into scale / shift), one launch per layer.  The [M][512] hidden activations never leave
// the SM: per 128-row tile they are produced 64 columns at a time in TMEM, read back by the H-epilogue warps (bias, ReLU),
// split into the two FP16 terms and written to shared memory as the A operand of the second contraction.
// The unfused pair (gemm_tc.cu: FF1 bias+ReLU -> HBM -> FF2 residual+BN) moves 0.52 + 2.0 GB and 2.6 + 0.5 GB per launch
// at M = 1.01 M rows and is bound by HBM; this kernel moves x in and y out (1.0 GB) and is bound by the tensor pipe /
// the L2 -> SM weight stream.
//
// Same two-term operand split as gemm_tc.cu (x = hi + lo'/2048, three MMAs per product, cross terms in their own
// accumulator); the weights are split ONCE per encoder call into FP16 hi / lo' matrices (split_ffn_weights_kernel) and
// streamed per tile by TMA straight into the UMMA layout -- no splitter pass over them.
//
//   warp 0      : TMA producer  - the x tile (4 raw fp32 boxes) once per tile; weight stages (32 KB: W1 rows of one
//                                 64-wide hidden chunk, or the W2 columns of one) through a ring of 3
//   warp 1      : MMA issuer    - per tile: A_0, then for j = 1..7 { A_j, B_{j-1} }, then B_7, where
//                                 A_j: H[128 x 64] = x W1_j^T (K = 128, 24 MMAs N = 64),  B_j: Y += H_j W2_j^T (K = 64, 12 MMAs)
//   warps 2..9  : x splitters (in place, as gemm_tc.cu) and the Y epilogue (+ b2 + residual, BatchNorm, store)
//   warps 10..17: H epilogue    - tcgen05.ld {main, cross} -> + b1, ReLU -> FP16 hi / lo' -> smem (double-buffered)
// TMEM (512 columns): H main 64 | H cross 64 | Y main0 128 | Y main1 128 | Y cross 128.  The K = 512 contraction spreads
// its hi.hi terms over two accumulators (16 truncating accumulations each; the unfused kernel uses three of ~11).
#include "common.cuh"
#include "gemm.cuh"
#include "tc.cuh"

namespace agh {

namespace tc {

#if AGH_TC_F16

namespace ffn {
constexpr int HC = 64;                       // hidden columns per chunk
constexpr int NCH = FF / HC;                 // 8 chunks per tile
constexpr int WST = 3;                       // weight ring stages
constexpr int WSTAGE = 2 * BLK;              // 32 KB
constexpr int X_OFF = 0, HS_OFF = 4 * BLK, W_OFF = 8 * BLK, BAR_OFF = W_OFF + WST * WSTAGE;
// barriers
constexpr int B_XRAW = 0, B_XSPLIT = 2, B_XEMPTY = 4, B_WFULL = 5, B_WEMPTY = B_WFULL + WST, B_HFULL = B_WEMPTY + WST,
              B_HEMPTY = B_HFULL + 1, B_HSFULL = B_HEMPTY + 1, B_HSEMPTY = B_HSFULL + 2, B_YFULL = B_HSEMPTY + 2, B_YEMPTY = B_YFULL + 1,
              NBARS = B_YEMPTY + 1;
constexpr int VEC_OFF = BAR_OFF + 256;       // b2 | bn_s | bn_t
constexpr int SMEM = VEC_OFF + 3 * D * 4 + 1024;
static_assert(8 * NBARS + 16 <= 256, "barrier block");
static_assert(SMEM <= 232448, "shared memory budget");
constexpr int T_HMAIN = 0, T_HCROSS = 64, T_Y = 128;      // TMEM columns; Y: main0, main1, cross at +0, +128, +256
constexpr int NT = 32 * 18;
}  // namespace ffn

struct FfnArgs {
  const float* x;          // [M][128] input = residual
  float* y;                // [M][128]
  const float* b1;         // [512]
  const float* b2;         // [128]
  const float* bn_s; const float* bn_t;
  int M;
};

// FP16 split of the FF weights of one layer: W1 [512][128] -> w1s [1024][128] halfs (hi rows, then lo' rows);
// W2 [128][512] -> w2s [256][512] halfs.
__global__ void split_ffn_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2, __half* __restrict__ w1s,
                                         __half* __restrict__ w2s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * FF * D) return;
  const bool second = i >= FF * D;
  const int e = second ? i - FF * D : i;
  const float w = second ? w2[e] : w1[e];
  const __half hi = __float2half_rn(w);
  const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.0f);
  __half* dst = second ? w2s : w1s;
  dst[e] = hi;
  dst[FF * D + e] = lo;
}

__global__ void __launch_bounds__(ffn::NT, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const FfnArgs g) {
  using namespace ffn;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* xs = smem + X_OFF;             // 2 stages (k 0..63, 64..127) of [hi 16 KB | lo' 16 KB]
  unsigned char* hs = smem + HS_OFF;            // 2 buffers of [hi 16 KB | lo' 16 KB]: H chunk, 128 rows x 64 k
  unsigned char* wr = smem + W_OFF;
  unsigned char* bars = smem + BAR_OFF;
  const uint32_t bar_base = smem_u32(bars);
  auto bar = [&](int i) { return bar_base + 8 * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 * NBARS);
  float* s_vec = reinterpret_cast<float*>(smem + VEC_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_mt = (g.M + BM - 1) / BM;

  if (threadIdx.x >= 64 && threadIdx.x < 64 + D) {
    const int t = threadIdx.x - 64;
    s_vec[t] = g.b2[t];
    s_vec[D + t] = g.bn_s[t];
    s_vec[2 * D + t] = g.bn_t[t];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(B_XRAW + s), 1);
      mbar_init(bar(B_XSPLIT + s), 256);
      mbar_init(bar(B_HSFULL + s), 256);
      mbar_init(bar(B_HSEMPTY + s), 1);
    }
    mbar_init(bar(B_XEMPTY), 1);
    for (int s = 0; s < WST; ++s) {
      mbar_init(bar(B_WFULL + s), 1);
      mbar_init(bar(B_WEMPTY + s), 1);
    }
    mbar_init(bar(B_HFULL), 1);
    mbar_init(bar(B_HEMPTY), 8);
    mbar_init(bar(B_YFULL), 1);
    mbar_init(bar(B_YEMPTY), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      uint32_t wit = 0, tile = 0;
      auto load_w1 = [&](int j) {          // rows 64 j .. +63 of W1: [hi k 0..63 | hi k 64..127 | lo' k 0..63 | lo' k 64..127], 8 KB each
        const int s = wit % WST;
        mbar_wait(bar(B_WEMPTY + s), ((wit / WST) & 1) ^ 1);
        const uint32_t dst = smem_u32(wr + s * WSTAGE);
        mbar_expect_tx(bar(B_WFULL + s), WSTAGE);
        tma_load_2d(dst, &tmW1, bar(B_WFULL + s), 0, j * HC);
        tma_load_2d(dst + BLK / 2, &tmW1, bar(B_WFULL + s), 64, j * HC);
        tma_load_2d(dst + BLK, &tmW1, bar(B_WFULL + s), 0, FF + j * HC);
        tma_load_2d(dst + BLK + BLK / 2, &tmW1, bar(B_WFULL + s), 64, FF + j * HC);
        ++wit;
      };
      auto load_w2 = [&](int j) {          // columns 64 j .. +63 of W2, all 128 rows: [hi 16 KB | lo' 16 KB]
        const int s = wit % WST;
        mbar_wait(bar(B_WEMPTY + s), ((wit / WST) & 1) ^ 1);
        const uint32_t dst = smem_u32(wr + s * WSTAGE);
        mbar_expect_tx(bar(B_WFULL + s), WSTAGE);
        tma_load_2d(dst, &tmW2, bar(B_WFULL + s), j * HC, 0);
        tma_load_2d(dst + BLK, &tmW2, bar(B_WFULL + s), j * HC, D);
        ++wit;
      };
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
        mbar_wait(bar(B_XEMPTY), (tile & 1) ^ 1);
        for (int s = 0; s < 2; ++s) {
          mbar_expect_tx(bar(B_XRAW + s), 2 * BLK);
          tma_load_2d(smem_u32(xs + s * 2 * BLK), &tmX, bar(B_XRAW + s), s * 64, mt * BM);
          tma_load_2d(smem_u32(xs + s * 2 * BLK + BLK), &tmX, bar(B_XRAW + s), s * 64 + BK, mt * BM);
        }
        load_w1(0);
        for (int j = 1; j < NCH; ++j) {
          load_w1(j);
          load_w2(j - 1);
        }
        load_w2(NCH - 1);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_a = umma_idesc(BM, HC), idesc_b = umma_idesc(BM, BN);
    uint32_t wit = 0, tile = 0, chunk = 0;      // chunk: running count of A ops (H accumulator phases)
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
      auto op_a = [&](int j) {
        const int s = wit % WST;
        mbar_wait(bar(B_HEMPTY), (chunk & 1) ^ 1);                 // H accumulator drained by the H epilogue
        mbar_wait(bar(B_WFULL + s), (wit / WST) & 1);
        if (j == 0) mbar_wait(bar(B_XSPLIT), tile & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t wb = smem_u32(wr + s * WSTAGE);
        for (int kb = 0; kb < 2; ++kb) {
          if (j == 0 && kb == 1) {
            mbar_wait(bar(B_XSPLIT + 1), tile & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          if (elect_one()) {
            const uint32_t xb = smem_u32(xs + kb * 2 * BLK);
            const uint64_t da_hi = umma_desc(xb), da_lo = umma_desc(xb + BLK);
            const uint64_t dw_hi = umma_desc(wb + kb * (BLK / 2)), dw_lo = umma_desc(wb + BLK + kb * (BLK / 2));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              const uint32_t first = (kb | k) != 0;
              umma_tf32(tmem_base + T_HCROSS, da_lo + adv, dw_hi + adv, idesc_a, first);
              umma_tf32(tmem_base + T_HCROSS, da_hi + adv, dw_lo + adv, idesc_a, 1);
              umma_tf32(tmem_base + T_HMAIN, da_hi + adv, dw_hi + adv, idesc_a, first);
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(bar(B_WEMPTY + s));
          umma_commit(bar(B_HFULL));
          if (j == NCH - 1) umma_commit(bar(B_XEMPTY));           // the x tile is no longer read
        }
        __syncwarp();
        ++wit; ++chunk;
      };
      auto op_b = [&](int j) {
        const int s = wit % WST;
        const uint32_t c = tile * NCH + j;                         // running chunk index, as counted by the H epilogue
        const int b = c & 1;
        if (j == 0) mbar_wait(bar(B_YEMPTY), (tile & 1) ^ 1);      // Y accumulators drained by the Y epilogue
        mbar_wait(bar(B_WFULL + s), (wit / WST) & 1);
        mbar_wait(bar(B_HSFULL + b), (c >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t hb = smem_u32(hs + b * 2 * BLK), wb = smem_u32(wr + s * WSTAGE);
          const uint64_t da_hi = umma_desc(hb), da_lo = umma_desc(hb + BLK);
          const uint64_t dw_hi = umma_desc(wb), dw_lo = umma_desc(wb + BLK);
          const uint32_t t_main = tmem_base + T_Y + (j & 1) * BN, t_cross = tmem_base + T_Y + 2 * BN;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_tf32(t_cross, da_lo + adv, dw_hi + adv, idesc_b, (j | k) != 0);
            umma_tf32(t_cross, da_hi + adv, dw_lo + adv, idesc_b, 1);
            umma_tf32(t_main, da_hi + adv, dw_hi + adv, idesc_b, (j >= 2) || k != 0);
          }
          umma_commit(bar(B_WEMPTY + s));
          umma_commit(bar(B_HSEMPTY + b));
          if (j == NCH - 1) umma_commit(bar(B_YFULL));
        }
        __syncwarp();
        ++wit;
      };
      op_a(0);
      for (int j = 1; j < NCH; ++j) {
        op_a(j);
        op_b(j - 1);
      }
      op_b(NCH - 1);
    }
  } else if (warp < 10) {
    // ================= x splitters + Y epilogue (warps 2..9) =================
    const int t = threadIdx.x - 64;
    const int q = warp & 3;                                   // TMEM lane quarter
    const int chalf = (warp - 2) >> 2;                        // 0: column chunks 0,1   1: chunks 2,3
    const int g8 = lane >> 2, t4 = lane & 3;
    auto split_x = [&](uint32_t tile) {
      for (int s = 0; s < 2; ++s) {
        mbar_wait(bar(B_XRAW + s), tile & 1);
        split_f16_half(xs + s * 2 * BLK, xs + s * 2 * BLK + BLK, t);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(bar(B_XSPLIT + s));
      }
    };
    uint32_t tile = 0;
    if ((int)blockIdx.x < num_mt) split_x(0);
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
      if (mt + (int)gridDim.x < num_mt) split_x(tile + 1);       // next tile's x: lands once this tile's last A op has retired
      const int row0 = mt * BM + q * 32;
      const int c_first = chalf * 64, c_last = c_first + 32;
      float2 fres[16];
      auto load_res_frag = [&](int c0) {
#pragma unroll
        for (int rq = 0; rq < 4; ++rq) {
          const int row = row0 + rq * 8 + g8;
          const float* src = g.x + (size_t)row * D + c0 + 2 * t4;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            fres[rq * 4 + k] = row < g.M ? __ldg(reinterpret_cast<const float2*>(src + 8 * k)) : make_float2(0.f, 0.f);
        }
      };
      load_res_frag(c_first);
      mbar_wait(bar(B_YFULL), tile & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_acc = tmem_base + T_Y + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = c_first; c0 <= c_last; c0 += 32) {
        float r[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t ra[16];
          const uint32_t ta = t_acc + ((uint32_t)(hh * 16) << 16) + (uint32_t)c0;
          tmem_ld16x256b_x4(ta, ra);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) r[hh * 16 + i] = __uint_as_float(ra[i]);
          tmem_ld16x256b_x4(ta + BN, ra);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) r[hh * 16 + i] += __uint_as_float(ra[i]);
          tmem_ld16x256b_x4(ta + 2 * BN, ra);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) r[hh * 16 + i] = fmaf(__uint_as_float(ra[i]), CROSS_SCALE, r[hh * 16 + i]);
        }
        if (c0 == c_last) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_YEMPTY));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int cl = c0 + 8 * k + 2 * t4;
          const float2 bias = *reinterpret_cast<const float2*>(s_vec + cl);
          const float2 sc = *reinterpret_cast<const float2*>(s_vec + D + cl);
          const float2 sh = *reinterpret_cast<const float2*>(s_vec + 2 * D + cl);
#pragma unroll
          for (int rq = 0; rq < 4; ++rq) {
            const int row = row0 + rq * 8 + g8;
            const int ri = (rq >> 1) * 16 + k * 4 + (rq & 1) * 2;
            const float2 rs = fres[rq * 4 + k];
            float2 v;
            v.x = fmaf(rs.x + (r[ri] + bias.x), sc.x, sh.x);
            v.y = fmaf(rs.y + (r[ri + 1] + bias.y), sc.y, sh.y);
            if (row < g.M) *reinterpret_cast<float2*>(g.y + (size_t)row * D + cl) = v;
          }
        }
        if (c0 < c_last) load_res_frag(c0 + 32);
      }
    }
  } else {
    // ================= H epilogue (warps 10..17): TMEM lane quarter = warp % 4, column half = (warp - 10) / 4 =================
    const int q = warp & 3, ch = (warp - 10) >> 2;
    const int row = q * 32 + lane, x7 = row & 7;
    uint32_t chunk = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x) {
#pragma unroll 1
      for (int j = 0; j < NCH; ++j, ++chunk) {
        const float4* b1p = reinterpret_cast<const float4*>(g.b1 + j * HC + ch * 32);      // warp-uniform: L1 broadcast
        mbar_wait(bar(B_HFULL), chunk & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32);
        uint32_t ra[32], rb[32];
        tmem_ld32(ta + T_HMAIN, ra);
        tmem_ld32(ta + T_HCROSS, rb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_HEMPTY));
        const int b = chunk & 1;
        mbar_wait(bar(B_HSEMPTY + b), ((chunk >> 1) & 1) ^ 1);    // the B op that read this buffer two chunks ago has retired
        unsigned char* p0 = hs + b * 2 * BLK + row * 128;
        unsigned char* p1 = p0 + BLK;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 ba = __ldg(b1p + 2 * cc), bc = __ldg(b1p + 2 * cc + 1);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bc.x, bc.y, bc.z, bc.w};
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i0 = cc * 8 + 2 * e;
            const float f0 = fmaxf(fmaf(__uint_as_float(rb[i0]), CROSS_SCALE, __uint_as_float(ra[i0])) + bv[2 * e], 0.f);
            const float f1 = fmaxf(fmaf(__uint_as_float(rb[i0 + 1]), CROSS_SCALE, __uint_as_float(ra[i0 + 1])) + bv[2 * e + 1], 0.f);
            uint32_t h2;
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(f1), "f"(f0));
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
            const float d0 = (f0 - hf.x) * 2048.0f, d1 = (f1 - hf.y) * 2048.0f;
            uint32_t l2;
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(d1), "f"(d0));
            hw[e] = h2; lw[e] = l2;
          }
          const int cpos = ((ch * 4 + cc) ^ x7) << 4;
          *reinterpret_cast<uint4*>(p0 + cpos) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(p1 + cpos) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(bar(B_HSFULL + b));
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// 2D fp16 row-major [rows][cols halfs], box = [box_rows][64 halfs] (128 bytes: one swizzle row)
static int make_map_f16(CUtensorMap* map, const __half* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return AGH_E_CUDA;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * sizeof(__half)};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp16) failed (%d) rows=%llu cols=%llu", (int)rc, (unsigned long long)rows, (unsigned long long)cols);
    return AGH_E_CUDA;
  }
  return AGH_OK;
}

#endif  // AGH_TC_F16

}  // namespace tc

size_t ffn_tc_scratch_bytes() {
#if AGH_TC_F16
  return (size_t)LAYERS * 4 * FF * D * sizeof(__half);     // per layer: w1s [1024][128] + w2s [256][512]
#else
  return 0;
#endif
}

bool ffn_tc_available() { return AGH_TC_F16 != 0; }

int ffn_tc_split_weights(const float* wblob, void* scratch, cudaStream_t st) {
#if AGH_TC_F16
  __half* base = reinterpret_cast<__half*>(scratch);
  for (int l = 0; l < LAYERS; ++l) {
    const float* tw = wblob + W_TC0 + (size_t)l * T_SIZE;
    __half* w1s = base + (size_t)l * 4 * FF * D;
    tc::split_ffn_weights_kernel<<<(2 * FF * D + 255) / 256, 256, 0, st>>>(tw + T_W1, tw + T_W2, w1s, w1s + 2 * FF * D);
    AGH_LAUNCH_CHECK();
  }
  return AGH_OK;
#else
  (void)wblob; (void)scratch; (void)st;
  return AGH_E_UNSUPPORTED;
#endif
}

int launch_ffn_tc(const float* x, const void* scratch, int layer, const float* b1, const float* b2, const float* bn_s,
                  const float* bn_t, float* y, int M, cudaStream_t st) {
#if AGH_TC_F16
  using namespace tc;
  const __half* w1s = reinterpret_cast<const __half*>(scratch) + (size_t)layer * 4 * FF * D;
  const __half* w2s = w1s + 2 * FF * D;
  CUtensorMap tmX, tmW1, tmW2;
  int rc = make_map(&tmX, x, (uint64_t)M, (uint64_t)D, (uint64_t)D);
  if (rc) return rc;
  rc = make_map_f16(&tmW1, w1s, 2 * FF, D, ffn::HC);
  if (rc) return rc;
  rc = make_map_f16(&tmW2, w2s, 2 * D, FF, BM);
  if (rc) return rc;
  FfnArgs a{x, y, b1, b2, bn_s, bn_t, M};
  const int m_tiles = ceil_div(M, BM);
  const int grid = m_tiles < sm_count() ? m_tiles : sm_count();
  AGH_CUDA(cudaFuncSetAttribute(ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ffn::SMEM));
  ffn_tc_kernel<<<grid, ffn::NT, ffn::SMEM, st>>>(tmX, tmW1, tmW2, a);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
#else
  return AGH_E_UNSUPPORTED;
#endif
}

}  // namespace agh
