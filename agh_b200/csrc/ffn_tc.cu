// Fused feed-forward sublayer of the encoder on tcgen05:   y = BN( x + W2 relu(W1 x + b1) + b2 )   (graph_encoder.py:159-172,
// eval-mode BatchNorm folded into scale / shift), one launch per layer.  The [M][512] hidden activations never leave
// the SM: per 128-row tile they are produced 64 columns at a time in TMEM, read back by the H-epilogue warps (bias, ReLU),
// split into the two FP16 terms and written BACK TO TMEM as the A operand of the second contraction.
// The unfused pair (gemm_tc.cu: FF1 bias+ReLU -> HBM -> FF2 residual+BN) moves 0.52 + 2.0 GB and 2.6 + 0.5 GB per launch
// at M = 1.01 M rows and is bound by HBM; this kernel moves x in and y out (1.0 GB).
//
// Same two-term operand split as gemm_tc.cu (x = hi + lo'/2048, three MMAs per product, cross terms in their own
// accumulator); the weights are split ONCE per encoder call into FP16 hi / lo' matrices (split_ffn_weights_kernel) and
// streamed per tile by TMA straight into the UMMA layout -- no splitter pass over them.
//
// What bounded the first version was the shared-memory bandwidth the tensor core needs for its operands (a 128 x 128 x 16
// FP16 MMA with both operands in shared memory reads 8 KB = 64 clocks at 128 B/clk, exactly its math time, and the split
// form reads every block two or three times; profiles/r2_ffn_fused.md has the timeline of that all-shared-memory version:
// 12.3 us per tile).  So the operands that are reused most sit in TMEM and are fed to tcgen05.mma as the A operand from
// there: the hi half of the x tile (used by two of the three products of every k-step of all eight A ops) and both halves of
// the current H chunk.  The shipped kernel keeps the tensor pipe busy 54 % of the time (ncu; 6.6 us of MMA
// time per 11.9 us tile): with TMEM full there is one H accumulator and one H-chunk buffer, so A_j -> H epilogue -> B_j is a chain.
//
//   warp 0      : TMA producer  - the raw x tile (4 fp32 boxes, a tile ahead); weight stages (32 KB: the W1 rows of one
//                                 64-wide hidden chunk, or the W2 columns of one) through a ring of 4
//   warp 1      : MMA issuer    - per tile: A_0, then for j = 1..7 { A_j, B_{j-1} }, then B_7, where
//                                 A_j: H[128 x 64] = x W1_j^T (K = 128, 16 MMAs),  B_j: Y += H_j W2_j^T (K = 64, 8 MMAs)
//   warps 2..9  : Y epilogue    - tcgen05.ld {main, cross} -> + b2 + residual, BatchNorm -> HBM
//   warps 10..17: H epilogue    - tcgen05.ld {main, cross} -> + b1, ReLU -> FP16 hi / lo' -> tcgen05.st; between tiles they
//                                 split the next raw x tile (one thread per row and 64 k: hi -> TMEM, lo' -> smem)
// TMEM (512 columns): Y main 128 | Y cross 128 | H main 64 | H cross 64 | H chunk hi 32 | lo' 32 | x hi 64.
#include "common.cuh"
#include "gemm.cuh"
#include "tc.cuh"

namespace agh {

namespace tc {

#if AGH_TC_F16

namespace ffn {
constexpr int HC = 64;                       // hidden columns per chunk
constexpr int NCH = FF / HC;                 // 8 chunks per tile
constexpr int WST = 4;                       // weight ring stages
constexpr int WSTAGE = 2 * BLK;              // 32 KB
// shared memory: raw x tile (2 stages of two [128 x 32] fp32 boxes) | x lo' (2 K-major blocks of [128 x 64] halfs) | weight ring
constexpr int XRAW_OFF = 0, XLO_OFF = 4 * BLK, W_OFF = 6 * BLK, BAR_OFF = W_OFF + WST * WSTAGE;
// barriers
constexpr int B_XRAW = 0, B_XRAWEMPTY = 2, B_XSPLIT = 3, B_XEMPTY = 5, B_WFULL = 6, B_WEMPTY = B_WFULL + WST, B_HFULL = B_WEMPTY + WST,
              B_HEMPTY = B_HFULL + 1, B_HSFULL = B_HEMPTY + 1, B_HSEMPTY = B_HSFULL + 1, B_YFULL = B_HSEMPTY + 1, B_YEMPTY = B_YFULL + 1,
              NBARS = B_YEMPTY + 1;
constexpr int VEC_OFF = BAR_OFF + 256;       // b2 | bn_s | bn_t
constexpr int SMEM = VEC_OFF + 3 * D * 4 + 1024;
static_assert(8 * NBARS + 16 <= 256, "barrier block");
static_assert(SMEM <= 232448, "shared memory budget");
// TMEM columns
constexpr int T_YMAIN = 0, T_YCROSS = 128, T_HMAIN = 256, T_HCROSS = 320, T_HSHI = 384, T_HSLO = 416, T_XHI = 448;
constexpr int NT = 32 * 18;
}  // namespace ffn

// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (128 rows = lanes, K halfs packed two per 32-bit column) read from TMEM
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// two fp32 -> packed FP16 hi pair and scaled-remainder pair (element 0 in the low half)
__device__ __forceinline__ void split_pair(float f0, float f1, uint32_t& h2, uint32_t& l2) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(f1), "f"(f0));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
  const float d0 = (f0 - hf.x) * 2048.0f, d1 = (f1 - hf.y) * 2048.0f;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(d1), "f"(d0));
}

// -DAGH_FFN_TRACE: CTA 0 of the first large launches stamps %globaltimer at the pipeline hand-offs and prints them at the
// end (tools/ffn_timeline.py).  Off in the product build.
#ifdef AGH_FFN_TRACE
constexpr int FTR_CAP = 128;
__device__ unsigned long long g_ftrace[14][FTR_CAP];
__device__ int g_ftrace_launches = 0;
__device__ __forceinline__ unsigned long long ftime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FTR(role, idx) do { if (blockIdx.x == 0 && (int)(idx) < FTR_CAP) g_ftrace[role][idx] = ftime(); } while (0)
#else
#define FTR(role, idx) do { } while (0)
#endif

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

__global__ void __launch_bounds__(ffn::NT, 1)      // 18 warps: five on one scheduler -> 96 registers per thread
ffn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const FfnArgs g) {
  using namespace ffn;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* xraw = smem + XRAW_OFF;        // stage kb (k 64 kb .. +63): [box k 0..31 | box k 32..63], 16 KB each, as TMA wrote them
  unsigned char* xlo = smem + XLO_OFF;          // block kb: K-major SWIZZLE_128B [128 rows][64 halfs] of the scaled remainders
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
      mbar_init(bar(B_XSPLIT + s), 4);          // one arrival per splitter warp of the stage
    }
    mbar_init(bar(B_XRAWEMPTY), 8);
    mbar_init(bar(B_XEMPTY), 1);
    for (int s = 0; s < WST; ++s) {
      mbar_init(bar(B_WFULL + s), 1);
      mbar_init(bar(B_WEMPTY + s), 1);
    }
    mbar_init(bar(B_HFULL), 1);
    mbar_init(bar(B_HEMPTY), 8);
    mbar_init(bar(B_HSFULL), 8);
    mbar_init(bar(B_HSEMPTY), 1);
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
      auto load_x = [&](int mt, uint32_t tl) {
        mbar_wait(bar(B_XRAWEMPTY), (tl & 1) ^ 1);                // the splitters have read the previous raw tile
        for (int s = 0; s < 2; ++s) {
          mbar_expect_tx(bar(B_XRAW + s), 2 * BLK);
          tma_load_2d(smem_u32(xraw + s * 2 * BLK), &tmX, bar(B_XRAW + s), s * 64, mt * BM);
          tma_load_2d(smem_u32(xraw + s * 2 * BLK + BLK), &tmX, bar(B_XRAW + s), s * 64 + BK, mt * BM);
        }
      };
      auto load_w1 = [&](int j) {          // rows 64 j .. +63 of W1: [hi k 0..63 | lo' k 0..63 | hi k 64..127 | lo' k 64..127], 8 KB each
        const int s = wit % WST;
        mbar_wait(bar(B_WEMPTY + s), ((wit / WST) & 1) ^ 1);
        FTR(0, wit);
        const uint32_t dst = smem_u32(wr + s * WSTAGE);
        mbar_expect_tx(bar(B_WFULL + s), WSTAGE);
        tma_load_2d(dst, &tmW1, bar(B_WFULL + s), 0, j * HC);
        tma_load_2d(dst + BLK / 2, &tmW1, bar(B_WFULL + s), 0, FF + j * HC);
        tma_load_2d(dst + BLK, &tmW1, bar(B_WFULL + s), 64, j * HC);
        tma_load_2d(dst + BLK + BLK / 2, &tmW1, bar(B_WFULL + s), 64, FF + j * HC);
        ++wit;
      };
      auto load_w2 = [&](int j) {          // columns 64 j .. +63 of W2, all 128 rows: [hi 16 KB | lo' 16 KB]
        const int s = wit % WST;
        mbar_wait(bar(B_WEMPTY + s), ((wit / WST) & 1) ^ 1);
        FTR(0, wit);
        const uint32_t dst = smem_u32(wr + s * WSTAGE);
        mbar_expect_tx(bar(B_WFULL + s), WSTAGE);
        tma_load_2d(dst, &tmW2, bar(B_WFULL + s), j * HC, 0);
        tma_load_2d(dst + BLK, &tmW2, bar(B_WFULL + s), j * HC, D);
        ++wit;
      };
      if ((int)blockIdx.x < num_mt) load_x(blockIdx.x, 0);
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
        load_w1(0);
        load_w1(1);
        if (mt + (int)gridDim.x < num_mt) load_x(mt + gridDim.x, tile + 1);      // next tile's raw x, a whole tile ahead
        load_w2(0);
        for (int j = 2; j < NCH; ++j) {
          load_w1(j);
          load_w2(j - 1);
        }
        load_w2(NCH - 1);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // One MMA with the hi operand from TMEM covers the main AND one cross term: its B operand is [w_hi rows ; w_lo' rows]
    // (adjacent in the stage) and its N twice the output width, so that D = [main | cross] (adjacent TMEM columns); the
    // other cross term (lo' operand x w_hi) is a second, narrower MMA onto the cross columns.  Two wide instructions per
    // k-step instead of three narrow ones: the tensor pipe shows ~26 clocks of fixed cost per instruction on top of N / 2.
    constexpr uint32_t idesc_a2 = umma_idesc(BM, 2 * HC), idesc_a1 = umma_idesc(BM, HC);
    constexpr uint32_t idesc_b2 = umma_idesc(BM, 2 * BN), idesc_b1 = umma_idesc(BM, BN);
    uint32_t wit = 0, tile = 0, chunk = 0;      // chunk: running count of A ops (H accumulator phases)
    const bool leader = elect_one();            // ONE thread issues every MMA and every commit (a commit covers the issuing thread's MMAs)
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
      auto op_a = [&](int j) {
        const int s = wit % WST;
        if (lane == 0) FTR(1, wit);
        mbar_wait(bar(B_HEMPTY), (chunk & 1) ^ 1);                 // H accumulator drained by the H epilogue
        if (lane == 0) FTR(2, wit);
        mbar_wait(bar(B_WFULL + s), (wit / WST) & 1);
        if (lane == 0) FTR(3, wit);
        if (j == 0) mbar_wait(bar(B_XSPLIT), tile & 1);
        if (lane == 0) FTR(4, wit);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t wb = smem_u32(wr + s * WSTAGE);
        for (int kb = 0; kb < 2; ++kb) {
          if (j == 0 && kb == 1) {
            mbar_wait(bar(B_XSPLIT + 1), tile & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          if (leader) {
            const uint64_t da_lo = umma_desc(smem_u32(xlo + kb * BLK));
            const uint64_t dw = umma_desc(wb + kb * BLK);                        // 64 rows of w_hi, then 64 rows of w_lo'
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              const uint32_t a_hi = tmem_base + T_XHI + kb * 32 + k * 8;          // 16 halfs = 8 columns per k-step
              umma_ts(tmem_base + T_HMAIN, a_hi, dw + adv, idesc_a2, (kb | k) != 0);     // [x_hi w_hi | x_hi w_lo']
              umma_tf32(tmem_base + T_HCROSS, da_lo + adv, dw + adv, idesc_a1, 1);      // x_lo' w_hi
            }
          }
          __syncwarp();
        }
        if (leader) {
          umma_commit(bar(B_WEMPTY + s));
          umma_commit(bar(B_HFULL));
          if (j == NCH - 1) umma_commit(bar(B_XEMPTY));           // the x tile (TMEM hi, smem lo') is no longer read
        }
        __syncwarp();
        if (lane == 0) FTR(5, wit);
        ++wit; ++chunk;
      };
      auto op_b = [&](int j) {
        const int s = wit % WST;
        const uint32_t c = tile * NCH + j;                         // running chunk index, as counted by the H epilogue
        if (lane == 0) FTR(1, wit);
        if (j == 0) mbar_wait(bar(B_YEMPTY), (tile & 1) ^ 1);      // Y accumulators drained by the Y epilogue
        if (lane == 0) FTR(2, wit);
        mbar_wait(bar(B_WFULL + s), (wit / WST) & 1);
        if (lane == 0) FTR(3, wit);
        mbar_wait(bar(B_HSFULL), c & 1);
        if (lane == 0) FTR(4, wit);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) {
          const uint32_t wb = smem_u32(wr + s * WSTAGE);
          const uint64_t dw = umma_desc(wb);                                     // 128 rows of w_hi, then 128 rows of w_lo'
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            const uint32_t a_hi = tmem_base + T_HSHI + k * 8, a_lo = tmem_base + T_HSLO + k * 8;
            umma_ts(tmem_base + T_YMAIN, a_hi, dw + adv, idesc_b2, (j | k) != 0);        // [h_hi w_hi | h_hi w_lo']
            umma_ts(tmem_base + T_YCROSS, a_lo, dw + adv, idesc_b1, 1);                  // h_lo' w_hi
          }
          umma_commit(bar(B_WEMPTY + s));
          umma_commit(bar(B_HSEMPTY));
          if (j == NCH - 1) umma_commit(bar(B_YFULL));
        }
        __syncwarp();
        if (lane == 0) FTR(5, wit);
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
    uint32_t tile = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
      const int row0 = mt * BM + q * 32;
      const int c_first = chalf * 64;
      mbar_wait(bar(B_YFULL), tile & 1);
      if (t == 0) FTR(11, tile);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16);
      // Drain this warp's 32 lanes x 64 columns of both accumulators into registers FIRST and hand the accumulators back:
      // the next tile's first B op waits for this, everything below overlaps with its MMAs.
      float r[64];                                                // [column chunk][lane half][16]
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t ra[16], rb[16];
          const uint32_t ta = t_acc + ((uint32_t)(hh * 16) << 16) + (uint32_t)(c_first + cc * 32);
          tmem_ld16x256b_x4(ta + T_YMAIN, ra);
          tmem_ld16x256b_x4(ta + T_YCROSS, rb);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) r[cc * 32 + hh * 16 + i] = fmaf(__uint_as_float(rb[i]), CROSS_SCALE, __uint_as_float(ra[i]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_YEMPTY));
      if (t == 0) FTR(12, tile);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = c_first + cc * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int cl = c0 + 8 * k + 2 * t4;
          const float2 bias = *reinterpret_cast<const float2*>(s_vec + cl);
          const float2 sc = *reinterpret_cast<const float2*>(s_vec + D + cl);
          const float2 sh = *reinterpret_cast<const float2*>(s_vec + 2 * D + cl);
#pragma unroll
          for (int rq = 0; rq < 4; ++rq) {
            const int row = row0 + rq * 8 + g8;
            const int ri = cc * 32 + (rq >> 1) * 16 + k * 4 + (rq & 1) * 2;
            if (row < g.M) {
              const float2 rs = __ldg(reinterpret_cast<const float2*>(g.x + (size_t)row * D + cl));      // residual (L2: the tile was just loaded)
              float2 v;
              v.x = fmaf(rs.x + (r[ri] + bias.x), sc.x, sh.x);
              v.y = fmaf(rs.y + (r[ri + 1] + bias.y), sc.y, sh.y);
              *reinterpret_cast<float2*>(g.y + (size_t)row * D + cl) = v;
            }
          }
        }
      }
      if (t == 0) FTR(13, tile);
    }
  } else {
    // ================= H epilogue (warps 10..17): TMEM lane quarter = warp % 4, column half = (warp - 10) / 4 =================
    const int q = warp & 3, ch = (warp - 10) >> 2;
    // Split of the raw x tile: this thread owns row 32 q + lane and k = 64 ch .. +63.  hi -> TMEM (two halfs per 32-bit
    // column, the A-operand layout of tcgen05.mma), lo' -> the K-major SWIZZLE_128B block of the stage.
    auto split_x = [&](uint32_t tile) {
      const int row = q * 32 + lane, x7 = row & 7;
      const unsigned char* p = xraw + ch * 2 * BLK + row * 128;
      mbar_wait(bar(B_XRAW + ch), tile & 1);
      mbar_wait(bar(B_XEMPTY), (tile & 1) ^ 1);                  // the previous tile's last A op has retired
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + T_XHI + ch * 32;
      unsigned char* pl = xlo + ch * BLK + row * 128;
#pragma unroll
      for (int bx = 0; bx < 2; ++bx) {                           // raw box bx: k = 64 ch + 32 bx .. +31
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(p + bx * BLK + ((c ^ x7) << 4));
          split_pair(v.x, v.y, hw[2 * c], lw[2 * c]);
          split_pair(v.z, v.w, hw[2 * c + 1], lw[2 * c + 1]);
        }
        tmem_st16(ta + bx * 16, hw);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(pl + (((bx * 4 + c) ^ x7) << 4)) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_XRAWEMPTY));              // raw tile read: the producer may fetch the next one
      if (lane == 0) mbar_arrive(bar(B_XSPLIT + ch));
    };
    uint32_t chunk = 0, tile = 0;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    if ((int)blockIdx.x < num_mt) split_x(0);
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
#pragma unroll 1
      for (int j = 0; j < NCH; ++j, ++chunk) {
        const float4* b1p = reinterpret_cast<const float4*>(g.b1 + j * HC + ch * 32);      // warp-uniform: L1 broadcast
        mbar_wait(bar(B_HFULL), chunk & 1);
        if (threadIdx.x == 320) FTR(6, chunk);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t ra[32], rb[32];
        tmem_ld32(tq + T_HMAIN + ch * 32, ra);
        tmem_ld32(tq + T_HCROSS + ch * 32, rb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_HEMPTY));
        if (threadIdx.x == 320) FTR(7, chunk);
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 ba = __ldg(b1p + 2 * cc), bc = __ldg(b1p + 2 * cc + 1);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bc.x, bc.y, bc.z, bc.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i0 = cc * 8 + 2 * e;
            const float f0 = fmaxf(fmaf(__uint_as_float(rb[i0]), CROSS_SCALE, __uint_as_float(ra[i0])) + bv[2 * e], 0.f);
            const float f1 = fmaxf(fmaf(__uint_as_float(rb[i0 + 1]), CROSS_SCALE, __uint_as_float(ra[i0 + 1])) + bv[2 * e + 1], 0.f);
            split_pair(f0, f1, hw[cc * 4 + e], lw[cc * 4 + e]);
          }
        }
        mbar_wait(bar(B_HSEMPTY), (chunk & 1) ^ 1);               // the B op of the previous chunk has retired
        if (threadIdx.x == 320) FTR(8, chunk);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_st16(tq + T_HSHI + ch * 16, hw);
        tmem_st16(tq + T_HSLO + ch * 16, lw);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_HSFULL));
        if (threadIdx.x == 320) FTR(9, chunk);
      }
      // next tile's x: its last reader (A_7) retired before this tile's last H chunk was ready, so this does not wait;
      // B_6 and B_7 keep the tensor pipe busy meanwhile
      if (mt + (int)gridDim.x < num_mt) split_x(tile + 1);
      if (threadIdx.x == 320) FTR(10, tile);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
#ifdef AGH_FFN_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0 && g.M > 500000) {
    const int l = atomicAdd(&g_ftrace_launches, 1);
    if (l < 3) {
      const unsigned long long base = g_ftrace[1][0];
      for (int role = 0; role < 14; ++role) {
        const int cnt = role < 6 ? FTR_CAP : (role < 10 ? FTR_CAP / 2 : FTR_CAP / 16);
        for (int i = 0; i < cnt; ++i) printf("FTRACE %d %d %d %lld\n", l, role, i, (long long)(g_ftrace[role][i] - base));
      }
    }
  }
#endif
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
