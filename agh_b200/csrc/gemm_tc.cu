// Tensor-core GEMM for the encoder / precompute contractions:  C[M,Nc] = epi(A[M,K] @ W[Nc,K]^T), fp32 in,
// fp32 out, computed on the 5th-generation tensor cores (tcgen05, accumulators in TMEM) with a two-term operand split
//      a = a_hi + a_lo,  w = w_hi + w_lo
//      a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w_hi            (dropped term ~2^-22 relative)
// so the result keeps fp32-level accuracy (the 1e-4 log-prob tolerance rules out plain tf32 / bf16 / fp16 operands).
// Shipped form (AGH_TC_F16 = 1): hi = rn_f16(x), lo' = rn_f16((x - hi) * 2048), kind::f16 MMAs of K = 16, the cross
// accumulator scaled by 2^-11 in the epilogue; AGH_TC_F16 = 0: hi = rn_tf32(x), lo = rn_tf32(x - hi), kind::tf32 MMAs
// of K = 8 (the form the scheme was first validated with).  Stage sizes below are those of the FP16 form.
//
// Persistent, warp-specialised kernel: each CTA owns one 128-column slab of W and loops over 128-row tiles of A.
//   warp 0      : TMA producer  - cp.async.bulk.tensor (SWIZZLE_128B) boxes of [128 x 32] fp32, two per operand and
//                                 stage (K = 64), into a ring of 4 (K = 128 kernels) or 3 (K = 512) stages
//   warp 1      : MMA issuer    - one elected lane issues 3 tcgen05.mma per K=16 step; tcgen05.commit
//                                 releases the smem stage / hands the TMEM accumulator to the epilogue
//   warps 2..9  : splitters     - two threads per row (one per raw box) rewrite the two landed fp32 boxes of a stage IN
//                                 PLACE as the K-major SWIZZLE_128B fp16 blocks of hi and lo' ([128 x 64 halfs] each),
//                                 fence.proxy.async  (TF32 build: warps 2..5, epilogue 6..13)
//   warps 10..17: epilogue      - tcgen05.ld -> bias / ReLU / residual + BatchNorm -> HBM, overlapped
//                                 with the next tile's main loop through a double-buffered TMEM accumulator
// K = 128 (QKV, out-projection, FF1, decoder projection): the W slab (hi + lo', 64 KB) is loaded and split ONCE
// per CTA and stays resident; only A streams.  K = 512 (FF2): W blocks stream with A.
// Wide outputs (QKV, FF1, decoder projection) are written slab-major -- one [M][128] matrix per 128-column slab, so a
// CTA's tile is a contiguous 64 KB block -- and FF2 reads its A operand from that form (GemmArgs::c_slab / a_slab).
//
// The tensor core adds into its fp32 accumulator with truncation, so the error grows with the number of sequential
// accumulations: the small cross terms go to their own accumulator (and, for K = 512, the hi.hi terms are spread
// round-robin over three), and the epilogue adds the accumulators in fp32 round-to-nearest.
//
// Replaces the cuBLAS calls behind graph_encoder.py:83-86,106-109,165-168 and attention_model.py:405-406.
#include "common.cuh"
#include "gemm.cuh"
#include "tc.cuh"

namespace agh {

namespace tc {

// WRES = true : K == 128, W slab resident (hi 64 KB + lo 64 KB), stage = A hi|lo (32 KB), 2 TMEM buffers of {main, cross}
// WRES = false: K % 32 == 0, stage = A hi|lo + W hi|lo (64 KB), 1 TMEM buffer of {3 main, cross}
template <bool WRES>
struct Cfg {
  static constexpr int STAGES = WRES ? STAGES_WRES : STAGES_STREAM;
  static constexpr int STAGE_BYTES = WRES ? 2 * BLK : 4 * BLK;      // F16: the same bytes hold twice the K extent
  static constexpr int WRES_BYTES = WRES ? (F16 ? 4 : 8) * BLK : 0;
  static constexpr int NACC = WRES ? 1 : 3;
#ifndef AGH_TC_NBUF
#define AGH_TC_NBUF 2
#endif
  static constexpr int NBUF = WRES ? AGH_TC_NBUF : 1;
  static constexpr int NBARS = 3 * STAGES + 2 * NBUF + 2;
  static constexpr int VEC_OFF = (8 * NBARS + 16 + 15) & ~15;      // per-column epilogue vectors: bias, bn scale, bn shift
  static constexpr int SMEM = WRES_BYTES + STAGES * STAGE_BYTES + 1024 /*align*/ + VEC_OFF + 3 * BN * 4;
};

// Which epilogues go through the in-register 32x32 transpose (coalesced 128-byte rows) instead of the row-wise
// 16-byte accesses: bit EPI of the mask.  Overridable at compile time for A/B measurements
// (AGH_NVCC_EXTRA=-DAGH_TC_TRANSPOSE_MASK=..).  Measured per launch at M = 1.01 M rows with the eight epilogue
// warps (transposed vs row-wise): QKV/plain 612 vs ~1000 us, out-proj 508 vs ~630, FF2 1010 vs ~1250 (residual+BN
// with the residual prefetched), FF1/bias+ReLU 820 vs 1200 -> every epilogue is transposed (the decoder projection
// writes four plain [M][128] matrices; its former swizzled 16-byte scatter was 1250 us row-wise, 1690 transposed).
// (With four epilogue warps the shuffles made FF1 slower transposed.)
#ifndef AGH_TC_TRANSPOSE_MASK
#define AGH_TC_TRANSPOSE_MASK 0xf
#endif
__host__ __device__ constexpr bool epi_transposed(int epi) { return (AGH_TC_TRANSPOSE_MASK >> epi) & 1; }
// Bit EPI set: read the accumulators in the 16x256b fragment distribution and access global memory straight from it
// (whole 32-byte sectors, no shuffles) instead of the 32x32b row distribution with the epilogues above.  Measured per
// launch at M = 1.01 M rows (fragment vs transposed): out-proj 292 vs 503 us, FF2 968 vs 998 (residual + BN: the
// residual is read in the same distribution and the 32x32 transposes go away); QKV 683 vs 608, FF1 842 vs 812, decoder
// projection 860 vs 830 (store-heavy shapes: four times as many 32-byte store requests as 128-byte rows) -> only the
// residual epilogue uses it.
// Bit EPI set: the MMA computes the TRANSPOSED tile, D[c][r] = sum_k W[c][k] A[r][k]: the weight slab is the "A" operand
// (its 128 rows = output columns -> TMEM lanes) and the activation tile the "B" operand (its 128 rows -> TMEM columns).
// Both are K-major SWIZZLE_128B blocks, so this is only an exchange of the two shared-memory descriptors -- but a
// 32x32b TMEM load then hands every thread ONE output column and 32 consecutive rows, i.e. the register layout the
// in-register transpose used to build with 80 shuffles + 160 selects per chunk: every store (and residual load) of a
// warp is one coalesced 128-byte row segment for free.  Measured per launch at M = 1.01 M rows (swapped vs the best of
// the other epilogues): decoder projection 736 vs 830 us, but QKV 646 vs 608, FF1 923 vs 812, out-proj 377 vs 292,
// FF2 1024 vs 968 -- the shuffles were not what bounds those shapes -- so only the decoder projection uses it.
#ifndef AGH_TC_SWAP_MASK
#define AGH_TC_SWAP_MASK 0x1b
#endif
__host__ __device__ constexpr bool epi_swap(int epi) { return (AGH_TC_SWAP_MASK >> epi) & 1; }
#ifndef AGH_TC_FRAG_MASK
#define AGH_TC_FRAG_MASK 0x4
#endif
__host__ __device__ constexpr bool epi_frag(int epi) { return (AGH_TC_FRAG_MASK >> epi) & 1; }

// -DAGH_TC_TRACE: CTA (0,0) of the first large launches stamps %globaltimer at every pipeline hand-off and prints the
// per-role timeline at the end (tools/gemm_timeline.py turns it into per-stage durations).  Off in the product build.
#ifdef AGH_TC_TRACE
constexpr int TR_CAP = 96;
__device__ unsigned long long g_trace[8][TR_CAP];
__device__ int g_trace_launches = 0;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TR(role, idx) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (int)(idx) < TR_CAP) g_trace[role][idx] = gtime(); } while (0)
#else
#define TR(role, idx) do { } while (0)
#endif

// CSLAB: the output is slab-major (Nc/128 separate [M][128] matrices).  A template parameter, like the fixed pitch of
// the residual epilogue below, so that the row pitch is a compile-time 128: the 32 stores of a chunk then share one
// base register with immediate offsets instead of 32 64-bit multiply-adds (the epilogue warps are what bounds the tile
// period of the K = 128 shapes).
template <int EPI, bool WRES, bool CSLAB>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmArgs g) {
  using C = Cfg<WRES>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ unsigned char smem_raw[];
  // swizzle-128B tiles need 1024-byte alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* wres = smem;                                   // [4 k-blocks hi][4 k-blocks lo] when WRES
  unsigned char* ring = smem + C::WRES_BYTES;
  unsigned char* bars = ring + STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = smem_u32(bars);
  auto bar_raw = [&](int s) { return bar_base + 8 * s; };                       // TMA -> splitters
  auto bar_split = [&](int s) { return bar_base + 8 * (STAGES + s); };          // splitters -> MMA
  auto bar_empty = [&](int s) { return bar_base + 8 * (2 * STAGES + s); };      // MMA -> TMA
  auto bar_tfull = [&](int b) { return bar_base + 8 * (3 * STAGES + b); };      // MMA -> epilogue
  auto bar_tempty = [&](int b) { return bar_base + 8 * (3 * STAGES + C::NBUF + b); };   // epilogue -> MMA
  const uint32_t bar_wraw = bar_base + 8 * (3 * STAGES + 2 * C::NBUF);
  const uint32_t bar_wsplit = bar_wraw + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 * C::NBARS);
  float* s_vec = reinterpret_cast<float*>(bars + C::VEC_OFF);   // [bias | bn_s | bn_t] of this CTA's 128 columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const int nkb = g.K / KST;
  const int num_mt = (g.M + BM - 1) / BM;

  // The epilogue's per-column vectors are constant for the CTA: staged once, so that no global-memory latency sits
  // between the accumulator read and the stores of every tile.
  if (threadIdx.x >= 64 && threadIdx.x < 64 + BN) {
    const int t = threadIdx.x - 64, c = n0 + t;
    s_vec[t] = g.bias != nullptr ? g.bias[c] : 0.f;
    s_vec[BN + t] = g.bn_s != nullptr ? g.bn_s[c] : 1.f;
    s_vec[2 * BN + t] = g.bn_t != nullptr ? g.bn_t[c] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_raw(s), 1);
      mbar_init(bar_split(s), SPLIT_WARPS * 32);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < C::NBUF; ++b) {
      mbar_init(bar_tfull(b), 1);
      mbar_init(bar_tempty(b), EPI_WARPS);
    }
    mbar_init(bar_wraw, 1);
    mbar_init(bar_wsplit, SPLIT_WARPS * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation (whole warp): all 512 columns, one CTA per SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  constexpr int ACC_COLS = (C::NACC + 1) * BN;                  // TMEM columns of one accumulator set

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      if (WRES) {
        mbar_expect_tx(bar_wraw, 4 * BLK);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem_u32(wres + kb * BLK), &tmW, bar_wraw, kb * BK, n0);
      }
      uint32_t it = 0;
      // A coordinates of the box at k0: row-major (k0, row); slab-major input (a_slab: K/128 matrices [M][128] stacked
      // in one [K/128 * M][128] map) (k0 % 128, (k0 / 128) * M + row) -- rows past M of a slab then come from the next
      // slab, which only feeds output rows >= M that are never stored
      auto load_a = [&](uint32_t dst, uint32_t bar, int k0, int row) {
        if (g.a_slab) tma_load_2d(dst, &tmA, bar, k0 & (BN - 1), (k0 >> 7) * g.M + row);
        else tma_load_2d(dst, &tmA, bar, k0, row);
      };
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(bar_empty(s), ((it / STAGES) & 1) ^ 1);
          TR(0, it);
          unsigned char* st = ring + s * C::STAGE_BYTES;
          if (F16) {      // two raw boxes per operand: k = kb*64 .. +31 and +32 .. +63
            mbar_expect_tx(bar_raw(s), WRES ? 2 * BLK : 4 * BLK);
            load_a(smem_u32(st), bar_raw(s), kb * KST, mt * BM);
            load_a(smem_u32(st + BLK), bar_raw(s), kb * KST + BK, mt * BM);
            if (!WRES) {
              tma_load_2d(smem_u32(st + 2 * BLK), &tmW, bar_raw(s), kb * KST, n0);
              tma_load_2d(smem_u32(st + 3 * BLK), &tmW, bar_raw(s), kb * KST + BK, n0);
            }
          } else {
            mbar_expect_tx(bar_raw(s), WRES ? BLK : 2 * BLK);
            load_a(smem_u32(st), bar_raw(s), kb * BK, mt * BM);
            if (!WRES) tma_load_2d(smem_u32(st + 2 * BLK), &tmW, bar_raw(s), kb * BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc(BM, BN);
    if (WRES) {
      mbar_wait(bar_wsplit, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    uint32_t it = 0, tile = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
      const int ab = tile % C::NBUF;
      mbar_wait(bar_tempty(ab), ((tile / C::NBUF) & 1) ^ 1);       // epilogue has drained this accumulator set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_acc = tmem_base + ab * ACC_COLS;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(bar_split(s), (it / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          TR(3, it);
          const uint32_t st = smem_u32(ring + s * C::STAGE_BYTES);
          const uint64_t da_hi = umma_desc(st), da_lo = umma_desc(st + BLK);
          // resident W: TF32 [4 hi blocks][4 lo blocks]; F16 [hi, lo'] pairs, one pair per 64 of K
          const uint32_t wb = WRES ? smem_u32(wres + (F16 ? 2 * kb : kb) * BLK) : st + 2 * BLK;
          const uint64_t dw_hi = umma_desc(wb), dw_lo = umma_desc(wb + ((WRES && !F16) ? 4 * BLK : BLK));
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 8 tf32 = 16 fp16 = 32 bytes inside the swizzle atom
            const int ks = kb * (BK / 8) + k;
            const uint32_t t_cross = t_acc + C::NACC * BN;
            const uint32_t t_main = t_acc + (ks % C::NACC) * BN;
            if (epi_swap(EPI)) {   // D^T: weights are the M-side operand
              umma_tf32(t_cross, dw_hi + adv, da_lo + adv, idesc, ks != 0);
              umma_tf32(t_cross, dw_lo + adv, da_hi + adv, idesc, 1);
              umma_tf32(t_main, dw_hi + adv, da_hi + adv, idesc, ks >= C::NACC);
            } else {
              umma_tf32(t_cross, da_lo + adv, dw_hi + adv, idesc, ks != 0);
              umma_tf32(t_cross, da_hi + adv, dw_lo + adv, idesc, 1);
              umma_tf32(t_main, da_hi + adv, dw_hi + adv, idesc, ks >= C::NACC);
            }
          }
          umma_commit(bar_empty(s));                             // frees the stage when these MMAs retire
          TR(4, it);
          if (kb == nkb - 1) umma_commit(bar_tfull(ab));         // accumulator set complete
        }
        __syncwarp();
      }
    }
  } else if (warp < FIRST_EPI) {
    // ================= splitters (warps 2 .. FIRST_EPI-1) =================
    const int t = threadIdx.x - 64;
    if (WRES) {
      mbar_wait(bar_wraw, 0);
      if (F16) {
        split_f16_half(wres, wres + BLK, t);
        split_f16_half(wres + 2 * BLK, wres + 3 * BLK, t);
      } else {
        split_block(wres, wres + 4 * BLK, 4 * BLK / 16, t);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_wsplit);
    }
    uint32_t it = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(bar_raw(s), (it / STAGES) & 1);
        if (t == 0) TR(1, it);
        unsigned char* st = ring + s * C::STAGE_BYTES;
        if (F16) {
          split_f16_half(st, st + BLK, t);
          if (!WRES) split_f16_half(st + 2 * BLK, st + 3 * BLK, t);
        } else {
          split_block(st, st + BLK, BLK / 16, t);
          if (!WRES) split_block(st + 2 * BLK, st + 3 * BLK, BLK / 16, t);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(bar_split(s));
        if (t == 0) TR(2, it);
      }
    }
  } else {
    // ================= epilogue (warps FIRST_EPI ..): TMEM lane quarter = warp % 4; the two warps of a quarter split the
    // four 32-column chunks of the tile between them =================
    const int q = warp & 3;
    const int chalf = (warp - FIRST_EPI) >> 2;                          // 0: chunks 0,1   1: chunks 2,3
    // EPI_DEC: the four 128-column slabs are four plain [M][128] matrices -- K', V, LK' (keys[0..2]) and P_sc -- so the
    // epilogue is the plain one with a per-slab base (shifted by n0 because `col` below is the GEMM's global column)
    float* Cb = g.C;
    int ldc = g.ldc;
    if (EPI == EPI_DEC) {
      Cb = (blockIdx.y < 3 ? g.keys + (size_t)blockIdx.y * g.M * D : g.psc) - n0;
      ldc = D;
    } else if (CSLAB) {         // slab-major output: this CTA's 128 columns are their own [M][128] matrix
      Cb = g.C + (size_t)blockIdx.y * g.M * BN - n0;
      ldc = BN;
    } else if (EPI == EPI_RES_BN) {
      ldc = BN;                 // residual + BatchNorm outputs and residuals are [M][128] activations (checked by the launcher)
    }
    const int ldr = (EPI == EPI_RES_BN) ? BN : g.ldr;
    constexpr int CH_PER_WARP = (BN / 32) / (EPI_WARPS / 4);
    uint32_t tile = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tile) {
      const int ab = tile % C::NBUF;
      if (epi_swap(EPI)) {
        // thread = one output column (TMEM lane), chunks of 32 consecutive rows (TMEM columns); the two warps of a lane
        // quarter split the tile's four row chunks between them
        const int cl = q * 32 + lane, col = n0 + cl;
        const int r_first = chalf * CH_PER_WARP * 32, r_last = r_first + (CH_PER_WARP - 1) * 32;
        float bias = 0.f, sc = 1.f, sh = 0.f;
        if (EPI == EPI_BIAS_RELU || ((EPI == EPI_RES_BN || EPI == EPI_TRAIN) && g.bias != nullptr)) bias = s_vec[cl];
        if (EPI == EPI_RES_BN) { sc = s_vec[BN + cl]; sh = s_vec[2 * BN + cl]; }
        float rs[32];                                           // residual of rows mt*BM + rc + i, this thread's column
        auto load_res_sw = [&](int rc) {
          const int rbase = mt * BM + rc;
          const float* src = g.res + (size_t)rbase * ldr + col;
#pragma unroll
          for (int i = 0; i < 32; ++i) rs[i] = (rbase + i < g.M) ? __ldg(src + (size_t)i * ldr) : 0.f;
        };
        if (EPI == EPI_RES_BN) load_res_sw(r_first);            // independent of the MMAs: issued before the accumulator is ready
        mbar_wait(bar_tfull(ab), (tile / C::NBUF) & 1);
        if (warp == FIRST_EPI && lane == 0) TR(5, tile);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t t_acc = tmem_base + ab * ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int rc = r_first; rc <= r_last; rc += 32) {
          float r[32];
          {   // accumulators one after the other (hi.hi partials, then the cross terms), added in fp32 round-to-nearest
            uint32_t ra[32];
            tmem_ld32(t_acc + (uint32_t)rc, ra);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(ra[i]);
#pragma unroll
            for (int a = 1; a <= C::NACC; ++a) {
              tmem_ld32(t_acc + (uint32_t)(a * BN + rc), ra);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = fmaf(__uint_as_float(ra[i]), a == C::NACC ? CROSS_SCALE : 1.0f, r[i]);
            }
          }
          if (rc == r_last) {             // all of this warp's share of the accumulator set is in registers: hand it back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(ab));
            if (warp == FIRST_EPI && lane == 0) TR(6, tile);
          }
#ifdef AGH_TC_SKIP_EPILOGUE
          if (r[0] != 12345.678f) continue;
#endif
          if (EPI == EPI_RES_BN) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = fmaf(rs[i] + (r[i] + bias), sc, sh);
            if (rc < r_last) load_res_sw(rc + 32);              // next chunk's residual flies while this chunk is stored
          }
          const int rbase = mt * BM + rc;
          float* dst = Cb + (size_t)rbase * ldc + col;
          if (EPI == EPI_TRAIN) {           // training primitive on the tensor cores: v = acc [+ bias] [relu] (agh_gemm keeps the
#pragma unroll                              // accumulate / mask forms on the CUDA-core kernels: their extra loads do not fit the
            for (int i = 0; i < 32; ++i) {  // register budget of this epilogue)
              float v = r[i] + bias;
              if (g.relu) v = fmaxf(v, 0.f);
              if (rbase + i < g.M) dst[(size_t)i * ldc] = v;
            }
          } else if (rbase + 32 <= g.M) {          // every row tile but the last: no per-row predicates
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float v = r[i];
              if (EPI == EPI_BIAS_RELU) v = fmaxf(v + bias, 0.f);
              dst[(size_t)i * ldc] = v;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float v = r[i];
              if (EPI == EPI_BIAS_RELU) v = fmaxf(v + bias, 0.f);
              if (rbase + i < g.M) dst[(size_t)i * ldc] = v;
            }
          }
        }
        if (warp == FIRST_EPI && lane == 0) TR(7, tile);
        continue;
      }
      const int row0 = mt * BM + q * 32;
      const int c_first = chalf * CH_PER_WARP * 32, c_last = c_first + (CH_PER_WARP - 1) * 32;
      // transposed domain: lane = column, rr[i] = residual of row row0 + i.  The residual does not depend on the
      // MMAs, so its (coalesced) loads are issued BEFORE waiting for the accumulator and one chunk ahead after that.
      float rr[32];
      auto load_res = [&](int c0, float (&dst)[32]) {
        const float* src = g.res + (size_t)row0 * ldr + n0 + c0 + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) dst[i] = (row0 + i < g.M) ? __ldg(src + (size_t)i * ldr) : 0.f;
      };
      // fragment domain: thread (g8, t4) owns rows row0 + 8 rq + g8 (rq = 0..3) and column pairs 8 k + 2 t4 (k = 0..3)
      const int g8 = lane >> 2, t4 = lane & 3;
      float2 fres[16];                                            // residual, index rq * 4 + k
      auto load_res_frag = [&](int c0) {
#pragma unroll
        for (int rq = 0; rq < 4; ++rq) {
          const int row = row0 + rq * 8 + g8;
          const float* src = g.res + (size_t)row * ldr + n0 + c0 + 2 * t4;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            fres[rq * 4 + k] = row < g.M ? __ldg(reinterpret_cast<const float2*>(src + 8 * k)) : make_float2(0.f, 0.f);
        }
      };
      if (EPI == EPI_RES_BN) {
        if (epi_frag(EPI)) load_res_frag(c_first);
        else if (epi_transposed(EPI)) load_res(c_first, rr);
      }
      mbar_wait(bar_tfull(ab), (tile / C::NBUF) & 1);
      if (warp == FIRST_EPI && lane == 0) TR(5, tile);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_acc = tmem_base + ab * ACC_COLS + ((uint32_t)(q * 32) << 16);
      if (epi_frag(EPI)) {
#pragma unroll 1
        for (int c0 = c_first; c0 <= c_last; c0 += 32) {
          float r[32];                                            // index (rq / 2) * 16 + k * 4 + (rq % 2) * 2 + {0, 1}
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {                        // TMEM lanes [0,16) and [16,32) of the quarter
            uint32_t ra[16];
            const uint32_t ta = t_acc + ((uint32_t)(hh * 16) << 16) + (uint32_t)c0;
            tmem_ld16x256b_x4(ta, ra);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) r[hh * 16 + i] = __uint_as_float(ra[i]);
#pragma unroll
            for (int a = 1; a <= C::NACC; ++a) {                   // further hi.hi partials, then the cross terms
              tmem_ld16x256b_x4(ta + (uint32_t)(a * BN), ra);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int i = 0; i < 16; ++i) r[hh * 16 + i] = fmaf(__uint_as_float(ra[i]), a == C::NACC ? CROSS_SCALE : 1.0f, r[hh * 16 + i]);
            }
          }
          if (c0 == c_last) {             // all of this warp's columns of the accumulator set are in registers: hand it back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(ab));
            if (warp == FIRST_EPI && lane == 0) TR(6, tile);
          }
#ifdef AGH_TC_SKIP_EPILOGUE
          if (r[0] != 12345.678f) continue;
#endif
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cl = c0 + 8 * k + 2 * t4;                   // column within the CTA's 128-column slab
            float2 bias = make_float2(0.f, 0.f), sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
            if (EPI == EPI_BIAS_RELU || (EPI == EPI_RES_BN && g.bias != nullptr)) bias = *reinterpret_cast<const float2*>(s_vec + cl);
            if (EPI == EPI_RES_BN) {
              sc = *reinterpret_cast<const float2*>(s_vec + BN + cl);
              sh = *reinterpret_cast<const float2*>(s_vec + 2 * BN + cl);
            }
#pragma unroll
            for (int rq = 0; rq < 4; ++rq) {
              const int row = row0 + rq * 8 + g8;
              const int ri = (rq >> 1) * 16 + k * 4 + (rq & 1) * 2;
              float2 v = make_float2(r[ri], r[ri + 1]);
              if (EPI == EPI_BIAS_RELU) {
                v.x = fmaxf(v.x + bias.x, 0.f); v.y = fmaxf(v.y + bias.y, 0.f);
              } else if (EPI == EPI_RES_BN) {
                const float2 rs = fres[rq * 4 + k];
                v.x = fmaf(rs.x + (v.x + bias.x), sc.x, sh.x); v.y = fmaf(rs.y + (v.y + bias.y), sc.y, sh.y);
              }
              if (row < g.M) *reinterpret_cast<float2*>(Cb + (size_t)row * ldc + n0 + cl) = v;
            }
          }
          if (EPI == EPI_RES_BN && c0 < c_last) load_res_frag(c0 + 32);      // next chunk's residual: in flight during its TMEM reads
        }
        if (warp == FIRST_EPI && lane == 0) TR(7, tile);
        continue;
      }
#pragma unroll 1
      for (int c0 = c_first; c0 <= c_last; c0 += 32) {
        float r[32];
        {   // accumulators one after the other (hi.hi partials, then the cross terms), added in fp32 round-to-nearest
          uint32_t ra[32];
          tmem_ld32(t_acc + (uint32_t)c0, ra);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(ra[i]);
#pragma unroll
          for (int a = 1; a <= C::NACC; ++a) {
            tmem_ld32(t_acc + (uint32_t)(a * BN + c0), ra);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = fmaf(__uint_as_float(ra[i]), a == C::NACC ? CROSS_SCALE : 1.0f, r[i]);
          }
        }
        if (c0 == c_last) {               // all of this warp's columns of the accumulator set are in registers: hand it back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty(ab));
          if (warp == FIRST_EPI && lane == 0) TR(6, tile);
        }
#ifdef AGH_TC_SKIP_EPILOGUE      // measurement only: main-loop-bound time (results are NOT written)
        if (r[0] != 12345.678f) continue;
#endif
        if (!epi_transposed(EPI)) {
          // row-wise epilogue: this thread owns row `row0 + lane`, 8 x 16-byte accesses per 32-column chunk
          const int row = row0 + lane;
          if (row >= g.M) continue;
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4) {
            const int col = n0 + c0 + v4 * 4;
            float4 v = make_float4(r[4 * v4], r[4 * v4 + 1], r[4 * v4 + 2], r[4 * v4 + 3]);
            if (EPI == EPI_BIAS_RELU) {
              const float4 bb = *reinterpret_cast<const float4*>(s_vec + c0 + v4 * 4);
              v.x = fmaxf(v.x + bb.x, 0.f); v.y = fmaxf(v.y + bb.y, 0.f); v.z = fmaxf(v.z + bb.z, 0.f); v.w = fmaxf(v.w + bb.w, 0.f);
            } else if (EPI == EPI_RES_BN) {
              if (g.bias != nullptr) {
                const float4 bb = *reinterpret_cast<const float4*>(s_vec + c0 + v4 * 4);
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
              }
              const float4 rr = __ldg(reinterpret_cast<const float4*>(g.res + (size_t)row * ldr + col));
              const float4 sc = *reinterpret_cast<const float4*>(s_vec + BN + c0 + v4 * 4);
              const float4 sh = *reinterpret_cast<const float4*>(s_vec + 2 * BN + c0 + v4 * 4);
              v.x = fmaf(rr.x + v.x, sc.x, sh.x); v.y = fmaf(rr.y + v.y, sc.y, sh.y);
              v.z = fmaf(rr.z + v.z, sc.z, sh.z); v.w = fmaf(rr.w + v.w, sc.w, sh.w);
            }
            *reinterpret_cast<float4*>(Cb + (size_t)row * ldc + col) = v;
          }
          continue;
        }
        // EPI_PLAIN: 32x32 in-register transpose (5 butterfly stages): before, lane = row and r[c] = column c; after,
        // lane = column and r[i] = row i, so that every store below is one coalesced 128-byte line per row
#pragma unroll
        for (int k = 16; k >= 1; k >>= 1) {
          const bool up = (lane & k) != 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if ((i & k) == 0) {
              const float send = up ? r[i] : r[i + k];
              const float recv = __shfl_xor_sync(0xffffffffu, send, k);
              if (up) r[i] = recv; else r[i + k] = recv;
            }
          }
        }
        const int col = n0 + c0 + lane;
        float bias = 0.f, sc = 1.f, sh = 0.f;
        if (EPI == EPI_BIAS_RELU || (EPI == EPI_RES_BN && g.bias != nullptr)) bias = s_vec[c0 + lane];
        if (EPI == EPI_RES_BN) { sc = s_vec[BN + c0 + lane]; sh = s_vec[2 * BN + c0 + lane]; }
        if (EPI == EPI_RES_BN) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = fmaf(rr[i] + (r[i] + bias), sc, sh);
          if (c0 < c_last) load_res(c0 + 32, rr);         // next chunk's residual flies while this chunk is stored
        }
#pragma unroll      // full unroll: r[] must stay in registers (static indices only)
        for (int i = 0; i < 32; ++i) {
          const int row = row0 + i;
          float v = r[i];
          if (EPI == EPI_BIAS_RELU) v = fmaxf(v + bias, 0.f);
          if (row < g.M) Cb[(size_t)row * ldc + col] = v;
        }
      }
      if (warp == FIRST_EPI && lane == 0) TR(7, tile);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
#ifdef AGH_TC_TRACE
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && g.M > 500000) {
    const int l = atomicAdd(&g_trace_launches, 1);
    if (l < 10) {
      const int tiles = (num_mt + (int)gridDim.x - 1) / (int)gridDim.x;
      const unsigned long long base = g_trace[0][0];
      for (int role = 0; role < 8; ++role) {
        const int cnt = role < 5 ? tiles * nkb : tiles;
        for (int i = 0; i < cnt && i < TR_CAP; ++i)
          printf("TRACE %d %d %d %d %d %lld\n", l, EPI, (int)WRES, role, i, (long long)(g_trace[role][i] - base));
      }
    }
  }
#endif
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------- host side

template <int EPI, bool WRES>
static int launch_cfg(const GemmArgs& g, cudaStream_t st) {
  using C = Cfg<WRES>;
  CUtensorMap tmA, tmW;
  int rc = g.a_slab ? make_map(&tmA, g.A, (uint64_t)g.M * (uint64_t)(g.K / BN), (uint64_t)BN, (uint64_t)BN)
                    : make_map(&tmA, g.A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda);
  if (rc) return rc;
  rc = make_map(&tmW, g.Wt, (uint64_t)g.Nc, (uint64_t)g.K, (uint64_t)g.K);
  if (rc) return rc;
  const int n_tiles = g.Nc / BN, m_tiles = ceil_div(g.M, BM);
  int per_slab = sm_count() / n_tiles;                 // persistent: one CTA per SM, CTAs of a slab stride over the row tiles
  if (per_slab < 1) per_slab = 1;
  if (per_slab > m_tiles) per_slab = m_tiles;
  dim3 grid(per_slab, n_tiles);
  if (EPI == EPI_RES_BN && (g.ldc != BN || g.ldr != BN)) {
    set_error("gemm_tc: the residual epilogue expects [M][128] output and residual (ldc=%d ldr=%d)", g.ldc, g.ldr);
    return AGH_E_UNSUPPORTED;
  }
  constexpr bool SLAB_OK = WRES && (EPI == EPI_PLAIN || EPI == EPI_BIAS_RELU);
  if (g.c_slab && !SLAB_OK) {
    set_error("gemm_tc: slab-major output is built for the K = 128 plain / bias+ReLU kernels only");
    return AGH_E_UNSUPPORTED;
  }
  if constexpr (SLAB_OK) {
    if (g.c_slab) {
      auto kern = gemm_tc_kernel<EPI, WRES, true>;
      AGH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
      kern<<<grid, NTHREADS, C::SMEM, st>>>(tmA, tmW, g);
      AGH_LAUNCH_CHECK();
      return AGH_OK;
    }
  }
  auto kern = gemm_tc_kernel<EPI, WRES, false>;
  AGH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  kern<<<grid, NTHREADS, C::SMEM, st>>>(tmA, tmW, g);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

}  // namespace tc

template <int EPI>
int launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
  if (g.K % tc::KST != 0 || (g.lda % 4) != 0 || g.Nc % tc::BN != 0 || g.M <= 0) {
    set_error("gemm_tc: unsupported M=%d K=%d Nc=%d lda=%d", g.M, g.K, g.Nc, g.lda);
    return AGH_E_UNSUPPORTED;
  }
  if (g.K == 128) return tc::launch_cfg<EPI, true>(g, st);
  return tc::launch_cfg<EPI, false>(g, st);
}

template int launch_gemm_tc<EPI_PLAIN>(const GemmArgs&, cudaStream_t);
template int launch_gemm_tc<EPI_BIAS_RELU>(const GemmArgs&, cudaStream_t);
template int launch_gemm_tc<EPI_RES_BN>(const GemmArgs&, cudaStream_t);
template int launch_gemm_tc<EPI_DEC>(const GemmArgs&, cudaStream_t);
template int launch_gemm_tc<EPI_TRAIN>(const GemmArgs&, cudaStream_t);

}  // namespace agh
