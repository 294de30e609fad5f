// Encoder self-attention on tcgen05 (graph_encoder.py:83-104) for graphs of up to 112 nodes: one persistent CTA per SM walks
// instances; per (instance, head) the scores S = Q_h K_h^T and the values O = softmax(S) V_h are two tensor-core
// contractions with the accumulators in TMEM, and the softmax runs in registers between them, one thread per query row.
//
// Same two-term FP16 operand split as the GEMMs (x = hi + lo'/2048; hi.hi into a main accumulator, the two cross products
// into a second one that the reader scales by 2^-11):
//   S_h : q_hi x [k_hi ; k_lo']  -> [S main | S cross]   (one MMA of N = 224: the lo' rows follow the hi rows in shared memory)
//         q_lo' x k_hi            -> S cross              (N = 112)            K = 16 = the head dimension: ONE k-step
//   P   : exp2(S - max) per row, split into (hi, lo') FP16 pairs and written back to TMEM OVER the scores it came from
//         (P hi over S main, P lo' over S cross: a thread's writes always trail its reads)
//   O_h : p_hi x [v_hi | v_lo'] -> [O main | O cross]    (N = 32; V is the MN-major B operand, keys are the K dimension)
//         p_lo' x v_hi           -> O cross               (N = 16)             7 k-steps of 16 keys
// The head slice of Q / K (16 of the 128 dims) is a 32-byte advance of the descriptor inside the 128-byte swizzle row,
// exactly a k-step of a GEMM; V is re-packed per box as [head a: hi | lo'][head b: hi | lo'] so that (hi | lo') of a head is
// one contiguous 64-byte N-slice.
//
//   warp 0      : TMA producer  - raw fp32 Q / K / V tiles as UNITS of two [112 rows][32 columns] boxes (64 dims = four heads of
//                                 one matrix) into a rotation of eight 28 KB slots; every unit lands and is split at least
//                                 half an instance before the first MMA that reads it
//   warp 1      : MMA issuer    - S_G, PV_{G-1}, S_{G+1}, PV_G, ... over the running head index G (two score buffers)
//   warps 2..5  : splitters     - one thread per row rewrites the raw boxes in place as FP16 (hi, lo') operand blocks
//   warps 6..13 : softmax       - two groups of four warps (even / odd heads), thread = query row = TMEM lane; before the
//                                 softmax of a head the group drains O of its previous head (x 1 / row sum) to HBM
// TMEM (512 columns): S0 224 | S1 224 | O0 32 | O1 32.
#include "common.cuh"
#include "gemm.cuh"
#include "tc.cuh"
#include <math_constants.h>

namespace agh {

namespace tc {

#if AGH_TC_F16

namespace att {
constexpr int ROWS = 112;                    // rows per tile (queries / keys), 14 swizzle atoms of 8
constexpr int BOX = ROWS * 128;              // one [112 x 32] fp32 box = one [112 x 64] fp16 block = 14 KB
constexpr int BUF = 2 * BOX;                 // one UNIT: two boxes = half of one matrix of one instance (64 dims = 4 heads)
constexpr int NBUF = 8;
constexpr int PAD = 2048;                    // the M = 128 A operand reads 16 rows past a 112-row block
constexpr int BAR_OFF = NBUF * BUF + PAD;
constexpr int B_RAW = 0, B_READY = 8, B_EMPTY = 16, B_SFULL = 24, B_SFREE = 26, B_PFULL = 28, B_OFULL = 30, B_OEMPTY = 32, NBARS = 34;
constexpr int SMEM = BAR_OFF + 384;
static_assert(8 * NBARS + 16 <= 384, "barrier block");
static_assert(SMEM <= 232448, "shared memory budget");
constexpr int T_S = 0, T_O = 448;            // S buffer g at T_S + 224 g (main 112 | cross 112), O buffer g at T_O + 32 g (main 16 | cross 16)
constexpr int NT = 32 * 14;
constexpr float QSCALE = 0.25f * 1.4426950408889634f;      // 1/sqrt(16) (graph_encoder.py:40,89) and log2(e): base-2 softmax
constexpr float CS = 1.0f / 2048.0f;
}  // namespace att

__device__ __forceinline__ void att_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void att_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]),
        "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]),
        "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
}
__device__ __forceinline__ void att_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void att_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void att_split_pair(float f0, float f1, uint32_t& h2, uint32_t& l2) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(f1), "f"(f0));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
  const float d0 = (f0 - hf.x) * 2048.0f, d1 = (f1 - hf.y) * 2048.0f;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(d1), "f"(d0));
}
__device__ __forceinline__ float att_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32 pairs (fma.rn.f32x2 / add / mul -> FFMA2-class SASS): two IEEE operations per issued instruction
__device__ __forceinline__ uint64_t att_pack2(float x, float y) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void att_unpack2(uint64_t v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ uint64_t att_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t att_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t att_mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float att_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// -DAGH_MHA_TRACE: CTA 0 stamps %globaltimer at the pipeline hand-offs of its first instances and prints them (tools/mha_timeline.py)
#ifdef AGH_MHA_TRACE
constexpr int ATR_CAP = 64;
__device__ unsigned long long g_atrace[16][ATR_CAP];
__device__ int g_atrace_launches = 0;
__device__ __forceinline__ unsigned long long atime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ATR(role, idx) do { if (blockIdx.x == 0 && (int)(idx) < ATR_CAP) g_atrace[role][idx] = atime(); } while (0)
#else
#define ATR(role, idx) do { } while (0)
#endif

struct AttArgs {
  float* heads;            // [B*n][128]
  int B, n;
  int M;                   // B * n: rows of one of the three stacked matrices
};

__global__ void __launch_bounds__(att::NT, 1) mha_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttArgs g) {
  using namespace att;
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();      // the swizzled blocks need 1024-byte alignment and there is no room for slack
  unsigned char* bars = smem + BAR_OFF;
  const uint32_t bar_base = smem_u32(bars);
  auto bar = [&](int i) { return bar_base + 8 * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 * NBARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = g.n;

  if (threadIdx.x == 0) {
    for (int u = 0; u < NBUF; ++u) {
      mbar_init(bar(B_RAW + u), 1);
      mbar_init(bar(B_READY + u), 4);
      mbar_init(bar(B_EMPTY + u), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(B_SFULL + s), 1);
      mbar_init(bar(B_SFREE + s), 1);
      mbar_init(bar(B_PFULL + s), 4);
      mbar_init(bar(B_OFULL + s), 1);
      mbar_init(bar(B_OEMPTY + s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the pad behind the last buffer is read (never used) by the tensor core: keep it finite
  for (int i = threadIdx.x; i < PAD / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + NBUF * BUF)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    // Units of two boxes (64 dims = the four heads 4 p .. 4 p + 3 of one matrix), loaded in the order the MMAs first need them:
    // k = 0 Q p0, 1 K p0, 2 V p0, 3 Q p1, 4 K p1, 5 V p1 of the CTA's i-th instance -> load j = 6 i + k into slot j % 8.  A slot
    // is free again when the last MMA that reads load j - 8 has retired (Q / K p0 after head 3's scores, V p0 after head 3's
    // values, ...), so every unit is fetched and split at least half an instance before its first use.
    if (elect_one()) {
      uint32_t j = 0;
      for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
        for (int k = 0; k < 6; ++k, ++j) {
          const int u = j & 7, m = k % 3, pr = k / 3;
          mbar_wait(bar(B_EMPTY + u), ((j >> 3) & 1) ^ 1);
          ATR(0, j);
          mbar_expect_tx(bar(B_RAW + u), BUF);
          const uint32_t dst = smem_u32(smem + u * BUF);
          const int row = m * g.M + b * n;
          tma_load_2d(dst, &tmQKV, bar(B_RAW + u), 64 * pr, row);
          tma_load_2d(dst + BOX, &tmQKV, bar(B_RAW + u), 64 * pr + 32, row);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t MN_B = 1u << 16;                               // idesc: B operand MN-major
    constexpr uint32_t id_s2 = umma_idesc(128, 2 * ROWS), id_s1 = umma_idesc(128, ROWS);
    constexpr uint32_t id_o2 = umma_idesc(128, 32) | MN_B, id_o1 = umma_idesc(128, 16) | MN_B;
    const bool leader = elect_one();
    int nloc = 0;
    for (int b = blockIdx.x; b < g.B; b += gridDim.x) ++nloc;
    const uint32_t total = 8u * (uint32_t)nloc;                       // running head index G = 8 i + h
    auto buf_of = [&](uint32_t j) { return smem_u32(smem + (j & 7) * BUF); };
    auto issue_pv = [&](uint32_t G) {
      const uint32_t i = G >> 3, h = G & 7, gi = G & 1, c = G >> 1;   // c: count of heads this buffer pair has served
      const uint32_t jv = 6 * i + 3 * (h >> 2) + 2;                   // the V unit of heads 4 (h / 4) .. + 3
      if (lane == 0) ATR(6, G);
      if ((h & 3) == 0) mbar_wait(bar(B_READY + (jv & 7)), (jv >> 3) & 1);
      mbar_wait(bar(B_PFULL + gi), c & 1);
      mbar_wait(bar(B_OEMPTY + gi), (c & 1) ^ 1);
      if (lane == 0) ATR(7, G);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        const uint32_t vb = buf_of(jv) + ((h >> 1) & 1) * BOX + (h & 1) * 64;   // [hi 16 dims | lo' 16 dims] of head h: 64 bytes of the key row
        const uint32_t t_s = tmem_base + T_S + 224 * gi, t_o = tmem_base + T_O + 32 * gi;
#pragma unroll
        for (int k = 0; k < ROWS / 16; ++k) {
          const uint64_t dv = umma_desc(vb + k * 2048);                      // 16 keys = two 8-key atoms
          att_umma_ts(t_o, t_s + 8 * k, dv, id_o2, k != 0);                  // p_hi x [v_hi | v_lo'] -> [O main | O cross]
          att_umma_ts(t_o + 16, t_s + ROWS + 8 * k, dv, id_o1, 1);           // p_lo' x v_hi -> O cross
        }
        umma_commit(bar(B_OFULL + gi));
        umma_commit(bar(B_SFREE + gi));
        if ((h & 3) == 3) umma_commit(bar(B_EMPTY + (jv & 7)));
      }
      __syncwarp();
      if (lane == 0) ATR(8, G);
    };
    for (uint32_t G = 0; G < total; ++G) {
      const uint32_t i = G >> 3, h = G & 7, gi = G & 1, c = G >> 1;
      const uint32_t jq = 6 * i + 3 * (h >> 2), jk = jq + 1;          // the Q and K units of heads 4 (h / 4) .. + 3
      if (lane == 0) ATR(3, G);
      if ((h & 3) == 0) {
        mbar_wait(bar(B_READY + (jq & 7)), (jq >> 3) & 1);
        mbar_wait(bar(B_READY + (jk & 7)), (jk >> 3) & 1);
      }
      mbar_wait(bar(B_SFREE + gi), (c & 1) ^ 1);                      // the values MMA that read P from this buffer has retired
      if (lane == 0) ATR(4, G);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        const uint32_t adv = (h & 3) * 32;                            // unit = blocks [hi | lo'] of 64 dims; 16 dims = 32 bytes
        const uint64_t da_hi = umma_desc(buf_of(jq) + adv), da_lo = umma_desc(buf_of(jq) + BOX + adv);
        const uint64_t dk = umma_desc(buf_of(jk) + adv);              // 112 rows of k_hi, then 112 rows of k_lo'
        const uint32_t t_s = tmem_base + T_S + 224 * gi;
        umma_tf32(t_s, da_hi, dk, id_s2, 0);                          // q_hi x [k_hi ; k_lo'] -> [S main | S cross]
        umma_tf32(t_s + ROWS, da_lo, dk, id_s1, 1);                   // q_lo' x k_hi -> S cross
        umma_commit(bar(B_SFULL + gi));
        if ((h & 3) == 3) {                                           // these Q and K units are no longer read
          umma_commit(bar(B_EMPTY + (jq & 7)));
          umma_commit(bar(B_EMPTY + (jk & 7)));
        }
      }
      __syncwarp();
      if (lane == 0) ATR(5, G);
      if (G >= 1) issue_pv(G - 1);
    }
    if (total > 0) issue_pv(total - 1);
  } else if (warp < 6) {
    // ================= splitters (warps 2..5): thread = row =================
    const int t = threadIdx.x - 64;
    const int x7 = t & 7;
    uint32_t j = 0;
    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
      for (int k = 0; k < 6; ++k, ++j) {
        const int u = j & 7, m = k % 3;
        mbar_wait(bar(B_RAW + u), (j >> 3) & 1);
        if (t == 0) ATR(1, j);
        if (t < ROWS) {
          unsigned char* p0 = smem + u * BUF + t * 128;
          unsigned char* p1 = p0 + BOX;
          if (m < 2) {
            // Q (scaled) / K: the two boxes = 64 dims -> K-major blocks hi (in the first box) and lo' (in the second)
            const float sc = m == 0 ? QSCALE : 1.0f;
            float4 v[16];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              v[c] = *reinterpret_cast<const float4*>(p0 + ((c ^ x7) << 4));
              v[8 + c] = *reinterpret_cast<const float4*>(p1 + ((c ^ x7) << 4));
            }
#pragma unroll
            for (int jc = 0; jc < 8; ++jc) {              // fp16 chunk jc = dims 8 jc .. 8 jc + 7 = fp32 chunks 2 jc, 2 jc + 1
              const float4 a = v[2 * jc], c4 = v[2 * jc + 1];
              uint32_t hw[4], lw[4];
              att_split_pair(a.x * sc, a.y * sc, hw[0], lw[0]);
              att_split_pair(a.z * sc, a.w * sc, hw[1], lw[1]);
              att_split_pair(c4.x * sc, c4.y * sc, hw[2], lw[2]);
              att_split_pair(c4.z * sc, c4.w * sc, hw[3], lw[3]);
              *reinterpret_cast<uint4*>(p0 + ((jc ^ x7) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(p1 + ((jc ^ x7) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          } else {
            // V: a box = two heads (16 dims each) -> in place [head a: hi 16 | lo' 16][head b: hi 16 | lo' 16] per key row
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              unsigned char* pb = p0 + c * BOX;
              float4 v[8];
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) v[q4] = *reinterpret_cast<const float4*>(pb + ((q4 ^ x7) << 4));
#pragma unroll
              for (int e = 0; e < 2; ++e) {                 // head e of the box: fp32 chunks 4e .. 4e+3
                uint32_t hw[8], lw[8];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const float4 a = v[4 * e + q4];
                  att_split_pair(a.x, a.y, hw[2 * q4], lw[2 * q4]);
                  att_split_pair(a.z, a.w, hw[2 * q4 + 1], lw[2 * q4 + 1]);
                }
                *reinterpret_cast<uint4*>(pb + (((4 * e + 0) ^ x7) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(pb + (((4 * e + 1) ^ x7) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                *reinterpret_cast<uint4*>(pb + (((4 * e + 2) ^ x7) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                *reinterpret_cast<uint4*>(pb + (((4 * e + 3) ^ x7) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
              }
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_READY + u));
        if (t == 0) ATR(2, j);
      }
    }
  } else {
    // ================= softmax + O epilogue (warps 6..13): group gi = (warp - 6) / 4 takes heads h = gi, gi + 2, .. =================
    const int gi = (warp - 6) >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_s = lane_base + T_S + 224 * gi, t_o = lane_base + T_O + 32 * gi;
    uint32_t cnt = 0;
    float inv_prev = 0.f;
    float* out_prev = nullptr;
    auto drain_o = [&]() {                                  // O of this group's previous head -> heads[row][16 h ..]
      mbar_wait(bar(B_OFULL + gi), (cnt - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t o[32];
      tmem_ld32(t_o, o);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_OEMPTY + gi));
      if (row < n) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          float4 r;
          r.x = fmaf(__uint_as_float(o[16 + 4 * v4]), CS, __uint_as_float(o[4 * v4])) * inv_prev;
          r.y = fmaf(__uint_as_float(o[17 + 4 * v4]), CS, __uint_as_float(o[4 * v4 + 1])) * inv_prev;
          r.z = fmaf(__uint_as_float(o[18 + 4 * v4]), CS, __uint_as_float(o[4 * v4 + 2])) * inv_prev;
          r.w = fmaf(__uint_as_float(o[19 + 4 * v4]), CS, __uint_as_float(o[4 * v4 + 3])) * inv_prev;
          *reinterpret_cast<float4*>(out_prev + 4 * v4) = r;
        }
      }
    };
    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
      for (int h = gi; h < 8; h += 2, ++cnt) {
        if (q == 0 && lane == 0) ATR(9 + 3 * gi, cnt);
        if (cnt > 0) drain_o();
        mbar_wait(bar(B_SFULL + gi), cnt & 1);
        if (q == 0 && lane == 0) ATR(10 + 3 * gi, cnt);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // tcgen05.wait::ld costs far more than the load it waits for, so the passes use few, wide loads:
        // pass 1: row maximum over the valid keys (the main term decides it; the cross term is 2^-11 of it): all 112 main
        // columns land behind ONE wait
        float mx = -CUDART_INF_F;
        {
          uint32_t a0[64], a1[32], a2[16];
          tmem_ld64(t_s, a0);
          tmem_ld32(t_s + 64, a1);
          att_tmem_ld16(t_s + 96, a2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          auto red = [&](const uint32_t* v, int c0, int cnt) {
            if (c0 + cnt <= n) {
#pragma unroll
              for (int i = 0; i < cnt; i += 2) mx = att_max3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            } else {
#pragma unroll
              for (int i = 0; i < cnt; ++i) mx = fmaxf(mx, (c0 + i < n) ? __uint_as_float(v[i]) : -CUDART_INF_F);
            }
          };
          red(a0, 0, 32); red(a0 + 32, 32, 32); red(a1, 64, 32); red(a2, 96, 16);
        }
        // pass 2: p = exp2(s - max), row sum, (hi, lo') split written over the scores; chunks of 32 keys (16 at the end)
        uint64_t l2 = att_pack2(0.f, 0.f);
        {
          const uint64_t nm2 = att_pack2(-mx, -mx), cs2 = att_pack2(CS, CS), k2 = att_pack2(2048.f, 2048.f), nk2 = att_pack2(-2048.f, -2048.f);
          auto pair = [&](uint32_t am0, uint32_t am1, uint32_t x0, uint32_t x1, int key, bool full, uint32_t& h2, uint32_t& lo2) {
            const uint64_t am = att_add2(att_pack2(__uint_as_float(am0), __uint_as_float(am1)), nm2);
            const uint64_t sv = att_fma2(att_pack2(__uint_as_float(x0), __uint_as_float(x1)), cs2, am);
            float s0, s1;
            att_unpack2(sv, s0, s1);
            float p0 = att_ex2(s0), p1 = att_ex2(s1);
            if (!full) {
              p0 = (key < n) ? p0 : 0.f;
              p1 = (key + 1 < n) ? p1 : 0.f;
            }
            const uint64_t pv = att_pack2(p0, p1);
            l2 = att_add2(l2, pv);
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(p1), "f"(p0));
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
            const uint64_t dv = att_fma2(pv, k2, att_mul2(att_pack2(hf.x, hf.y), nk2));      // (p - hi) * 2048
            float d0, d1;
            att_unpack2(dv, d0, d1);
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(d1), "f"(d0));
          };
#pragma unroll
          for (int c0 = 0; c0 < 96; c0 += 32) {
            uint32_t a[32], x[32];
            tmem_ld32(t_s + c0, a);
            tmem_ld32(t_s + ROWS + c0, x);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const bool full = c0 + 32 <= n;                  // warp-uniform: every key of the chunk is a real node
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pair(a[2 * i], a[2 * i + 1], x[2 * i], x[2 * i + 1], c0 + 2 * i, full, hw[i], lw[i]);
            att_tmem_st16(t_s + c0 / 2, hw);                 // P hi over S main, P lo' over S cross: behind every column still to be read
            att_tmem_st16(t_s + ROWS + c0 / 2, lw);
          }
          {
            uint32_t a[16], x[16];
            att_tmem_ld16(t_s + 96, a);
            att_tmem_ld16(t_s + ROWS + 96, x);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const bool full = ROWS <= n;
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pair(a[2 * i], a[2 * i + 1], x[2 * i], x[2 * i + 1], 96 + 2 * i, full, hw[i], lw[i]);
            att_tmem_st8(t_s + 48, hw);
            att_tmem_st8(t_s + ROWS + 48, lw);
          }
        }
        float l;
        {
          float la, lb;
          att_unpack2(l2, la, lb);
          l = la + lb;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PFULL + gi));
        if (q == 0 && lane == 0) ATR(11 + 3 * gi, cnt);
        inv_prev = 1.0f / l;
        out_prev = g.heads + ((size_t)b * n + row) * D + h * HD;
      }
    }
    if (cnt > 0) drain_o();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
#ifdef AGH_MHA_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0 && g.B > 5000) {
    const int l = atomicAdd(&g_atrace_launches, 1);
    if (l < 2) {
      const unsigned long long base = g_atrace[3][0];
      for (int role = 0; role < 15; ++role)
        for (int i = 0; i < ATR_CAP; ++i) printf("ATRACE %d %d %d %lld\n", l, role, i, (long long)(g_atrace[role][i] - base));
    }
  }
#endif
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

static int make_map_qkv(CUtensorMap* map, const float* base, uint64_t rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return AGH_E_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)D, rows};
  const cuuint64_t strides[1] = {D * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)att::ROWS};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (attention) failed (%d) rows=%llu", (int)rc, (unsigned long long)rows);
    return AGH_E_CUDA;
  }
  return AGH_OK;
}

#endif  // AGH_TC_F16

}  // namespace tc

bool mha_tc_supported(int n) { return AGH_TC_F16 != 0 && n >= 2 && n <= 112; }

// qkv: slab-major [3][B*n][128] (Q, K, V as three stacked row-major matrices)
int launch_mha_tc(const float* qkv, int B, int n, float* heads, cudaStream_t st) {
#if AGH_TC_F16
  using namespace tc;
  CUtensorMap tm;
  const int rc = make_map_qkv(&tm, qkv, (uint64_t)3 * (uint64_t)B * (uint64_t)n);
  if (rc) return rc;
  AttArgs a{heads, B, n, B * n};
  const int grid = B < sm_count() ? B : sm_count();
  AGH_CUDA(cudaFuncSetAttribute(mha_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, att::SMEM));
  mha_tc_kernel<<<grid, att::NT, att::SMEM, st>>>(tm, a);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
#else
  (void)qkv; (void)B; (void)n; (void)heads; (void)st;
  return AGH_E_UNSUPPORTED;
#endif
}

}  // namespace agh
