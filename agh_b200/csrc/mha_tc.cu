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
//   warp 0      : TMA producer  - raw fp32 Q / K / V tiles ([112 rows][128] as four boxes of 32 columns) into a rotation of
//                                 four 56 KB buffers: Q of the next instance lands a whole instance ahead, K and V as soon as
//                                 the current instance's last score MMA has retired
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
constexpr int BUF = 4 * BOX;                 // one matrix of one instance
constexpr int NBUF = 4;
constexpr int PAD = 2048;                    // the M = 128 A operand reads 16 rows past a 112-row block
constexpr int BAR_OFF = NBUF * BUF + PAD;
constexpr int B_RAW = 0, B_READY = 4, B_EMPTY = 8, B_SFULL = 12, B_SFREE = 14, B_PFULL = 16, B_OFULL = 18, B_OEMPTY = 20, NBARS = 22;
constexpr int SMEM = BAR_OFF + 256;
static_assert(8 * NBARS + 16 <= 256, "barrier block");
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
    // ================= TMA producer: loads j = 3 i + m (m = 0 Q, 1 K, 2 V of the CTA's i-th instance) into buffer j % 4 =================
    if (elect_one()) {
      uint32_t j = 0;
      for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
        for (int m = 0; m < 3; ++m, ++j) {
          const int u = j & 3;
          mbar_wait(bar(B_EMPTY + u), ((j >> 2) & 1) ^ 1);
          mbar_expect_tx(bar(B_RAW + u), BUF);
          const uint32_t dst = smem_u32(smem + u * BUF);
          const int row = m * g.M + b * n;
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_2d(dst + c * BOX, &tmQKV, bar(B_RAW + u), 32 * c, row);
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
    auto buf_of = [&](uint32_t j) { return smem_u32(smem + (j & 3) * BUF); };
    auto issue_pv = [&](uint32_t G) {
      const uint32_t i = G >> 3, h = G & 7, gi = G & 1, c = G >> 1;   // c: count of heads this buffer pair has served
      const uint32_t jv = 3 * i + 2;
      if (h == 0) mbar_wait(bar(B_READY + (jv & 3)), (jv >> 2) & 1);
      mbar_wait(bar(B_PFULL + gi), c & 1);
      mbar_wait(bar(B_OEMPTY + gi), (c & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        const uint32_t vb = buf_of(jv) + (h >> 1) * BOX + (h & 1) * 64;      // [hi 16 dims | lo' 16 dims] of head h: 64 bytes of the key row
        const uint32_t t_s = tmem_base + T_S + 224 * gi, t_o = tmem_base + T_O + 32 * gi;
#pragma unroll
        for (int k = 0; k < ROWS / 16; ++k) {
          const uint64_t dv = umma_desc(vb + k * 2048);                      // 16 keys = two 8-key atoms
          att_umma_ts(t_o, t_s + 8 * k, dv, id_o2, k != 0);                  // p_hi x [v_hi | v_lo'] -> [O main | O cross]
          att_umma_ts(t_o + 16, t_s + ROWS + 8 * k, dv, id_o1, 1);           // p_lo' x v_hi -> O cross
        }
        umma_commit(bar(B_OFULL + gi));
        umma_commit(bar(B_SFREE + gi));
        if (h == 7) umma_commit(bar(B_EMPTY + (jv & 3)));
      }
      __syncwarp();
    };
    for (uint32_t G = 0; G < total; ++G) {
      const uint32_t i = G >> 3, h = G & 7, gi = G & 1, c = G >> 1;
      const uint32_t jq = 3 * i, jk = 3 * i + 1;
      if (h == 0) {
        mbar_wait(bar(B_READY + (jq & 3)), (jq >> 2) & 1);
        mbar_wait(bar(B_READY + (jk & 3)), (jk >> 2) & 1);
      }
      mbar_wait(bar(B_SFREE + gi), (c & 1) ^ 1);                      // the values MMA that read P from this buffer has retired
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        const uint32_t blk = (h >> 2) * 2 * BOX, adv = (h & 3) * 32;  // dims 64 (h / 4) .. : blocks [hi | lo']; 16 dims = 32 bytes
        const uint64_t da_hi = umma_desc(buf_of(jq) + blk + adv), da_lo = umma_desc(buf_of(jq) + blk + BOX + adv);
        const uint64_t dk = umma_desc(buf_of(jk) + blk + adv);        // 112 rows of k_hi, then 112 rows of k_lo'
        const uint32_t t_s = tmem_base + T_S + 224 * gi;
        umma_tf32(t_s, da_hi, dk, id_s2, 0);                          // q_hi x [k_hi ; k_lo'] -> [S main | S cross]
        umma_tf32(t_s + ROWS, da_lo, dk, id_s1, 1);                   // q_lo' x k_hi -> S cross
        umma_commit(bar(B_SFULL + gi));
        if (h == 7) {                                                 // Q and K of this instance are no longer read
          umma_commit(bar(B_EMPTY + (jq & 3)));
          umma_commit(bar(B_EMPTY + (jk & 3)));
        }
      }
      __syncwarp();
      if (G >= 1) issue_pv(G - 1);
    }
    if (total > 0) issue_pv(total - 1);
  } else if (warp < 6) {
    // ================= splitters (warps 2..5): thread = row =================
    const int t = threadIdx.x - 64;
    const int x7 = t & 7;
    uint32_t j = 0;
    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
      for (int m = 0; m < 3; ++m, ++j) {
        const int u = j & 3;
        mbar_wait(bar(B_RAW + u), (j >> 2) & 1);
        if (t < ROWS) {
          unsigned char* base = smem + u * BUF + t * 128;
          if (m < 2) {
            // Q (scaled) / K: boxes (2p, 2p+1) = dims 64 p .. +63 -> K-major blocks hi (in box 2p) and lo' (in box 2p+1)
            const float sc = m == 0 ? QSCALE : 1.0f;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              unsigned char* p0 = base + (2 * p) * BOX;
              unsigned char* p1 = p0 + BOX;
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
            }
          } else {
            // V: box c = heads 2c, 2c+1 (16 dims each) -> in place [head 2c: hi 16 | lo' 16][head 2c+1: hi 16 | lo' 16] per key row
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              unsigned char* p0 = base + c * BOX;
              float4 v[8];
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) v[q4] = *reinterpret_cast<const float4*>(p0 + ((q4 ^ x7) << 4));
#pragma unroll
              for (int e = 0; e < 2; ++e) {                 // head 2c + e: fp32 chunks 4e .. 4e+3
                uint32_t hw[8], lw[8];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const float4 a = v[4 * e + q4];
                  att_split_pair(a.x, a.y, hw[2 * q4], lw[2 * q4]);
                  att_split_pair(a.z, a.w, hw[2 * q4 + 1], lw[2 * q4 + 1]);
                }
                *reinterpret_cast<uint4*>(p0 + (((4 * e + 0) ^ x7) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(p0 + (((4 * e + 1) ^ x7) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                *reinterpret_cast<uint4*>(p0 + (((4 * e + 2) ^ x7) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                *reinterpret_cast<uint4*>(p0 + (((4 * e + 3) ^ x7) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
              }
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_READY + u));
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
        if (cnt > 0) drain_o();
        mbar_wait(bar(B_SFULL + gi), cnt & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // pass 1: row maximum over the valid keys (the main term decides it; the cross term is 2^-11 of it)
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int c0 = 0; c0 < ROWS; c0 += 32) {
          if (c0 + 32 <= ROWS) {
            uint32_t a[32];
            tmem_ld32(t_s + c0, a);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c0 + 32 <= n) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(a[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c0 + i < n) ? __uint_as_float(a[i]) : -CUDART_INF_F);
            }
          } else {
            uint32_t a[16];
            att_tmem_ld16(t_s + c0, a);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (c0 + i < n) ? __uint_as_float(a[i]) : -CUDART_INF_F);
          }
        }
        // pass 2: p = exp2(s - max), row sum, (hi, lo') split written over the scores
        float l = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < ROWS; c0 += 32) {
          if (c0 + 32 <= ROWS) {
            uint32_t a[32], x[32];
            tmem_ld32(t_s + c0, a);
            tmem_ld32(t_s + ROWS + c0, x);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t hw[16], lw[16];
            const bool full = c0 + 32 <= n;                  // warp-uniform: every key of the chunk is a real node
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float s0 = fmaf(__uint_as_float(x[2 * i]), CS, __uint_as_float(a[2 * i]));
              const float s1 = fmaf(__uint_as_float(x[2 * i + 1]), CS, __uint_as_float(a[2 * i + 1]));
              float p0 = att_ex2(s0 - mx), p1 = att_ex2(s1 - mx);
              if (!full) {
                p0 = (c0 + 2 * i < n) ? p0 : 0.f;
                p1 = (c0 + 2 * i + 1 < n) ? p1 : 0.f;
              }
              l += p0 + p1;
              att_split_pair(p0, p1, hw[i], lw[i]);
            }
            att_tmem_st16(t_s + c0 / 2, hw);
            att_tmem_st16(t_s + ROWS + c0 / 2, lw);
          } else {
            uint32_t a[16], x[16];
            att_tmem_ld16(t_s + c0, a);
            att_tmem_ld16(t_s + ROWS + c0, x);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float s0 = fmaf(__uint_as_float(x[2 * i]), CS, __uint_as_float(a[2 * i]));
              const float s1 = fmaf(__uint_as_float(x[2 * i + 1]), CS, __uint_as_float(a[2 * i + 1]));
              const float p0 = (c0 + 2 * i < n) ? att_ex2(s0 - mx) : 0.f;
              const float p1 = (c0 + 2 * i + 1 < n) ? att_ex2(s1 - mx) : 0.f;
              l += p0 + p1;
              att_split_pair(p0, p1, hw[i], lw[i]);
            }
            att_tmem_st8(t_s + c0 / 2, hw);
            att_tmem_st8(t_s + ROWS + c0 / 2, lw);
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PFULL + gi));
        inv_prev = 1.0f / l;
        out_prev = g.heads + ((size_t)b * n + row) * D + h * HD;
      }
    }
    if (cnt > 0) drain_o();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
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
