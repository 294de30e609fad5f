// Encoder self-attention on the warp-level tensor-core path (mma.sync): graph_encoder.py:83-104.
//
// The attention of one (instance, head) is two tiny GEMMs, S = Q K^T ([n,16] x [16,n]) and O = softmax(S) V
// ([n,n] x [n,16]) with n = N+1 <= 301 -- far too small for a tcgen05 tile pipeline, but exactly the shape of
// mma.sync fragments held in registers.  One warp owns 16 query rows; it walks the keys in chunks of CH score tiles
// of 8 keys with a base-2 online softmax between the two products, so the [n,n] score matrix lives only in
// registers.  fp32 accuracy comes from the same two-term operand split the tcgen05 GEMM uses,
// a*b ~= hi_a*hi_b + (lo_a*hi_b + hi_a*lo_b), with the cross products in their own accumulators so that they are not
// absorbed by the growing partial sums.  Shipped form (AGH_MHA_F16 = 1): fp16 (hi, lo' = (x - hi) 2^11) operands on
// m16n8k16 -- one instruction spans the head dimension of S and 16 keys of P V; AGH_MHA_F16 = 0: TF32 (hi, lo) on m16n8k8.
//
// One CTA walks the 8 heads of one instance.  K_h and V_h are rewritten in shared memory as (hi, lo') words in
// fragment order: the two B-fragment registers of every mma are adjacent (FP16 form -- K: dims {2t,2t+1 | 2t+8,2t+9} of
// one key, 96-byte key stride; V: keys {2t,2t+1 | 2t+8,2t+9} of one dim, 1 KB per 16 keys), so each fragment is ONE
// conflict-free LDS.64 straight into the register pair the instruction needs.  The P fragments of the second product
// are the score fragments themselves (FP16 form: a pair of score tiles is the m16k16 A fragment as it stands), so no
// shuffle is needed between the two products.
//
// Two kernels share that inner routine:
//  * mha_ws_kernel (default when it fits): warp-specialised.  One producer warp streams the raw K_h|V_h rows of the
//    next head with 16-byte cp.async into a double-buffered raw buffer (64-byte cp.async.bulk copies, 303 per head,
//    were 2x slower to issue) and converts them into the fragment-ordered buffer of a two-deep ring; the compute
//    warps read their Q fragments from global memory one head ahead and otherwise only wait on the ring's
//    full/empty mbarriers, so the HMMA pipe is never idled by a staging phase or a CTA-wide barrier.
//  * mha_mma_kernel: every warp stages and computes (cp.async prefetch of the next head, two __syncthreads per
//    head); used for small graphs (several heads per CTA) and when the ring does not fit in shared memory.
#include "common.cuh"
#include "gemm.cuh"
#include <math_constants.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace agh {

namespace {

constexpr int RAWS = 52;          // mha_mma_kernel raw row stride in floats: Q 16 | K 16 | V 16 | pad 4
constexpr int KVS = 36;           // mha_ws_kernel raw K|V row stride in floats: K 16 | V 16 | pad 4
// AGH_MHA_F16 = 1 (shipped): the same two-term FP16 split as the tcgen05 GEMM (x = hi + lo' / 2048, gemm_tc.cu) on
// mma.sync.m16n8k16: one instruction covers the whole head dimension of S = Q K^T and 16 keys of P V, i.e. HALF the
// tensor instructions of the m16n8k8 TF32 form at the same 11 + 11 significant bits.  AGH_MHA_F16 = 0 keeps the
// 3xTF32 form the scheme was validated against.
#ifndef AGH_MHA_F16
#define AGH_MHA_F16 1
#endif
#if AGH_MHA_F16
constexpr int KSTR = 12;          // split K row (one key), uint2 units: 4 hi {d2t,d2t+1 | d2t+8,d2t+9}, 4 lo', 4 pad (96 B: conflict-free)
constexpr int VGRP = 128;         // split V, per 16 keys: [d = 0,1][g = 0..7][t = 0..3] hi {keys 2t,2t+1 | 2t+8,2t+9} of dim 8d+g (64 units), then lo'
constexpr int KEY_PAIR = 2;       // chunk lengths are even: P V consumes the score tiles in pairs (16 keys per mma)
__host__ __device__ inline size_t split_bytes(int npad) { return (size_t)npad * KSTR * 8 + (size_t)(npad / 16) * VGRP * 8; }
#else
constexpr int KSTR = 20;          // split K row (one key): 8 hi pairs, 8 lo pairs, 4 pad
constexpr int VSTR = 36;          // split V row (two keys): 16 hi pairs, 16 lo pairs, 4 pad
constexpr int KEY_PAIR = 1;
__host__ __device__ inline size_t split_bytes(int npad) { return (size_t)npad * KSTR * 8 + (size_t)(npad / 2) * VSTR * 8; }
#endif
constexpr int MHA_MAX_THREADS = 640;
// norm_factor = 1/sqrt(16) (graph_encoder.py:40,89) and log2(e) go into q once: the softmax runs in base 2
constexpr float QSCALE = 0.25f * 1.4426950408889634f;

__device__ __forceinline__ uint32_t tf32_rn(float x) {          // round to nearest TF32 (ties away), finite inputs
  return (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void split_rn(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rn(x);
  lo = tf32_rn(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
#if AGH_MHA_F16
struct QFrag { uint32_t hi[4], lo[4]; };
__device__ __forceinline__ uint32_t pack_f16(float lo_elem, float hi_elem) {        // {lower half, upper half}
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ float2 unpack_f16(uint32_t h2) { return __half22float2(*reinterpret_cast<const __half2*>(&h2)); }
// (a, b) -> hi = f16x2(a, b), lo' = f16x2((a - hi_a) 2^11, (b - hi_b) 2^11)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16(a, b);
  const float2 hf = unpack_f16(hi);
  lo = pack_f16((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
// this lane's A fragment of 16 query rows x 16 dims from its 8 raw elements q[i] = {row r0 | r1}[i & 1], dims 2t + 8 (i >> 1) + {0, 1}
__device__ __forceinline__ void make_qfrag(const float2 (&q)[4], QFrag& f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair(q[i].x * QSCALE, q[i].y * QSCALE, f.hi[i], f.lo[i]);
}
// K_h / V_h rows [0, npad) (zeros beyond n) -> fragment-ordered FP16 (hi, lo') words; `raw` rows hold K at koff, V at voff
__device__ __forceinline__ void split_kv(const float* raw, int rstride, int koff, int voff, int n, int npad, uint2* ksp,
                                         uint2* vsp, int tid, int nthreads) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = tid; j < npad; j += nthreads) {                      // K: one key per thread
    const float4* src = reinterpret_cast<const float4*>(raw + j * rstride + koff);
    const bool ok = j < n;
    const float4 v0 = ok ? src[0] : z, v1 = ok ? src[1] : z, v2 = ok ? src[2] : z, v3 = ok ? src[3] : z;
    const float d[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
    uint32_t hw[8], lw[8];                                          // word 2t: dims 2t, 2t+1; word 2t+1: dims 2t+8, 2t+9
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      split_pair(d[2 * t], d[2 * t + 1], hw[2 * t], lw[2 * t]);
      split_pair(d[2 * t + 8], d[2 * t + 9], hw[2 * t + 1], lw[2 * t + 1]);
    }
    uint4* dst = reinterpret_cast<uint4*>(ksp + j * KSTR);
    dst[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); dst[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    dst[2] = make_uint4(lw[0], lw[1], lw[2], lw[3]); dst[3] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
  }
  uint32_t* vw = reinterpret_cast<uint32_t*>(vsp);
  for (int idx = tid; idx < npad * 2; idx += nthreads) {            // V: (keys 2 kp, 2 kp + 1; dims 4 c .. 4 c + 3)
    const int kp = idx >> 2, c = idx & 3;
    const float4 a = (2 * kp < n) ? *reinterpret_cast<const float4*>(raw + (2 * kp) * rstride + voff + c * 4) : z;
    const float4 b = (2 * kp + 1 < n) ? *reinterpret_cast<const float4*>(raw + (2 * kp + 1) * rstride + voff + c * 4) : z;
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    const int kg = kp >> 3, pp = kp & 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int dd = 4 * c + i;
      const int unit = kg * VGRP + (dd * 4) + (pp & 3);             // ((d * 8 + g) * 4 + t) with dd = 8 d + g
      uint32_t hi, lo;
      split_pair(av[i], bv[i], hi, lo);
      vw[unit * 2 + (pp >> 2)] = hi;
      vw[(unit + 64) * 2 + (pp >> 2)] = lo;
    }
  }
}

// One warp, 16 query rows (fragment q), all keys of one head: out[d][e] = normalised O in the mma C layout
// (e = 0,1: row r0, dims 8d + 2t, +1;  e = 2,3: row r0 + 8).  kr / vr point at this lane's first K / V fragment.
template <int CH>
__device__ __forceinline__ void attend(const uint2* kr, const uint2* vr, const QFrag& q, int n, int nchunk, int t,
                                       float (&out)[2][4]) {
  static_assert(CH % 2 == 0, "score tiles are consumed in pairs");
  constexpr float CS = 1.0f / 2048.0f;
  // O accumulators, dims [0,8) and [8,16): main product and the two (2^11-scaled) cross products
  float om[2][4], ox[2][4], oy[2][4];
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int e = 0; e < 4; ++e) { om[d][e] = 0.f; ox[d][e] = 0.f; oy[d][e] = 0.f; }
  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;

  for (int c = 0; c < nchunk; ++c, kr += CH * 8 * KSTR, vr += (CH / 2) * VGRP) {
    float s[CH][4];
    // ---- S = Q K^T for 8 CH keys: one k16 mma per product ----
#pragma unroll
    for (int jt = 0; jt < CH; ++jt) {
      const uint2* kt = kr + jt * 8 * KSTR;
      const uint2 kh = kt[0], kl = kt[4];
      float sx[4] = {0.f, 0.f, 0.f, 0.f};
      s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
      mma_f16(sx, q.lo, kh);
      mma_f16(sx, q.hi, kl);
      mma_f16(s[jt], q.hi, kh);
#pragma unroll
      for (int e = 0; e < 4; ++e) s[jt][e] = fmaf(sx[e], CS, s[jt][e]);
    }
    if (c == nchunk - 1) {                  // the last chunk holds the padding keys: mask them
      constexpr int TAIL = CH < 3 ? CH : 3;
      if (nchunk * CH * 8 - n <= 8 * TAIL) {      // ... which (for every graph size served) sit in its last three tiles:
#pragma unroll                                    // one uniform branch, static tile indices (masking all CH tiles costs 10 %
        for (int jt = CH - TAIL; jt < CH; ++jt) { // of the kernel's instructions at N = 100)
          const int key = (c * CH + jt) * 8 + 2 * t;
          if (key >= n) { s[jt][0] = -CUDART_INF_F; s[jt][2] = -CUDART_INF_F; }
          if (key + 1 >= n) { s[jt][1] = -CUDART_INF_F; s[jt][3] = -CUDART_INF_F; }
        }
      } else {
#pragma unroll
        for (int jt = 0; jt < CH; ++jt) {
          const int key = (c * CH + jt) * 8 + 2 * t;
          if (key >= n) { s[jt][0] = -CUDART_INF_F; s[jt][2] = -CUDART_INF_F; }
          if (key + 1 >= n) { s[jt][1] = -CUDART_INF_F; s[jt][3] = -CUDART_INF_F; }
        }
      }
    }
    // ---- online softmax: raise the running maxima of rows r0 / r0 + 8 ----
    float x0 = -CUDART_INF_F, x1 = -CUDART_INF_F;
#pragma unroll
    for (int jt = 0; jt < CH; ++jt) {
      x0 = fmaxf(x0, fmaxf(s[jt][0], s[jt][1]));
      x1 = fmaxf(x1, fmaxf(s[jt][2], s[jt][3]));
    }
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);        // finite: every chunk holds a valid key
    if (c > 0) {
      const float f0 = ex2(m0 - n0), f1 = ex2(m1 - n1);
      l0 *= f0; l1 *= f1;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        om[d][0] *= f0; om[d][1] *= f0; ox[d][0] *= f0; ox[d][1] *= f0; oy[d][0] *= f0; oy[d][1] *= f0;
        om[d][2] *= f1; om[d][3] *= f1; ox[d][2] *= f1; ox[d][3] *= f1; oy[d][2] *= f1; oy[d][3] *= f1;
      }
    }
    m0 = n0; m1 = n1;
    // ---- O += P V, 16 keys per mma: the two score tiles of a pair ARE the A fragment (rows g / g+8, keys 2t,2t+1 | 2t+8,2t+9) ----
#pragma unroll
    for (int jp = 0; jp < CH / 2; ++jp) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jt = 2 * jp + u;
        const float p0 = ex2(s[jt][0] - m0), p1 = ex2(s[jt][1] - m0), p2 = ex2(s[jt][2] - m1), p3 = ex2(s[jt][3] - m1);
        l0 += p0 + p1; l1 += p2 + p3;
        split_pair(p0, p1, ph[2 * u], pl[2 * u]);
        split_pair(p2, p3, ph[2 * u + 1], pl[2 * u + 1]);
      }
      const uint2* vt = vr + jp * VGRP;
      const uint2 vh0 = vt[0], vh1 = vt[32], vl0 = vt[64], vl1 = vt[96];
      mma_f16(om[0], ph, vh0);
      mma_f16(om[1], ph, vh1);
      mma_f16(ox[0], pl, vh0);
      mma_f16(ox[1], pl, vh1);
      mma_f16(oy[0], ph, vl0);
      mma_f16(oy[1], ph, vl1);
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    out[d][0] = fmaf(ox[d][0] + oy[d][0], CS, om[d][0]) * i0; out[d][1] = fmaf(ox[d][1] + oy[d][1], CS, om[d][1]) * i0;
    out[d][2] = fmaf(ox[d][2] + oy[d][2], CS, om[d][2]) * i1; out[d][3] = fmaf(ox[d][3] + oy[d][3], CS, om[d][3]) * i1;
  }
}
#else
// (a_i, b_i), i = 0..3  ->  4 hi pairs at dst[0..3], 4 lo pairs at dst[lo_off..lo_off+3]
__device__ __forceinline__ void store_pairs(uint2* dst, int lo_off, const float4 a, const float4 b) {
  uint32_t ah[4], al[4], bh[4], bl[4];
  split_rn(a.x, ah[0], al[0]); split_rn(a.y, ah[1], al[1]); split_rn(a.z, ah[2], al[2]); split_rn(a.w, ah[3], al[3]);
  split_rn(b.x, bh[0], bl[0]); split_rn(b.y, bh[1], bl[1]); split_rn(b.z, bh[2], bl[2]); split_rn(b.w, bh[3], bl[3]);
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = make_uint4(ah[0], bh[0], ah[1], bh[1]);
  d[1] = make_uint4(ah[2], bh[2], ah[3], bh[3]);
  uint4* e = reinterpret_cast<uint4*>(dst + lo_off);
  e[0] = make_uint4(al[0], bl[0], al[1], bl[1]);
  e[1] = make_uint4(al[2], bl[2], al[3], bl[3]);
}
// K_h / V_h rows [0, npad) (zeros beyond n) -> fragment-ordered (hi, lo) pairs; `raw` rows hold K at koff, V at voff
__device__ __forceinline__ void split_kv(const float* raw, int rstride, int koff, int voff, int n, int npad, uint2* ksp,
                                         uint2* vsp, int tid, int nthreads) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = tid; idx < npad * 2; idx += nthreads) {            // K: (key j, dims 8 k2 .. 8 k2 + 7)
    const int j = idx >> 1, k2 = idx & 1;
    const float4* src = reinterpret_cast<const float4*>(raw + j * rstride + koff + k2 * 8);
    const bool ok = j < n;
    store_pairs(ksp + j * KSTR + k2 * 4, 8, ok ? src[0] : z, ok ? src[1] : z);
  }
  for (int idx = tid; idx < npad * 2; idx += nthreads) {            // V: (keys 2 kp, 2 kp + 1; dims 4 c .. 4 c + 3)
    const int kp = idx >> 2, c = idx & 3;
    const float4 a = (2 * kp < n) ? *reinterpret_cast<const float4*>(raw + (2 * kp) * rstride + voff + c * 4) : z;
    const float4 b = (2 * kp + 1 < n) ? *reinterpret_cast<const float4*>(raw + (2 * kp + 1) * rstride + voff + c * 4) : z;
    store_pairs(vsp + kp * VSTR + c * 4, 16, a, b);
  }
}

// One warp, 16 query rows (fragments qhi/qlo), all keys of one head: out[d][e] = normalised O in the mma C layout
// (e = 0,1: row r0, dims 8d + 2t, +1;  e = 2,3: row r0 + 8).  kr / vr point at this lane's first K / V fragment.
template <int CH>
__device__ __forceinline__ void attend(const uint2* kr, const uint2* vr, const uint32_t (&qhi)[2][4],
                                       const uint32_t (&qlo)[2][4], int n, int nchunk, int t, float (&out)[2][4]) {
  // O accumulators, dims [0,8) and [8,16): main product and the two cross products (three independent mma chains)
  float om[2][4], ox[2][4], oy[2][4];
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int e = 0; e < 4; ++e) { om[d][e] = 0.f; ox[d][e] = 0.f; oy[d][e] = 0.f; }
  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;

  for (int c = 0; c < nchunk; ++c, kr += CH * 8 * KSTR, vr += CH * 4 * VSTR) {
    float s[CH][4];
    // ---- S = Q K^T for 8 CH keys ----
#pragma unroll
    for (int jt = 0; jt < CH; ++jt) {
      const uint2* kt = kr + jt * 8 * KSTR;
      const uint2 kh0 = kt[0], kh1 = kt[4], kl0 = kt[8], kl1 = kt[12];
      s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
      mma_tf32(s[jt], qlo[0], kh0);
      mma_tf32(s[jt], qhi[0], kl0);
      mma_tf32(s[jt], qlo[1], kh1);
      mma_tf32(s[jt], qhi[1], kl1);
      mma_tf32(s[jt], qhi[0], kh0);
      mma_tf32(s[jt], qhi[1], kh1);
    }
    if ((c + 1) * CH * 8 > n) {             // the chunk holds padding keys: mask them
#pragma unroll
      for (int jt = 0; jt < CH; ++jt) {
        const int key = (c * CH + jt) * 8 + 2 * t;
        if (key >= n) { s[jt][0] = -CUDART_INF_F; s[jt][2] = -CUDART_INF_F; }
        if (key + 1 >= n) { s[jt][1] = -CUDART_INF_F; s[jt][3] = -CUDART_INF_F; }
      }
    }
    // ---- online softmax: raise the running maxima of rows r0 / r0 + 8 ----
    float x0 = -CUDART_INF_F, x1 = -CUDART_INF_F;
#pragma unroll
    for (int jt = 0; jt < CH; ++jt) {
      x0 = fmaxf(x0, fmaxf(s[jt][0], s[jt][1]));
      x1 = fmaxf(x1, fmaxf(s[jt][2], s[jt][3]));
    }
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);        // finite: every chunk holds a valid key
    if (c > 0) {
      const float f0 = ex2(m0 - n0), f1 = ex2(m1 - n1);
      l0 *= f0; l1 *= f1;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        om[d][0] *= f0; om[d][1] *= f0; ox[d][0] *= f0; ox[d][1] *= f0; oy[d][0] *= f0; oy[d][1] *= f0;
        om[d][2] *= f1; om[d][3] *= f1; ox[d][2] *= f1; ox[d][3] *= f1; oy[d][2] *= f1; oy[d][3] *= f1;
      }
    }
    m0 = n0; m1 = n1;
    // ---- O += P V ----
#pragma unroll
    for (int jt = 0; jt < CH; ++jt) {
      const float p0 = ex2(s[jt][0] - m0), p1 = ex2(s[jt][1] - m0), p2 = ex2(s[jt][2] - m1), p3 = ex2(s[jt][3] - m1);
      l0 += p0 + p1; l1 += p2 + p3;
      // A fragment: column t <- key 2t, column t+4 <- key 2t+1 (the V pairs follow the same order)
      uint32_t ph[4], pl[4];
      ph[0] = __float_as_uint(p0) & 0xffffe000u; ph[1] = __float_as_uint(p2) & 0xffffe000u;
      ph[2] = __float_as_uint(p1) & 0xffffe000u; ph[3] = __float_as_uint(p3) & 0xffffe000u;
      pl[0] = __float_as_uint(p0 - __uint_as_float(ph[0])) & 0xffffe000u;
      pl[1] = __float_as_uint(p2 - __uint_as_float(ph[1])) & 0xffffe000u;
      pl[2] = __float_as_uint(p1 - __uint_as_float(ph[2])) & 0xffffe000u;
      pl[3] = __float_as_uint(p3 - __uint_as_float(ph[3])) & 0xffffe000u;
      const uint2* vt = vr + jt * 4 * VSTR;
      const uint2 vh0 = vt[0], vh1 = vt[8], vl0 = vt[16], vl1 = vt[24];
      mma_tf32(om[0], ph, vh0);
      mma_tf32(om[1], ph, vh1);
      mma_tf32(ox[0], pl, vh0);
      mma_tf32(ox[1], pl, vh1);
      mma_tf32(oy[0], ph, vl0);
      mma_tf32(oy[1], ph, vl1);
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    out[d][0] = (om[d][0] + (ox[d][0] + oy[d][0])) * i0; out[d][1] = (om[d][1] + (ox[d][1] + oy[d][1])) * i0;
    out[d][2] = (om[d][2] + (ox[d][2] + oy[d][2])) * i1; out[d][3] = (om[d][3] + (ox[d][3] + oy[d][3])) * i1;
  }
}

#endif

__device__ __forceinline__ void store_rows(float* heads_bh, int n, int r0, int t, const float (&out)[2][4]) {
  // heads_bh = heads + b * n * D + h * HD
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    if (r0 < n) *reinterpret_cast<float2*>(heads_bh + (size_t)r0 * D + d * 8 + 2 * t) = make_float2(out[d][0], out[d][1]);
    if (r0 + 8 < n) *reinterpret_cast<float2*>(heads_bh + (size_t)(r0 + 8) * D + d * 8 + 2 * t) = make_float2(out[d][2], out[d][3]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// warp-specialised kernel: grid = B CTAs, block = (mt + 1) warps; warps 0..mt-1 compute, warp mt produces
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t ws_smem_bytes(int n, int npad) {
  return 2 * (size_t)n * KVS * 4 + 2 * split_bytes(npad) + 64;
}

template <int CH, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) mha_ws_kernel(const float* __restrict__ qkv, int rs, size_t ps, int n, int mt,
                                                            float* __restrict__ heads) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nchunk = (n + 8 * CH - 1) / (8 * CH), npad = nchunk * 8 * CH;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (size_t)b * n * rs;      // row j, part p (Q/K/V), column c at base + j * rs + p * ps + c
  float* raw_kv = reinterpret_cast<float*>(smem_raw);                       // [2][n][KVS]  raw K_h | V_h rows
  uint2* ring = reinterpret_cast<uint2*>(raw_kv + 2 * n * KVS);             // [2]{ K [npad][KSTR], V [npad/2][VSTR] }
  const int ring_pairs = (int)(split_bytes(npad) / 8);
  uint64_t* split_full = reinterpret_cast<uint64_t*>(ring + 2 * ring_pairs);
  uint64_t* split_empty = split_full + 2;

  if (threadIdx.x == 0) {
    mbar_init(split_full, 1); mbar_init(split_full + 1, 1);
    mbar_init(split_empty, mt); mbar_init(split_empty + 1, mt);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == mt) {
    // ================= producer warp =================
    auto issue_load = [&](int h) {           // K_h | V_h rows -> raw_kv[h & 1], 8 chunks of 16 bytes per row
      float* dst = raw_kv + (h & 1) * n * KVS;
      for (int idx = lane; idx < n * 8; idx += 32) {
        const int j = idx >> 3, c = idx & 7;
        cp_async16(dst + j * KVS + c * 4, base + (size_t)j * rs + (size_t)(1 + (c >> 2)) * ps + h * HD + (c & 3) * 4);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue_load(0);
    for (int h = 0; h < H; ++h) {
      const int bf = h & 1, use = h >> 1;
      asm volatile("cp.async.wait_group 0;" ::: "memory");           // rows of head h have landed
      __syncwarp();                                                  // ... for every lane; head h-1's raw rows are dead
      if (h + 1 < H) issue_load(h + 1);                              // flies during the conversion below
      if (use >= 1) mbar_wait(split_empty + bf, (use - 1) & 1);      // compute warps are done with this ring slot
      uint2* ksp = ring + bf * ring_pairs;
      split_kv(raw_kv + bf * n * KVS, KVS, 0, HD, n, npad, ksp, ksp + npad * KSTR, lane, 32);
      __syncwarp();
      if (lane == 0) mbar_arrive(split_full + bf);                   // release: fragments visible to the consumers
    }
  } else {
    // ================= compute warps =================
    const int g = lane >> 2, t = lane & 3;
    const int r0 = warp * 16 + g, r1 = r0 + 8;
#if AGH_MHA_F16
    auto load_q = [&](int h, float2 (&q)[4]) {                       // this lane's A-fragment elements of Q_h
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (i & 1) ? r1 : r0;
        q[i] = (r < n) ? __ldg(reinterpret_cast<const float2*>(base + (size_t)r * rs + h * HD + 2 * t + 8 * (i >> 1))) : make_float2(0.f, 0.f);
      }
    };
    float2 qraw[4];
#else
    auto load_q = [&](int h, float (&q)[2][4]) {                     // this lane's A-fragment elements of Q_h
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = (e & 1) ? r1 : r0;
          q[ks][e] = (r < n) ? __ldg(base + (size_t)r * rs + h * HD + ks * 8 + t + ((e & 2) ? 4 : 0)) : 0.f;
        }
      }
    };
    float qraw[2][4];
#endif
    load_q(0, qraw);
    for (int h = 0; h < H; ++h) {
      const int bf = h & 1;
#if AGH_MHA_F16
      QFrag qf;
      make_qfrag(qraw, qf);
#else
      uint32_t qhi[2][4], qlo[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int e = 0; e < 4; ++e) split_rn(qraw[ks][e] * QSCALE, qhi[ks][e], qlo[ks][e]);
#endif
      if (h + 1 < H) load_q(h + 1, qraw);                            // in flight during this head
      mbar_wait(split_full + bf, (h >> 1) & 1);
      const uint2* ksp = ring + bf * ring_pairs;
      float out[2][4];
#if AGH_MHA_F16
      attend<CH>(ksp + g * KSTR + t, ksp + npad * KSTR + g * 4 + t, qf, n, nchunk, t, out);
#else
      attend<CH>(ksp + g * KSTR + t, ksp + npad * KSTR + t * VSTR + g, qhi, qlo, n, nchunk, t, out);
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive(split_empty + bf);                  // the ring slot may be refilled
      store_rows(heads + (size_t)b * n * D + h * HD, n, r0, t, out);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// uniform kernel: grid = B CTAs, block = hpc * mt warps.  The CTA's warps form hpc "head lanes" of mt warps; lane hl
// walks heads hl, hl + hpc, ... of the instance, warp (w % mt) of a lane owning query rows [16 (w % mt), +16).  The raw
// rows of the next head arrive by cp.async while the current head is computed.
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t lane_bytes(int npad) {
  return (size_t)npad * RAWS * 4 + split_bytes(npad);
}

template <int CH, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) mha_mma_kernel(const float* __restrict__ qkv, int rs, size_t ps, int n, int hpc, int mt,
                                                             float* __restrict__ heads) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nchunk = (n + 8 * CH - 1) / (8 * CH), npad = nchunk * 8 * CH;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int hl = warp / mt, mi = warp - hl * mt;
  const int lt = threadIdx.x - hl * mt * 32, lthreads = mt * 32;
  const float* base = qkv + (size_t)b * n * rs;      // row j, part p (Q/K/V), column c at base + j * rs + p * ps + c
  float* raw = reinterpret_cast<float*>(smem_raw + hl * lane_bytes(npad));
  uint2* ksp = reinterpret_cast<uint2*>(raw + npad * RAWS);
  uint2* vsp = ksp + npad * KSTR;
  const int r0 = mi * 16 + g, r1 = r0 + 8;
  const int nit = H / hpc;

  auto issue_load = [&](int h) {            // rows [0, n) of Q_h | K_h | V_h, 12 chunks of 16 bytes per row
    for (int idx = lt; idx < n * 12; idx += lthreads) {
      const int j = idx / 12, c = idx - j * 12;
      cp_async16(raw + j * RAWS + c * 4, base + (size_t)j * rs + (size_t)(c >> 2) * ps + h * HD + (c & 3) * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_load(hl);

  for (int it = 0; it < nit; ++it) {
    const int h = hl + it * hpc;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                        // raw rows of head h visible; previous head's fragments no longer read
#if AGH_MHA_F16
    QFrag qf;
    {
      float2 qraw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (i & 1) ? r1 : r0;
        qraw[i] = (r < n) ? *reinterpret_cast<const float2*>(raw + r * RAWS + 2 * t + 8 * (i >> 1)) : make_float2(0.f, 0.f);
      }
      make_qfrag(qraw, qf);
    }
#else
    uint32_t qhi[2][4], qlo[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = (e & 1) ? r1 : r0;
        const float q = (r < n) ? raw[r * RAWS + ks * 8 + t + ((e & 2) ? 4 : 0)] : 0.f;
        split_rn(q * QSCALE, qhi[ks][e], qlo[ks][e]);
      }
    }
#endif
    split_kv(raw, RAWS, HD, 2 * HD, n, npad, ksp, vsp, lt, lthreads);
    __syncthreads();                        // fragments visible; raw buffer free for the next head
    if (it + 1 < nit) issue_load(h + hpc);
    float out[2][4];
#if AGH_MHA_F16
    attend<CH>(ksp + g * KSTR + t, vsp + g * 4 + t, qf, n, nchunk, t, out);
#else
    attend<CH>(ksp + g * KSTR + t, vsp + t * VSTR + g, qhi, qlo, n, nchunk, t, out);
#endif
    store_rows(heads + (size_t)b * n * D + h * HD, n, r0, t, out);
  }
}

template <int CH, int MAXT, int MINB>
int launch_uniform(const float* qkv, int rs, size_t ps, int B, int n, int hpc, int mt, float* heads, cudaStream_t st) {
  const int npad = ceil_div(n, 8 * CH) * 8 * CH;
  const size_t smem = (size_t)hpc * lane_bytes(npad);
  // the opt-in is a per-device attribute: set it on every launch (a process may drive several GPUs)
  if (smem > 48 * 1024)
    AGH_CUDA(cudaFuncSetAttribute(mha_mma_kernel<CH, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mha_mma_kernel<CH, MAXT, MINB><<<B, hpc * mt * 32, smem, st>>>(qkv, rs, ps, n, hpc, mt, heads);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

template <int CH, int MAXT, int MINB>
int launch_ws(const float* qkv, int rs, size_t ps, int B, int n, int mt, float* heads, cudaStream_t st) {
  const int npad = ceil_div(n, 8 * CH) * 8 * CH;
  const size_t smem = ws_smem_bytes(n, npad);
  // the opt-in is a per-device attribute: set it on every launch (a process may drive several GPUs)
  if (smem > 48 * 1024)
    AGH_CUDA(cudaFuncSetAttribute(mha_ws_kernel<CH, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mha_ws_kernel<CH, MAXT, MINB><<<B, (mt + 1) * 32, smem, st>>>(qkv, rs, ps, n, mt, heads);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

// CTAs of up to 8 warps (N <= 111 with the producer warp) may use up to 128 registers; larger graphs need up to 20.
template <int CH>
int launch_ch(const float* qkv, int rs, size_t ps, int B, int n, int hpc, int mt, float* heads, cudaStream_t st) {
  static const bool no_ws = getenv("AGH_MHA_WS") != nullptr && atoi(getenv("AGH_MHA_WS")) == 0;
  const int npad = ceil_div(n, 8 * CH) * 8 * CH;
  if (!no_ws && hpc == 1 && ws_smem_bytes(n, npad) <= 227 * 1024) {
    if (mt + 1 <= 8) return launch_ws<CH, 256, 2>(qkv, rs, ps, B, n, mt, heads, st);
    return launch_ws<CH, MHA_MAX_THREADS, 1>(qkv, rs, ps, B, n, mt, heads, st);
  }
  if (hpc * mt <= 8) return launch_uniform<CH, 256, 2>(qkv, rs, ps, B, n, hpc, mt, heads, st);
  return launch_uniform<CH, MHA_MAX_THREADS, 1>(qkv, rs, ps, B, n, hpc, mt, heads, st);
}

}  // namespace

int launch_mha_mma(const float* qkv, bool slab_major, int B, int n, float* heads, cudaStream_t st) {
  // row-major [B*n][384] rows Q|K|V, or slab-major: three [B*n][128] matrices (the tensor-core QKV GEMM's output form)
  const int rs = slab_major ? D : 3 * D;
  const size_t ps = slab_major ? (size_t)B * n * D : (size_t)D;
  const int mt = ceil_div(n, 16);
  int hpc = 1;
  while (hpc < H && 2 * hpc * mt <= 8) hpc *= 2;                 // about 8 warps per CTA for small graphs
  AGH_CHECK_ARG((hpc * mt + 1) * 32 <= MHA_MAX_THREADS);
  // chunk length: least padded work (about 60 instructions per key fragment, 60 per chunk for the rescale)
#if AGH_MHA_F16
  const int ntile = 2 * ceil_div(n, 16);                         // score tiles come in pairs (16 keys per P V mma)
  // CTAs of more than 8 warps have 102 registers per thread: chunks of at most 8 tiles there (longer ones spill)
  const int ch_max = (hpc * mt + 1) * 32 <= 256 ? 14 : 8;
  int best = 8, best_cost = 1 << 30;
  for (int ch = ch_max; ch >= 4; ch -= 2) {
    const int chunks = ceil_div(ntile, ch), cost = chunks * ch * 60 + chunks * 60;
    if (cost < best_cost) { best = ch; best_cost = cost; }
  }
  switch (best) {
    case 4: return launch_ch<4>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    case 6: return launch_ch<6>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    case 8: return launch_ch<8>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    case 10: return launch_ch<10>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    case 12: return launch_ch<12>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    default: return launch_ch<14>(qkv, rs, ps, B, n, hpc, mt, heads, st);
  }
#else
  const int ntile = ceil_div(n, 8);
  int best = 8, best_cost = 1 << 30;
  for (int ch = 8; ch >= 3; --ch) {
    const int chunks = ceil_div(ntile, ch), cost = chunks * ch * 60 + chunks * 60;
    if (cost < best_cost) { best = ch; best_cost = cost; }
  }
  switch (best) {
    case 3: return launch_ch<3>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    case 4: return launch_ch<4>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    case 5: return launch_ch<5>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    case 6: return launch_ch<6>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    case 7: return launch_ch<7>(qkv, rs, ps, B, n, hpc, mt, heads, st);
    default: return launch_ch<8>(qkv, rs, ps, B, n, hpc, mt, heads, st);
  }
#endif
}

}  // namespace agh
