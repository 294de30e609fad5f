// Encoder self-attention on the warp-level tensor-core path (mma.sync m16n8k8, 3xTF32): graph_encoder.py:83-104.
//
// The attention of one (instance, head) is two tiny GEMMs, S = Q K^T ([n,16] x [16,n]) and O = softmax(S) V
// ([n,n] x [n,16]) with n = N+1 <= 301 -- far too small for a tcgen05 tile pipeline, but exactly the shape of
// mma.sync fragments held in registers.  One warp owns 16 query rows; it walks the keys in chunks of CH fragments
// of 8 keys with a base-2 online softmax between the two products, so the [n,n] score matrix lives only in
// registers.  fp32 accuracy comes from the same 3xTF32 split the tcgen05 GEMM uses: x = hi + lo with hi, lo
// representable in TF32, and  a*b ~= hi_a*hi_b + (lo_a*hi_b + hi_a*lo_b).  For S (a 16-term sum) the cross products
// are issued first so the small terms enter the accumulator before the large ones; for O (a sum over all keys) they
// have their own accumulator so that they are not absorbed by the growing partial sums.
//
// Shared memory holds K_h and V_h of the CTA's heads already split, as (hi, lo) pairs, so the inner loop has no
// conversion work for them and every fragment element is one LDS.64.  Row strides of 20 (K) and 18 (V) pairs make
// both fragment access patterns bank-conflict free.  The P fragments of the second product are the score
// fragments themselves: a thread's score columns (2t, 2t+1) are fed as A-columns (t, t+4), and the V fragment is
// loaded with rows (2t, 2t+1) to match, so no shuffle is needed between the two products.
#include "common.cuh"
#include "gemm.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace agh {

namespace {

constexpr int KSTR = 20;          // K row stride, in (hi, lo) pairs
constexpr int VSTR = 18;          // V row stride, in (hi, lo) pairs
constexpr int MHA_MAX_THREADS = 640;
#ifndef MHA_MINB_SMALL
#define MHA_MINB_SMALL 2
#endif

__device__ __forceinline__ uint32_t tf32_rn(float x) {          // round to nearest TF32 (ties away), finite inputs
  return (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void split_rn(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rn(x);
  lo = tf32_rn(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void stage_row(float2* dst, const float4 v) {       // 4 values -> 4 (hi, lo) pairs
  uint32_t hi[4], lo[4];
  split_rn(v.x, hi[0], lo[0]); split_rn(v.y, hi[1], lo[1]); split_rn(v.z, hi[2], lo[2]); split_rn(v.w, hi[3], lo[3]);
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = make_uint4(hi[0], lo[0], hi[1], lo[1]);
  d[1] = make_uint4(hi[2], lo[2], hi[3], lo[3]);
}

// grid = B * (H / hpc) CTAs, block = hpc * mt warps: warp w works on head (w / mt) of the CTA and query rows
// [16 (w % mt), 16 (w % mt) + 16).  CH = key fragments (8 keys each) per online-softmax chunk; the shared-memory
// rows are zero-padded to whole chunks, so the chunk body is branch-free.
template <int CH, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) mha_mma_kernel(const float* __restrict__ qkv, int n, int hpc, int mt,
                                                                  float* __restrict__ heads) {
  extern __shared__ __align__(16) float2 sm2[];
  const int nchunk = (n + 8 * CH - 1) / (8 * CH), npad = nchunk * 8 * CH;
  const int groups = H / hpc;
  const int b = blockIdx.x / groups;
  const int h0 = (blockIdx.x % groups) * hpc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float* base = qkv + (size_t)b * n * (3 * D);
  const int head_pairs = npad * (KSTR + VSTR);

  // ---- Q fragments of this warp's 16 rows, straight from global memory (in flight during the staging below) ----
  const int hl = warp / mt, mi = warp - hl * mt;
  const int h = h0 + hl;
  const int r0 = mi * 16 + g, r1 = r0 + 8;
  float qraw[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = (e & 1) ? r1 : r0;
      const int col = ks * 8 + t + ((e & 2) ? 4 : 0);
      qraw[ks][e] = (r < n) ? __ldg(base + (size_t)r * 3 * D + h * HD + col) : 0.f;
    }
  }

  // ---- stage K_h, V_h of the CTA's heads, split into (hi, lo); two rows in flight per thread ----
  const int items = hpc * npad * 4;
  for (int idx = threadIdx.x; idx < items; idx += 2 * blockDim.x) {
    const int idx2 = idx + blockDim.x;
    const int hla = idx / (npad * 4), ra = idx - hla * npad * 4;
    const int hlb = idx2 / (npad * 4), rb = idx2 - hlb * npad * 4;
    const int ja = ra >> 2, ca = ra & 3, jb = rb >> 2, cb = rb & 3;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ka = z, va = z, kb = z, vb = z;
    if (ja < n) {
      const float* p = base + (size_t)ja * 3 * D + D + (h0 + hla) * HD + ca * 4;
      ka = *reinterpret_cast<const float4*>(p);
      va = *reinterpret_cast<const float4*>(p + D);
    }
    if (idx2 < items && jb < n) {
      const float* p = base + (size_t)jb * 3 * D + D + (h0 + hlb) * HD + cb * 4;
      kb = *reinterpret_cast<const float4*>(p);
      vb = *reinterpret_cast<const float4*>(p + D);
    }
    float2* sKa = sm2 + hla * head_pairs;
    stage_row(sKa + ja * KSTR + ca * 4, ka);
    stage_row(sKa + npad * KSTR + ja * VSTR + ca * 4, va);
    if (idx2 < items) {
      float2* sKb = sm2 + hlb * head_pairs;
      stage_row(sKb + jb * KSTR + cb * 4, kb);
      stage_row(sKb + npad * KSTR + jb * VSTR + cb * 4, vb);
    }
  }

  // norm_factor = 1/sqrt(16) (graph_encoder.py:40,89) and log2(e) go into q once: the softmax runs in base 2
  const float qs = 0.25f * 1.4426950408889634f;
  uint32_t qhi[2][4], qlo[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) split_rn(qraw[ks][e] * qs, qhi[ks][e], qlo[ks][e]);
  __syncthreads();

  const uint2* kr = reinterpret_cast<const uint2*>(sm2 + hl * head_pairs) + g * KSTR + t;
  const uint2* vr = reinterpret_cast<const uint2*>(sm2 + hl * head_pairs) + npad * KSTR + 2 * t * VSTR + g;
  float om[2][4], ox[2][4];                 // O accumulators: main and cross products, dims [0,8) and [8,16)
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int e = 0; e < 4; ++e) { om[d][e] = 0.f; ox[d][e] = 0.f; }
  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;

  for (int c = 0; c < nchunk; ++c, kr += CH * 8 * KSTR, vr += CH * 8 * VSTR) {
    float s[CH][4];
    // ---- S = Q K^T for 8 CH keys ----
#pragma unroll
    for (int jt = 0; jt < CH; ++jt) {
      const uint2 k00 = kr[jt * 8 * KSTR], k01 = kr[jt * 8 * KSTR + 4], k10 = kr[jt * 8 * KSTR + 8], k11 = kr[jt * 8 * KSTR + 12];
      s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
      mma_tf32(s[jt], qlo[0], k00.x, k01.x);
      mma_tf32(s[jt], qhi[0], k00.y, k01.y);
      mma_tf32(s[jt], qlo[1], k10.x, k11.x);
      mma_tf32(s[jt], qhi[1], k10.y, k11.y);
      mma_tf32(s[jt], qhi[0], k00.x, k01.x);
      mma_tf32(s[jt], qhi[1], k10.x, k11.x);
    }
    if ((c + 1) * CH * 8 > n) {               // the chunk holds padding keys: mask them
#pragma unroll
      for (int jt = 0; jt < CH; ++jt) {
        const int key = (c * CH + jt) * 8 + 2 * t;
        if (key >= n) { s[jt][0] = -CUDART_INF_F; s[jt][2] = -CUDART_INF_F; }
        if (key + 1 >= n) { s[jt][1] = -CUDART_INF_F; s[jt][3] = -CUDART_INF_F; }
      }
    }
    // ---- online softmax: raise the running maxima of rows r0 / r1 ----
    float x0 = -CUDART_INF_F, x1 = -CUDART_INF_F;
#pragma unroll
    for (int jt = 0; jt < CH; ++jt) {
      x0 = fmaxf(x0, fmaxf(s[jt][0], s[jt][1]));
      x1 = fmaxf(x1, fmaxf(s[jt][2], s[jt][3]));
    }
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);          // finite: every chunk holds a valid key
    if (c > 0) {
      const float f0 = ex2(m0 - n0), f1 = ex2(m1 - n1);
      l0 *= f0; l1 *= f1;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        om[d][0] *= f0; om[d][1] *= f0; ox[d][0] *= f0; ox[d][1] *= f0;
        om[d][2] *= f1; om[d][3] *= f1; ox[d][2] *= f1; ox[d][3] *= f1;
      }
    }
    m0 = n0; m1 = n1;
    // ---- O += P V ----
#pragma unroll
    for (int jt = 0; jt < CH; ++jt) {
      const float p0 = ex2(s[jt][0] - m0), p1 = ex2(s[jt][1] - m0), p2 = ex2(s[jt][2] - m1), p3 = ex2(s[jt][3] - m1);
      l0 += p0 + p1; l1 += p2 + p3;
      // A fragment: column t <- key 2t, column t+4 <- key 2t+1 (the V rows below follow the same order)
      uint32_t ph[4], pl[4];
      ph[0] = __float_as_uint(p0) & 0xffffe000u; ph[1] = __float_as_uint(p2) & 0xffffe000u;
      ph[2] = __float_as_uint(p1) & 0xffffe000u; ph[3] = __float_as_uint(p3) & 0xffffe000u;
      pl[0] = __float_as_uint(p0 - __uint_as_float(ph[0])) & 0xffffe000u;
      pl[1] = __float_as_uint(p2 - __uint_as_float(ph[1])) & 0xffffe000u;
      pl[2] = __float_as_uint(p1 - __uint_as_float(ph[2])) & 0xffffe000u;
      pl[3] = __float_as_uint(p3 - __uint_as_float(ph[3])) & 0xffffe000u;
      const uint2 v00 = vr[jt * 8 * VSTR], v01 = vr[jt * 8 * VSTR + VSTR], v10 = vr[jt * 8 * VSTR + 8], v11 = vr[jt * 8 * VSTR + VSTR + 8];
      mma_tf32(om[0], ph, v00.x, v01.x);
      mma_tf32(om[1], ph, v10.x, v11.x);
      mma_tf32(ox[0], pl, v00.x, v01.x);
      mma_tf32(ox[1], pl, v10.x, v11.x);
      mma_tf32(ox[0], ph, v00.y, v01.y);
      mma_tf32(ox[1], ph, v10.y, v11.y);
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    if (r0 < n)
      *reinterpret_cast<float2*>(heads + ((size_t)b * n + r0) * D + h * HD + d * 8 + 2 * t) =
          make_float2((om[d][0] + ox[d][0]) * i0, (om[d][1] + ox[d][1]) * i0);
    if (r1 < n)
      *reinterpret_cast<float2*>(heads + ((size_t)b * n + r1) * D + h * HD + d * 8 + 2 * t) =
          make_float2((om[d][2] + ox[d][2]) * i1, (om[d][3] + ox[d][3]) * i1);
  }
}

template <int CH, int MAXT, int MINB>
int launch_ch2(const float* qkv, int B, int n, int hpc, int mt, float* heads, cudaStream_t st) {
  const int npad = ceil_div(n, 8 * CH) * 8 * CH;
  const size_t smem = (size_t)hpc * npad * (KSTR + VSTR) * sizeof(float2);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    AGH_CUDA(cudaFuncSetAttribute(mha_mma_kernel<CH, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  mha_mma_kernel<CH, MAXT, MINB><<<B * (H / hpc), hpc * mt * 32, smem, st>>>(qkv, n, hpc, mt, heads);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

// CTAs of up to 8 warps (N <= 127) may use up to 128 registers; larger graphs need up to 19 warps per CTA.
template <int CH>
int launch_ch(const float* qkv, int B, int n, int hpc, int mt, float* heads, cudaStream_t st) {
  static const int minb = getenv("AGH_MHA_MINB") ? atoi(getenv("AGH_MHA_MINB")) : MHA_MINB_SMALL;
  if (hpc * mt <= 8 && minb == 3) return launch_ch2<CH, 256, 3>(qkv, B, n, hpc, mt, heads, st);
  if (hpc * mt <= 8) return launch_ch2<CH, 256, MHA_MINB_SMALL>(qkv, B, n, hpc, mt, heads, st);
  return launch_ch2<CH, MHA_MAX_THREADS, 1>(qkv, B, n, hpc, mt, heads, st);
}

}  // namespace

int launch_mha_mma(const float* qkv, int B, int n, float* heads, cudaStream_t st) {
  const int mt = ceil_div(n, 16);
  int hpc = 1;
  while (hpc < H && 2 * hpc * mt <= 8) hpc *= 2;                 // about 8 warps per CTA for small graphs
  AGH_CHECK_ARG(hpc * mt * 32 <= MHA_MAX_THREADS);
  // chunk length: least padded work (about 100 instructions per key fragment, 60 per chunk for the rescale)
  const int ntile = ceil_div(n, 8);
  int best = 8, best_cost = 1 << 30;
  for (int ch = 8; ch >= 3; --ch) {
    const int chunks = ceil_div(ntile, ch), cost = chunks * ch * 100 + chunks * 60;
    if (cost < best_cost) { best = ch; best_cost = cost; }
  }
  switch (best) {
    case 3: return launch_ch<3>(qkv, B, n, hpc, mt, heads, st);
    case 4: return launch_ch<4>(qkv, B, n, hpc, mt, heads, st);
    case 5: return launch_ch<5>(qkv, B, n, hpc, mt, heads, st);
    case 6: return launch_ch<6>(qkv, B, n, hpc, mt, heads, st);
    case 7: return launch_ch<7>(qkv, B, n, hpc, mt, heads, st);
    default: return launch_ch<8>(qkv, B, n, hpc, mt, heads, st);
  }
}

}  // namespace agh
