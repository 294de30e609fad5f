// Sample-shared-key decode (second half): the `batch_rep` replicas of one instance that
// AttentionModel.sample_many rolls out (attention_model.py:358-370, utils/functions.py:171-188 `do_batch_rep`) share the
// SAME glimpse keys, values, logit keys and step-context table -- only the vehicle state differs.  The per-instance kernel
// (decode.cu) spends one CTA, 155 KB of staging and three block barriers per step on every replica; here ONE CTA stages the
// instance's matrices once and each of its WARPS owns one replica for all T steps:
//   * no block barrier in the step loop -- a replica's step is a chain of warp-synchronous phases (__syncwarp only), and
//     4..16 independent replicas per SM hide each other's latencies;
//   * no redundant work -- the state update is done once per replica (every warp of the per-instance kernel repeats it);
//   * every phase walks the compact list of the step's FEASIBLE nodes only;
//   * phase mappings chosen so that no cross-lane reduction of the 128-wide glimpse is needed:
//       scores   lane = (head h = lane / 4, r = lane % 4): q_h in 16 registers, nodes r, r+4, ... of the list, K' rows swizzled;
//                softmax statistics across the 4 lanes of a head (two shuffles)
//       values   lane = feature chunk (4 of the 128 features, head lane / 4): o_chunk = sum_j e_hj V_j[chunk], a whole
//                512-byte V row per warp load, weights broadcast from the warp's score buffer -- no butterfly at all
//       logits   8 lanes per node, 4 nodes per pass (as in decode.cu), o chunks in registers
//       select   lane = list slot: tanh, Gumbel noise (the SAME Philox4x32-10 stream keyed by (seed, instance id, step,
//                node) as decode.cu: both kernels draw identical noise), arg-max by two REDUX
// Feasibility mask and state transition use the reference's exact mixed f32 / f64 arithmetic (state_agh.py:99-174), as in
// decode.cu.  Sampling mode, N <= 127, temperature >= 0.25; everything else takes the per-instance kernel.
#include "common.cuh"
#include <math_constants.h>

namespace agh {

namespace sh {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

// Philox4x32-10, identical to decode.cu: counter = (instance id lo/hi, step, node), key = seed
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint64_t id, uint32_t step, uint32_t node) {
  uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32), c2 = step, c3 = node;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}
__device__ __forceinline__ float gumbel(uint64_t seed, uint64_t id, uint32_t step, uint32_t node) {
  const float u = ((float)philox_u32(seed, id, step, node) + 0.5f) * 2.3283064365386963e-10f;  // (0,1)
  return -logf(-logf(fminf(u, 0.99999994f)));
}
__device__ __forceinline__ int fkey(float x) {
  int k = __float_as_int(x);
  return k ^ ((k >> 31) & 0x7fffffff);
}
__device__ __forceinline__ uint64_t pack2(float x, float y) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float sum2(uint64_t v) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
  return x + y;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int NPL, typename T>
__device__ __forceinline__ T pick(const T (&v)[NPL], int idx) {
  T r = v[0];
#pragma unroll
  for (int p = 1; p < NPL; ++p) r = (idx == p) ? v[p] : r;
  return r;
}

struct Layout {     // byte offsets into dynamic shared memory
  size_t K, V, L, P, twr, dur, twl, dem, crd, bar, warp0, per_warp, total;
};

__host__ __device__ inline Layout layout(int n, int npl, bool resp, int warps) {
  Layout s{};
  const size_t npad = (size_t)npl * 32, mat = (size_t)n * KS * 4;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  s.K = take(mat); s.V = take(mat); s.L = take(mat); s.P = take(resp ? mat : 0);
  s.twr = take(npad * 8); s.dur = take(npad * 8); s.twl = take(npad * 4); s.dem = take(npad * 4); s.crd = take(npad * 4);
  s.bar = take(16);
  // per warp: scores / softmax weights [8][npad] f32 | query [128] | glimpse [128] | pre-tanh logits [npad] | list [npad] u16
  s.per_warp = (8 * npad * 4 + 128 * 4 + 128 * 4 + npad * 4 + npad * 2 + 15) & ~(size_t)15;
  s.warp0 = off;
  s.total = off + (size_t)warps * s.per_warp;
  return s;
}

template <int NPL, bool RESP>
__global__ void __launch_bounds__(512, 1) decode_shared_kernel(const agh_decode_args a, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int n = a.N + 1;
  const int warps = blockDim.x >> 5;
  const int KBn = a.key_mod;
  const int kb = blockIdx.x % KBn, grp = blockIdx.x / KBn;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NPAD = NPL * 32;
  const Layout Lo = layout(n, NPL, RESP, warps);
  float* sK = reinterpret_cast<float*>(smem + Lo.K);
  float* sV = reinterpret_cast<float*>(smem + Lo.V);
  float* sL = reinterpret_cast<float*>(smem + Lo.L);
  float* sP = reinterpret_cast<float*>(smem + Lo.P);
  double* s_twr = reinterpret_cast<double*>(smem + Lo.twr);
  double* s_dur = reinterpret_cast<double*>(smem + Lo.dur);
  float* s_twl = reinterpret_cast<float*>(smem + Lo.twl);
  float* s_dem = reinterpret_cast<float*>(smem + Lo.dem);
  int* s_crd = reinterpret_cast<int*>(smem + Lo.crd);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Lo.bar);
  unsigned char* mine = smem + Lo.warp0 + (size_t)warp * Lo.per_warp;
  float* s_a = reinterpret_cast<float*>(mine);                 // [8][NPAD]: scores, then un-normalised softmax weights
  float* s_q = s_a + 8 * NPAD;                                 // [128]
  float* s_o = s_q + D;                                        // [128]
  float* s_u = s_o + D;                                        // [NPAD] pre-tanh logits by list slot
  uint16_t* s_idx = reinterpret_cast<uint16_t*>(s_u + NPAD);   // [NPAD] feasible nodes, increasing index

  const size_t KB = (size_t)KBn;
  const float* gK = a.keys + (size_t)kb * n * KS;
  const float* gV = a.keys + (KB + kb) * n * KS;
  const float* gL = a.keys + (2 * KB + kb) * n * KS;
  const float* gpsc = a.psc + (size_t)kb * n * D;

  // ---- stage the instance's matrices ONCE for all replicas of this CTA (TMA bulk copies, one mbarrier) ----
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t mat = (uint32_t)n * KS * 4;
    mbar_expect_tx(bar, (RESP ? 4u : 3u) * mat);
    bulk_g2s(sK, gK, mat, bar);
    bulk_g2s(sV, gV, mat, bar);
    bulk_g2s(sL, gL, mat, bar);
    if (RESP) bulk_g2s(sP, gpsc, mat, bar);
  }
  for (int j = tid; j < NPAD; j += blockDim.x) {
    const bool ok = j < n;
    s_twr[j] = ok ? a.tw_right[(size_t)kb * n + j] : 0.0;
    s_dur[j] = ok ? a.duration[(size_t)kb * n + j] : 0.0;
    s_twl[j] = ok ? a.tw_left[(size_t)kb * n + j] : 0.f;
    s_dem[j] = (ok && j >= 1) ? a.demand[(size_t)kb * a.N + j - 1] : 0.f;
    s_crd[j] = ok ? (int)a.coord[(size_t)kb * n + j] : 0;
  }
  __syncthreads();
  mbar_wait(bar, 0);
  // K' rows: XOR-swizzle the 16-byte chunks of every aligned 8-chunk group by (node & 7) in place, so that the score
  // phase (lanes of one quarter-warp read the same chunk of different rows) is bank-conflict free
  for (int it = tid; it < n * 4; it += blockDim.x) {
    const int j = it >> 2, x = j & 7;
    if (x == 0) continue;
    float4* grp4 = reinterpret_cast<float4*>(sK + (size_t)j * KS) + (it & 3) * 8;
    const int rot = lane & 7;
    float4 tmp[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) tmp[c] = grp4[(c + rot) & 7];
#pragma unroll
    for (int c = 0; c < 8; ++c) grp4[((c + rot) & 7) ^ x] = tmp[c];
  }
  __syncthreads();

  const int rep = grp * warps + warp;                           // replica of this warp
  if (rep >= reps) return;
  const int b = rep * KBn + kb;                                 // row of the outputs (replica-major, as decode.cu)
  const size_t trow = (size_t)b * a.t_stride;
  const uint64_t rng_id = (uint64_t)(a.id_offset + b);
  const bool sampling = a.mode == AGH_SAMPLING;
  const bool unit_temp = a.temp == 1.0f;
  const float shift = __fdiv_rn(10.0f, a.temp);                 // log-sum-exp shift: logits / temp <= shift

  // per-lane constants: my nodes lane + 32 p, and my 4 query components 4 lane .. 4 lane + 3
  int cj[NPL];
  float demn[NPL], twl[NPL];
  double twr[NPL], dur[NPL];
  uint32_t visbits = 0u;
#pragma unroll
  for (int p = 0; p < NPL; ++p) {
    const int j = lane + 32 * p;
    cj[p] = s_crd[j]; demn[p] = s_dem[j]; twl[p] = s_twl[j]; twr[p] = s_twr[j]; dur[p] = s_dur[j];
  }
  const float4 qb4 = *reinterpret_cast<const float4*>(a.qbase + (size_t)kb * D + 4 * lane);
  const float4 wc4 = *reinterpret_cast<const float4*>(a.wblob + W_CAP + 4 * lane);
  const float4 wt4 = *reinterpret_cast<const float4*>(a.wblob + W_TIME + 4 * lane);
  for (int j = lane; j < n; j += 32) a.serve[(size_t)b * n + j] = 0.f;
  __syncwarp();

  // state (replicated in every lane): state_agh.py:60-93
  int prev = 0, t = 0, nvis = 0;
  bool depot_seen = false;
  float used = 0.f, len = 0.f, ll = 0.f;
  double cft = -60.0;
  float ddm[NPL], ttm[NPL];
#pragma unroll
  for (int p = 0; p < NPL; ++p) {
    const bool ok = lane + 32 * p < n;
    ddm[p] = ok ? __ldg(a.distance + NODE_SIZE * s_crd[0] + cj[p]) : 0.f;
    ttm[p] = ok ? __ldg(a.travel + NODE_SIZE * s_crd[0] + cj[p]) : 0.f;
  }
  float4 pv4 = RESP ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4*>(gpsc) + lane);

  const int h = lane >> 2, r4 = lane & 3;           // score phase: head / position in the head's lane group; value phase: chunk = lane
  const int ns = lane >> 3, fg = lane & 7;          // logit phase: node of the pass / chunk group
  const float4* K4 = reinterpret_cast<const float4*>(sK);
  const float4* V4 = reinterpret_cast<const float4*>(sV) + lane;
  const float4* L4 = reinterpret_cast<const float4*>(sL) + fg;

  while (t < a.t_stride && nvis < n) {
    // (1) feasibility of my nodes: state_agh.py:148-174, exact dtypes
    uint32_t mw[NPL];
    double fin[NPL];
    uint32_t feas = 0u;
#pragma unroll
    for (int p = 0; p < NPL; ++p) {
      const int j = lane + 32 * p;
      bool mj = true;
      fin[p] = 0.0;
      if (j < n && j >= 1) {
        const bool cap = __fadd_rn(demn[p], used) > 1.0f;
        fin[p] = __dadd_rn(fmax(__dadd_rn(cft, (double)ttm[p]), (double)twl[p]), dur[p]);
        mj = ((visbits >> p) & 1u) | cap | (fin[p] > twr[p]);
      }
      mw[p] = __ballot_sync(0xffffffffu, mj);
      feas |= (p == 0) ? (~mw[p] & 0xfffffffeu) : ~mw[p];
    }
    mw[0] = (mw[0] & 0xfffffffeu) | ((prev == 0 && feas != 0u) ? 1u : 0u);      // depot: state_agh.py:173
    // (2) compact list of the feasible nodes (increasing index)
    int nfeas = 0;
#pragma unroll
    for (int p = 0; p < NPL; ++p) {
      const uint32_t fw = ~mw[p];
      if ((fw >> lane) & 1u) s_idx[nfeas + __popc(fw & ((1u << lane) - 1u))] = (uint16_t)(32 * p + lane);
      nfeas += __popc(fw);
    }
    // (3) query, 4 components per lane: attention_model.py:432-435,510-522
    {
      const double inv = 1.0 / 1440.0;
      const double q0 = cft * inv;
      const float tf = (float)fma(fma(-q0, 1440.0, cft), inv, q0);
      const float rem = 1.0f - used;
      const float4 pv = RESP ? *(reinterpret_cast<const float4*>(sP + (size_t)prev * D) + lane) : pv4;
      float4 q;
      q.x = fmaf(wt4.x, tf, fmaf(wc4.x, rem, qb4.x + pv.x)); q.y = fmaf(wt4.y, tf, fmaf(wc4.y, rem, qb4.y + pv.y));
      q.z = fmaf(wt4.z, tf, fmaf(wc4.z, rem, qb4.z + pv.z)); q.w = fmaf(wt4.w, tf, fmaf(wc4.w, rem, qb4.w + pv.w));
      reinterpret_cast<float4*>(s_q)[lane] = q;
    }
    __syncwarp();

    // (4) compatibilities of head h with the feasible nodes r4, r4 + 4, ... (K' holds 1/sqrt(16) and log2(e))
    float inv_sum;
    {
      uint64_t qh[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 qv = reinterpret_cast<const float4*>(s_q + h * HD)[i];
        qh[2 * i] = pack2(qv.x, qv.y);
        qh[2 * i + 1] = pack2(qv.z, qv.w);
      }
      float* sa = s_a + h * NPAD;
      float mx = -CUDART_INF_F;
      for (int c = r4; c < nfeas; c += 4) {
        const int j = s_idx[c];
        const float4* kr = K4 + j * (KS / 4);
        uint64_t acc = 0ull;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 kv = kr[(h * 4 + i) ^ (j & 7)];
          acc = fma2(qh[2 * i], pack2(kv.x, kv.y), acc);
          acc = fma2(qh[2 * i + 1], pack2(kv.z, kv.w), acc);
        }
        const float s = sum2(acc);
        sa[c] = s;
        mx = fmaxf(mx, s);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
      for (int c = r4; c < nfeas; c += 4) {        // every lane re-reads only what it wrote
        const float e = exp2f(sa[c] - mx);
        sa[c] = e;
        sum += e;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      inv_sum = __frcp_rn(sum);                    // the 4 lanes of head h hold the same value; they own chunks 4h..4h+3 below
    }
    __syncwarp();

    // (5) glimpse: lane = feature chunk `lane` (head lane / 4 = h): o_chunk = inv_sum * sum_j e_hj V_j[chunk]
    {
      const float* sa = s_a + h * NPAD;
      uint64_t a01 = 0ull, a23 = 0ull, b01 = 0ull, b23 = 0ull;
      int c = 0;
      for (; c + 2 <= nfeas; c += 2) {
        const int j0 = s_idx[c], j1 = s_idx[c + 1];
        const float e0 = sa[c], e1 = sa[c + 1];
        const float4 v0 = V4[j0 * (KS / 4)], v1 = V4[j1 * (KS / 4)];
        const uint64_t ee0 = pack2(e0, e0), ee1 = pack2(e1, e1);
        a01 = fma2(ee0, pack2(v0.x, v0.y), a01); a23 = fma2(ee0, pack2(v0.z, v0.w), a23);
        b01 = fma2(ee1, pack2(v1.x, v1.y), b01); b23 = fma2(ee1, pack2(v1.z, v1.w), b23);
      }
      if (c < nfeas) {
        const int j0 = s_idx[c];
        const float e0 = sa[c];
        const float4 v0 = V4[j0 * (KS / 4)];
        const uint64_t ee0 = pack2(e0, e0);
        a01 = fma2(ee0, pack2(v0.x, v0.y), a01); a23 = fma2(ee0, pack2(v0.z, v0.w), a23);
      }
      float ax, ay, az, aw, bx, by, bz, bw;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(ax), "=f"(ay) : "l"(a01));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(az), "=f"(aw) : "l"(a23));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(bx), "=f"(by) : "l"(b01));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(bz), "=f"(bw) : "l"(b23));
      reinterpret_cast<float4*>(s_o)[lane] = make_float4((ax + bx) * inv_sum, (ay + by) * inv_sum, (az + bz) * inv_sum, (aw + bw) * inv_sum);
    }
    __syncwarp();

    // (6) logits of the feasible nodes: 8 lanes per node, lane fg takes chunks {fg, fg+8, fg+16, fg+24} of o and of LK'_j
    {
      uint64_t o01[4], o23[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 ov = reinterpret_cast<const float4*>(s_o)[fg + 8 * k];
        o01[k] = pack2(ov.x, ov.y);
        o23[k] = pack2(ov.z, ov.w);
      }
      for (int c0 = 0; c0 < nfeas; c0 += 4) {
        const int c = c0 + ns;
        const bool valid = c < nfeas;
        const int j = valid ? (int)s_idx[c] : 0;
        const float4* lr = L4 + j * (KS / 4);
        uint64_t ua = 0ull, ub = 0ull;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 lv = lr[8 * k];
          ua = fma2(o01[k], pack2(lv.x, lv.y), ua);
          ub = fma2(o23[k], pack2(lv.z, lv.w), ub);
        }
        float u = sum2(ua) + sum2(ub);
        u += __shfl_xor_sync(0xffffffffu, u, 1);
        u += __shfl_xor_sync(0xffffffffu, u, 2);
        u += __shfl_xor_sync(0xffffffffu, u, 4);
        if (valid && fg == 0) s_u[c] = u;
      }
    }
    __syncwarp();

    // (7) clip, temperature, Gumbel-max selection (lane = list slot): attention_model.py:446-447,580,372-390
    int best_key = (int)0x80000000, best_j = 0x7fffffff;
    float best_z = 0.f, ws = 0.f;
    for (int c = lane; c < nfeas; c += 32) {
      const int j = s_idx[c];
      float z = 10.0f * tanhf(s_u[c]);
      if (!unit_temp) z = __fdiv_rn(z, a.temp);
      const float sc = sampling ? z + gumbel(a.seed, rng_id, (uint32_t)t, (uint32_t)j) : z;
      ws += expf(z - shift);
      const int key = fkey(sc);
      if (key > best_key) { best_key = key; best_j = j; best_z = z; }       // slots are in index order: first index wins ties
    }
    const int kmax = __reduce_max_sync(0xffffffffu, best_key);
    const int sel = __reduce_min_sync(0xffffffffu, best_key == kmax ? best_j : 0x7fffffff);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, off);
    const int src = __ffs(__ballot_sync(0xffffffffu, best_key == kmax && best_j == sel)) - 1;
    const float z_sel = __shfl_sync(0xffffffffu, best_z, src);
    ll += (z_sel - shift) - logf(ws);

    // (8) state update: state_agh.py:99-137 (the owner lane of the selected node holds its distance, demand and finish time)
    const int p_sel = sel >> 5, l_sel = sel & 31;
    const float d = __shfl_sync(0xffffffffu, pick<NPL>(ddm, p_sel), l_sel);
    const float dm = __shfl_sync(0xffffffffu, pick<NPL>(demn, p_sel), l_sel);
    const double fsel = __shfl_sync(0xffffffffu, pick<NPL>(fin, p_sel), l_sel);
    len = __fadd_rn(len, d);
    if (sel != 0) {
      used = __fadd_rn(used, dm);
      cft = fsel;
      ++nvis;
    } else {
      used = 0.f;
      cft = -60.0;
      nvis += depot_seen ? 0 : 1;
      depot_seen = true;
    }
    if (lane == l_sel) visbits |= 1u << p_sel;
    if (lane == 0) {
      a.pi[trow + t] = (int16_t)sel;
      a.serve[(size_t)b * n + sel] = (float)cft;
    }
    {   // next step's rows come through L2: requested now, consumed after the next mask / query phases start
      const int cnext = s_crd[sel];
#pragma unroll
      for (int p = 0; p < NPL; ++p) {
        const bool ok = lane + 32 * p < n;
        ddm[p] = ok ? __ldg(a.distance + NODE_SIZE * cnext + cj[p]) : 0.f;
        ttm[p] = ok ? __ldg(a.travel + NODE_SIZE * cnext + cj[p]) : 0.f;
      }
      if (!RESP) pv4 = __ldg(reinterpret_cast<const float4*>(gpsc + (size_t)sel * D) + lane);
    }
    prev = sel;
    ++t;
    __syncwarp();      // the warp's buffers are rewritten by the next step
  }

  // ---- epilogue: closed-tour length (problem_agh.py:49-50,66), outputs, zero padding ----
  if (lane == 0) {
    a.cost[b] = __fadd_rn(len, __ldg(a.distance + NODE_SIZE * s_crd[prev]));
    a.ll[b] = ll;
    a.steps[b] = t;
    if (a.status != nullptr) a.status[b] = nvis < n ? AGH_DECODE_STUCK : 0;
  }
  for (int k = t + lane; k < a.t_stride; k += 32) a.pi[trow + k] = 0;
}

template <int NPL>
static int launch(const agh_decode_args& a, int reps, cudaStream_t st, bool* taken) {
  const int n = a.N + 1;
  int dev = 0, max_smem = 0;
  AGH_CUDA(cudaGetDevice(&dev));
  AGH_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // Replicas per CTA (= warps, 4..16): the FEWEST that still put all replicas on the GPU in one wave (fewer warps per SM
  // = a shorter step for each; e.g. 1 instance x 1280 replicas -> 143 CTAs of 9 warps instead of 80 of 16), else 16.
  // The step-context table is resident when it fits.
  int sms = 148;
  AGH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int sm_smem = 0;
  AGH_CUDA(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  for (int resp = 1; resp >= 0; --resp) {
    int pick = 0;
    for (int warps = 4; warps <= 16; ++warps) {
      const size_t bytes = layout(n, NPL, resp != 0, warps).total;
      if ((int)bytes > max_smem) break;
      pick = warps;                                                          // the largest that fits, unless one wave is reached
      int per_sm = sm_smem / ((int)bytes + 1024);
      per_sm = per_sm < 1 ? 1 : per_sm;
      if (per_sm * warps > 16) per_sm = 16 / warps;                          // 512 threads x 128 registers fill the register file
      const long long grid = (long long)a.key_mod * ((reps + warps - 1) / warps);
      if (grid <= (long long)sms * per_sm) break;
    }
    if (pick == 0) continue;
    const int warps = pick;
    const size_t bytes = layout(n, NPL, resp != 0, warps).total;
    const long long grid = (long long)a.key_mod * ((reps + warps - 1) / warps);
    if (grid > 2147483647LL) continue;
    if (resp) {
      AGH_CUDA(cudaFuncSetAttribute(decode_shared_kernel<NPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      decode_shared_kernel<NPL, true><<<(unsigned)grid, warps * 32, bytes, st>>>(a, reps);
    } else {
      AGH_CUDA(cudaFuncSetAttribute(decode_shared_kernel<NPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      decode_shared_kernel<NPL, false><<<(unsigned)grid, warps * 32, bytes, st>>>(a, reps);
    }
    AGH_LAUNCH_CHECK();
    *taken = true;
    return AGH_OK;
  }
  *taken = false;
  return AGH_OK;
}

}  // namespace sh

// Called by agh_decode: takes the launch when the arguments describe shared-key sampling that this kernel supports.
int try_decode_shared(const agh_decode_args& a, cudaStream_t st, bool* taken) {
  *taken = false;
  const bool dbg = a.mode == AGH_REPLAY || a.inject_logp || a.inject_z || a.logp_sel || a.logp_full || a.mask_dump || a.used_dump ||
                   a.cft_dump || a.q_dump || a.attn_dump || a.o_dump || a.lse_dump;
  if (dbg || a.mode != AGH_SAMPLING || a.key_mod <= 0 || a.B % a.key_mod != 0 || a.N > 126 || a.temp < 0.25f) return AGH_OK;
  if ((a.smem_mats & 64) != 0) return AGH_OK;                  // caller forces the per-instance kernel (A/B tests)
  const int reps = a.B / a.key_mod;
  if (reps < 4) return AGH_OK;
  const int npl = ceil_div(a.N + 1, 32);
  switch (npl) {
    case 1: return sh::launch<1>(a, reps, st, taken);
    case 2: return sh::launch<2>(a, reps, st, taken);
    case 3: return sh::launch<3>(a, reps, st, taken);
    case 4: return sh::launch<4>(a, reps, st, taken);
    default: return AGH_OK;
  }
}

}  // namespace agh
