// Fused persistent decode kernel: one CTA per (instance, fleet) runs ALL decode steps.
//
// Replaces, per step, attention_model.py:429-451 (_get_log_p), :510-522 (step context), :550-584
// (_one_to_many_logits), :372-390 (_select_node), state_agh.py:148-174 (get_mask) and :99-137
// (update) -- ~93 ATen launches and 4 host syncs per step in the reference -- and accumulates
// the log-likelihood (attention_model.py:238-250) and the tour length (problem_agh.py:66) on the
// fly so the [B,T,n] log-prob tensor is never materialised.
//
// Data layout:
//   * the instance's K' rows are loaded ONCE into registers (16 NPL per thread) and its V / LK' rows
//     ([n][128] f32, plain rows written by agh_precompute) are staged ONCE into shared memory with
//     cp.async.bulk (TMA bulk copy, mbarrier completion) and chunk-swizzled in place; all stay on chip for all T steps.
//     P_sc = W_sc[:, :128] h (the node part of the step-context projection) is staged too when that
//     does not cost a resident CTA.  At N = 100 this is 111 KB + 64 registers per instance, so TWO
//     instances share an SM and hide each other's step latency; N = 300 streams V through L2.
//   * W_out is folded into LK' (LK' = W_out^T W_lk / sqrt(128)), so a step is
//       q = qbase + P_sc[prev] + w_cap (1-used) + w_time (cft/1440)
//       s_hj = q_h . K'_jh ; a_h = softmax_j(s_h | mask) ; o_h = sum_j a_hj V_jh
//       u_j = 10 tanh(o . LK'_j) ; log_p = log_softmax(u / temp | mask)
//   * 8 warps = 8 heads for the glimpse (warp-local softmax, no block sync), then 16 nodes per
//     warp for the logits; three __syncthreads per step.  The kernel is bound by the latency of this
//     per-step dependency chain, so the chain is kept short: integer REDUX for every max / arg-max, the
//     next step's L2 reads (distance row, P_sc component) issued the moment the node is selected, the
//     selected node's finish time read back from the thread that evaluated it for the mask.
//   * graphs with more than 127 nodes walk a compact list of the step's feasible nodes in the value
//     and logit phases (feasibility is sparse: about 30 % of the nodes on average).
//   * feasibility mask and state transition follow the reference's mixed f32/f64 arithmetic
//     exactly (SURVEY.md section 8a numerics contract): f32 table lookup, f32 IEEE division by
//     110, f64 time accumulation, f32 capacity.
#include "common.cuh"
#include <math_constants.h>

namespace agh {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
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

// Philox4x32-10 (Salmon et al. 2011); counter = (instance id lo/hi, step, node), key = seed.
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

// Packed FP32 pairs (sm_100a fma.rn.f32x2 -> SASS FFMA2): two IEEE FMAs per issued instruction.  Measured on B200
// (tools/micro/ffma2.cu) the FP32 FLOP rate is the same as FFMA's, so the gain is issue slots and chain length: the
// 16-term dot products below are 8 dependent FFMA2 instead of 16 dependent FFMA.
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
__device__ __forceinline__ void unpack2(uint64_t v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

template <int NPL>
__device__ __forceinline__ uint32_t pick_word(const uint32_t (&w)[NPL], int idx) {
  uint32_t r = 0;
#pragma unroll
  for (int p = 0; p < NPL; ++p) r = (idx == p) ? w[p] : r;
  return r;
}

// order-preserving float <-> int map (its own inverse): integer REDUX / compares on float scores
__device__ __forceinline__ int fkey(float x) {
  int k = __float_as_int(x);
  return k ^ ((k >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

// Greedy selection is argmax_j exp(log_p_j) with the first index winning ties (attention_model.py:333,377), and
// log_p_j = (z_j - max z) - log(sum) is rounded on the grid of log(sum) (at most 2^-21 for n <= 301), which collapses
// logits closer than that to the maximum into a tie.  Logits further than NEAR_TIE below the maximum can never tie with
// it (>= 4 grid steps in every binade, and exp keeps them apart), so the fast path selects on z itself and only steps
// with a second logit inside the window re-decide with the reference's formula (rare: ~1e-5 of the steps).
constexpr float NEAR_TIE = 2e-6f;

// Residency of the per-instance matrices.  K' always lives in REGISTERS (lane l of warp h holds the 16
// features of head h for its nodes l, l+32, ...: 16*NPL registers, the exact operand layout of the
// compatibility dot products).  V / LK' / P_sc live in shared memory when the bit is set, else they are
// read through L2 every step.
enum { RES_V = 1, RES_L = 2, RES_P = 4 };

struct DecodeSmem {
  // byte offsets into dynamic shared memory.  The small per-step arrays come FIRST, at offsets that depend on the
  // template parameter NPL only (compile-time immediates in the step loop); the key matrices follow.
  size_t twr, dur, twl, dem, crd, dcur, z, serve, o, qb, red, mw, idx, fin, bar, V, L, P, total;
};

__host__ __device__ constexpr size_t al16(size_t b) { return (b + 15) & ~(size_t)15; }

__host__ __device__ inline DecodeSmem decode_smem_layout(int n, int npl, int res) {
  DecodeSmem s{};
  const size_t npad = (size_t)npl * 32;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += al16(bytes); return o; };
  s.o = take(D * 4); s.qb = take(D * 4);          // qb slot holds the step's query vector
  s.red = take(8 * 4 * 4);
  s.mw = take(16 * 4);
  s.bar = take(16);
  s.z = take(npad * 4); s.dcur = take(2 * npad * 4);
  s.fin = take(2 * npad * 8);                     // finish time of every node if it were served next (f64), double-buffered
  s.dem = take(npad * 4); s.crd = take(npad * 4); s.serve = take(npad * 4);
  s.idx = take(npad * 2);                         // compact list of the step's feasible nodes (u16)
  s.twr = take(npad * 8); s.dur = take(npad * 8); s.twl = take(npad * 4);
  off = (off + 127) & ~(size_t)127;
  s.V = take((res & RES_V) ? (size_t)n * KS * 4 : 0);
  s.L = take((res & RES_L) ? (size_t)n * KS * 4 : 0);
  s.P = take((res & RES_P) ? (size_t)n * D * 4 : 0);
  s.total = off;
  return s;
}

// compile-time offsets of the small arrays (must mirror decode_smem_layout)
template <int NPL>
struct SmallOff {
  static constexpr size_t npad = (size_t)NPL * 32;
  static constexpr size_t o = 0;
  static constexpr size_t qb = o + al16(D * 4);
  static constexpr size_t red = qb + al16(D * 4);
  static constexpr size_t mw = red + al16(8 * 4 * 4);
  static constexpr size_t bar = mw + al16(16 * 4);
  static constexpr size_t z = bar + al16(16);
  static constexpr size_t dcur = z + al16(npad * 4);
  static constexpr size_t fin = dcur + al16(2 * npad * 4);
  static constexpr size_t dem = fin + al16(2 * npad * 8);
  static constexpr size_t crd = dem + al16(npad * 4);
  static constexpr size_t serve = crd + al16(npad * 4);
  static constexpr size_t idx = serve + al16(npad * 4);
  static constexpr size_t twr = idx + al16(npad * 2);
  static constexpr size_t dur = twr + al16(npad * 8);
  static constexpr size_t twl = dur + al16(npad * 8);
  static constexpr size_t mats = (twl + al16(npad * 4) + 127) & ~(size_t)127;
};

// V rows (and, on the compact path, LK' rows) are held in shared memory unpadded ([n][128] f32) with the 16-byte chunks
// of every aligned 8-chunk group XOR-swizzled by (node & 7): lanes that read the same logical chunk of 8 consecutive
// nodes hit 8 different bank groups, so every LDS.128 is conflict-free without padding bytes.
__device__ __forceinline__ int swz(int chunk, int node) { return chunk ^ (node & 7); }

// resident CTAs per SM the register budget is compiled for (65536 / (256 * regs))
constexpr int decode_min_blocks(int npl) { return npl == 1 ? 3 : (npl <= 4 ? 2 : 1); }

// NPL: nodes per lane in the glimpse (ceil(n/32)); RES: which of {V, LK', P_sc} live in smem; DBG: the variant that
// carries the optional paths (replay, logit / log-prob injection, per-step dumps, full log-prob output) -- the
// production greedy / sampling rollout is compiled without them.
template <int NPL, int RES, bool DBG>
__global__ void __launch_bounds__(256, decode_min_blocks(NPL)) decode_kernel(const agh_decode_args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int n = a.N + 1;
  const int b = blockIdx.x;
  const int kb = a.key_mod > 0 ? b % a.key_mod : b;      // source of keys / node vectors
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  using SO = SmallOff<NPL>;
  constexpr int NPAD = NPL * 32;
  constexpr int MPT = (NPAD + 255) / 256;                 // mask nodes per thread (node = tid + 256 k)
  // large graphs walk a compact list of the feasible nodes in the value / logit phases (measured: N = 300 +41 %,
  // N = 200 +3 %, N <= 100 -4 %: there the fully unrolled fixed mapping wins)
  constexpr bool COMPACT = NPL > 4;

  float* sV = reinterpret_cast<float*>(smem + SO::mats);
  float* sL = sV + ((RES & RES_V) ? (size_t)n * KS : 0);
  float* sP = sL + ((RES & RES_L) ? (size_t)n * KS : 0);
  double* s_twr = reinterpret_cast<double*>(smem + SO::twr);
  double* s_dur = reinterpret_cast<double*>(smem + SO::dur);
  float* s_twl = reinterpret_cast<float*>(smem + SO::twl);
  float* s_dem = reinterpret_cast<float*>(smem + SO::dem);
  int* s_crd = reinterpret_cast<int*>(smem + SO::crd);
  float* s_dcur = reinterpret_cast<float*>(smem + SO::dcur);
  float* s_z = reinterpret_cast<float*>(smem + SO::z);
  float* s_serve = reinterpret_cast<float*>(smem + SO::serve);
  float* s_o = reinterpret_cast<float*>(smem + SO::o);
  float* s_q = reinterpret_cast<float*>(smem + SO::qb);        // query of the current step, written by threads 128..255
  float* s_red = reinterpret_cast<float*>(smem + SO::red);
  uint32_t* s_mw = reinterpret_cast<uint32_t*>(smem + SO::mw);
  uint16_t* s_idx = reinterpret_cast<uint16_t*>(smem + SO::idx);
  double* s_fin = reinterpret_cast<double*>(smem + SO::fin);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SO::bar);

  // keys = [3][KB][n][128]: K', V, LK' as three plain row-major matrices (written by agh_precompute's GEMM epilogue)
  const size_t KB = a.key_mod > 0 ? (size_t)a.key_mod : (size_t)a.B;
  const float* gK = a.keys + (size_t)kb * n * KS;
  const float* gV = a.keys + (KB + kb) * n * KS;
  const float* gL = a.keys + (2 * KB + kb) * n * KS;
  const float* gpsc = a.psc + (size_t)kb * n * D;

  // ---- stage the key matrices with TMA bulk copies (one mbarrier, one phase) ----
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t mat_bytes = (uint32_t)n * KS * 4;
    uint32_t total = 0;
    if (RES & RES_V) total += mat_bytes;
    if (RES & RES_L) total += mat_bytes;
    if (RES & RES_P) total += (uint32_t)n * D * 4;
    mbar_expect_tx(bar, total);
    if (RES & RES_V) bulk_g2s(sV, gV, mat_bytes, bar);
    if (RES & RES_L) bulk_g2s(sL, gL, mat_bytes, bar);
    if (RES & RES_P) bulk_g2s(sP, gpsc, (uint32_t)n * D * 4, bar);
  }
  // K' -> registers: the 16 features of head `warp` for nodes lane + 32 p (one-time global read), as FFMA2 operand pairs
  uint64_t kreg[NPL][HD / 2];
#pragma unroll
  for (int p = 0; p < NPL; ++p) {
    const int j = lane + 32 * p;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < n) v = __ldg(reinterpret_cast<const float4*>(gK + (size_t)j * KS) + warp * 4 + i);
      kreg[p][2 * i] = pack2(v.x, v.y);
      kreg[p][2 * i + 1] = pack2(v.z, v.w);
    }
  }
  // node vectors and per-instance constants with ordinary loads while the bulk copies fly
  for (int j = tid; j < NPAD; j += 256) {
    const bool ok = j < n;
    s_twr[j] = ok ? a.tw_right[(size_t)kb * n + j] : 0.0;
    s_dur[j] = ok ? a.duration[(size_t)kb * n + j] : 0.0;
    s_twl[j] = ok ? a.tw_left[(size_t)kb * n + j] : 0.f;
    s_dem[j] = (ok && j >= 1) ? a.demand[(size_t)kb * a.N + j - 1] : 0.f;
    s_crd[j] = ok ? (int)a.coord[(size_t)kb * n + j] : 0;
    s_serve[j] = 0.f;
  }
  // threads 128..255 each own one component of the query (base, capacity and time coefficients in registers): at
  // N <= 127 the feasibility of the nodes is evaluated by threads 0..127, so the two halves of a step's first phase run
  // side by side in different warps instead of one after the other in the same ones
  const int qi = tid - (256 - D);
  float qb_r = 0.f, wc_r = 0.f, wt_r = 0.f;
  if (qi >= 0) {
    qb_r = a.qbase[(size_t)kb * D + qi];
    wc_r = a.wblob[W_CAP + qi];
    wt_r = a.wblob[W_TIME + qi];
  }
  __syncthreads();
  mbar_wait(bar, 0);
  // the staged rows are plain: permute the 16-byte chunks of every aligned 8-chunk group by (node & 7) in place (swz()),
  // so that lanes reading the same logical chunk of 8 consecutive nodes hit 8 different bank groups.  LK' needs it on
  // the compact path only: the fixed logit mapping reads whole 128-byte groups of one row per quarter-warp.
  constexpr bool SWV = (RES & RES_V) != 0, SWL = COMPACT && (RES & RES_L) != 0;     // global rows are read as they are
  {
    auto swizzle_rows = [&](float* m) {
      for (int it = tid; it < n * 4; it += 256) {
        const int j = it >> 2, x = j & 7;
        if (x == 0) continue;
        float4* grp = reinterpret_cast<float4*>(m + (size_t)j * KS) + (it & 3) * 8;
        const int rot = lane & 7;                 // lanes start at different chunks: no bank conflicts in this pass either
        float4 tmp[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) tmp[c] = grp[(c + rot) & 7];
#pragma unroll
        for (int c = 0; c < 8; ++c) grp[((c + rot) & 7) ^ x] = tmp[c];
      }
    };
    if (SWV) swizzle_rows(sV);
    if (SWL) swizzle_rows(sL);
    __syncthreads();
  }

  const float* Vm = (RES & RES_V) ? sV : gV;
  const float* Lm = (RES & RES_L) ? sL : gL;
  const float* Pm = (RES & RES_P) ? sP : gpsc;

  // ---- constants of the node(s) whose feasibility this thread evaluates (node = tid + 256 k) ----
  int cjm[MPT];
  float demm[MPT], twlm[MPT];
  double twrm[MPT], durm[MPT];
  bool vis[MPT];
#pragma unroll
  for (int k = 0; k < MPT; ++k) {
    const int j = tid + 256 * k;
    const bool ok = j < NPAD;
    cjm[k] = ok ? s_crd[j] : 0; demm[k] = ok ? s_dem[j] : 0.f; twlm[k] = ok ? s_twl[j] : 0.f;
    twrm[k] = ok ? s_twr[j] : 0.0; durm[k] = ok ? s_dur[j] : 0.0;
    vis[k] = false;
  }

  // ---- state (replicated in every thread) : state_agh.py:60-93 ----
  int prev = 0, t = 0, nvis = 0;
  bool depot_seen = false;
  float used = 0.f, len = 0.f, ll = 0.f;
  double cft = -60.0;

  const int h = warp;
  const bool replay = DBG && a.mode == AGH_REPLAY;
  const bool sampling = a.mode == AGH_SAMPLING;
  const uint64_t rng_id = (uint64_t)(a.id_offset + b);
  const size_t trow = (size_t)b * a.t_stride;
  const int nwords = (n + 31) / 32;
  const bool unit_temp = a.temp == 1.0f;
  // 10 tanh(.) / temp <= shift: a fixed log-sum-exp shift needs no cross-warp max (exp(-2 shift) must stay
  // normal, hence temp >= 0.25); colder temperatures take the exact max-merging path
  const bool fixed_shift = a.temp >= 0.25f;
  const float shift = __fdiv_rn(10.0f, a.temp);
  const bool tie_check = a.mode == AGH_GREEDY && !(DBG && a.inject_logp != nullptr);     // production greedy selection

  // ---- per-lane row pointers of the fixed (non-compact) value / logit mappings: every LDS.128 of the step loop is
  //      [register + immediate].  The last 32-node pass may run past n: those lanes read a clamped (valid) row; their
  //      softmax weight is 0 and their logit is masked.
  //      value phase: lane <-> node lane + 32 p, chunks 4h..4h+3 of the swizzled V row;
  //      logit phase: 4 nodes per warp and pass, 8 lanes per node; lane fg takes chunks {fg, fg+8, fg+16, fg+24} of the
  //      plain LK' row (a quarter-warp reads 128 contiguous bytes: conflict-free) and keeps the matching chunks of o.
  const int ns = lane >> 3, fg = lane & 7;
  const float4* vb[4];
  const float4* vb_last[4];
  {
    const int jl = min(lane + 32 * (NPL - 1), n - 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      vb[i] = reinterpret_cast<const float4*>(Vm + (size_t)lane * KS) + (SWV ? swz(h * 4 + i, lane) : h * 4 + i);
      vb_last[i] = reinterpret_cast<const float4*>(Vm + (size_t)jl * KS) + (SWV ? swz(h * 4 + i, jl) : h * 4 + i);
    }
  }
  const float4* lbase = reinterpret_cast<const float4*>(Lm) + fg;

  // (1) distance and travel time (distance / 110 in f32, precomputed table: bit-identical to the reference's division)
  //     from the current location to the node(s) this thread masks, and this thread's component of the current node's
  //     P_sc row when that table is not in shared memory: all come through L2, so they are requested as soon as the node
  //     is known (end of the previous step) and land during the log-likelihood / state update
  float ddm[MPT], ttm[MPT];
#pragma unroll
  for (int k = 0; k < MPT; ++k) {
    const bool ok = tid + 256 * k < n;
    ddm[k] = ok ? __ldg(a.distance + NODE_SIZE * s_crd[0] + cjm[k]) : 0.f;
    ttm[k] = ok ? __ldg(a.travel + NODE_SIZE * s_crd[0] + cjm[k]) : 0.f;
  }
  float pv_r = 0.f;
  if (!(RES & RES_P) && qi >= 0) pv_r = __ldg(Pm + qi);

  // the step budget bounds the loop: an instance whose unvisited nodes are all infeasible (demand > capacity, a time
  // window that cannot be met from the depot) would select the depot forever -- the reference spins in its Python loop;
  // here the instance stops with AGH_DECODE_STUCK in its status word
  const int tmax = replay ? a.replay_steps : a.t_stride;
  while (t < tmax && (replay || nvis < n)) {
    float* dc = s_dcur + (t & 1) * NPAD;
    // like dc, double-buffered by step parity: a fast warp's writes of step t+1 must not overtake a slow warp's
    // reads of step t (no barrier separates a step's selection phase from the next step's mask phase)
    double* fin_t = s_fin + (t & 1) * NPAD;

    // (2) one component of the query per thread (threads 128..255), published through shared memory:
    //     attention_model.py:432-435,510-522
    if (qi >= 0) {
      // (float)(cft / 1440.0) with an FMA-corrected reciprocal instead of the f64 division subroutine
      const double inv = 1.0 / 1440.0;
      const double q0 = cft * inv;
      const float tf = (float)fma(fma(-q0, 1440.0, cft), inv, q0);
      const float pv = (RES & RES_P) ? Pm[(size_t)prev * D + qi] : pv_r;
      s_q[qi] = fmaf(wt_r, tf, fmaf(wc_r, 1.0f - used, qb_r + pv));
    }

    // (4) feasibility of node tid (+256k): state_agh.py:148-174, exact dtypes; one node per thread, the
    //     32-node words are published through shared memory
#pragma unroll
    for (int k = 0; k < MPT; ++k) {
      const int j = tid + 256 * k;
      if (NPAD <= 128 && warp >= 4) break;      // warp-uniform: warps 4..7 only build the query
      bool mj = true;
      if (j < n && j >= 1) {
        const bool cap = __fadd_rn(demm[k], used) > 1.0f;
        const double fin = __dadd_rn(fmax(__dadd_rn(cft, (double)ttm[k]), (double)twlm[k]), durm[k]);
        mj = vis[k] | cap | (fin > twrm[k]);
        fin_t[j] = fin;      // exactly the new cur_free_time if j is served next (state_agh.py:117-121): reused by the update
      }
      const uint32_t word = __ballot_sync(0xffffffffu, mj);
      if (j < NPAD) {
        if (lane == 0) s_mw[warp + 8 * k] = word;
        dc[j] = ddm[k];
      }
    }
    __syncthreads();   // (C) mask words + distances of this step visible

    uint32_t mw[NPL];
    uint32_t feas = 0u;
#pragma unroll
    for (int p = 0; p < NPL; ++p) {
      mw[p] = s_mw[p];
      feas |= (p == 0) ? (~mw[p] & 0xfffffffeu) : ~mw[p];
    }
    mw[0] = (mw[0] & 0xfffffffeu) | ((prev == 0 && feas != 0u) ? 1u : 0u);      // depot: state_agh.py:173
    int nfeas = 0;
    if constexpr (COMPACT) {
      // Feasibility is sparse (about a quarter of the nodes on average, SURVEY 3.5 workloads): the value and logit
      // phases below walk a COMPACT list of the feasible nodes instead of all n.  Every warp builds the same list
      // (ballot order = increasing node index) and reads it back after its own __syncwarp; masked nodes get their
      // -inf log-prob slot here.
  #pragma unroll
      for (int p = 0; p < NPL; ++p) {
        const uint32_t fw = ~mw[p];
        if ((fw >> lane) & 1u) s_idx[nfeas + __popc(fw & ((1u << lane) - 1u))] = (uint16_t)(32 * p + lane);
        nfeas += __popc(fw);
      }
  #pragma unroll
      for (int k = 0; k < MPT; ++k) {
        const int j = tid + 256 * k;
        if (j < n && ((pick_word<NPL>(mw, warp + 8 * k) >> lane) & 1u)) s_z[j] = -CUDART_INF_F;
      }
      __syncwarp();
    } else {
      // N <= 127: the LOGIT phase walks a compact list of the step's feasible nodes (about 30 % of them on average), which
      // cuts its shared-memory wavefronts -- the busiest unit of this kernel (profiles/r2c_decode_ncu_summary.md) -- to the
      // rows that matter.  Warp p builds the entries of its own 32-node word while the other phases of the step run; the
      // list is complete at barrier (A).  Increasing slot = increasing node index (first-index tie-break).
      int before = 0;
  #pragma unroll
      for (int p = 0; p < NPL; ++p) {
        const int c = __popc(~mw[p]);
        before += (p < warp) ? c : 0;
        nfeas += c;
      }
      if (warp < NPL) {
        const uint32_t fw = ~pick_word<NPL>(mw, warp);
        if ((fw >> lane) & 1u) s_idx[before + __popc(fw & ((1u << lane) - 1u))] = (uint16_t)(32 * warp + lane);
      }
      if constexpr (DBG) {     // the full log-prob output wants -inf at the masked nodes
        if (tid < n && ((pick_word<NPL>(mw, warp) >> lane) & 1u)) s_z[tid] = -CUDART_INF_F;
      }
    }
    if constexpr (DBG) {
      if (a.q_dump != nullptr && tid < D) a.q_dump[(trow + t) * D + tid] = s_q[tid];
      if (warp == 0) {
        if (a.mask_dump != nullptr && lane < nwords) a.mask_dump[(trow + t) * nwords + lane] = pick_word<NPL>(mw, lane);
        if (lane == 0) {
          if (a.used_dump != nullptr) a.used_dump[trow + t] = used;
          if (a.cft_dump != nullptr) a.cft_dump[trow + t] = cft;
        }
      }
    }

    // (3) compatibilities of my nodes with head h against the register-resident K' (which already holds the
    //     1/sqrt(16) and a log2(e), so the softmax below is exp2)
    float s[NPL];
    {
      uint64_t sacc[NPL];
#pragma unroll
      for (int p = 0; p < NPL; ++p) sacc[p] = 0ull;
      const float4* q4 = reinterpret_cast<const float4*>(s_q + h * HD);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 qv = q4[i];
        const uint64_t q01 = pack2(qv.x, qv.y), q23 = pack2(qv.z, qv.w);
#pragma unroll
        for (int p = 0; p < NPL; ++p)
          if (!COMPACT || mw[p] != 0xffffffffu) {       // warp-uniform: a fully masked 32-node word needs no scores
            sacc[p] = fma2(kreg[p][2 * i], q01, sacc[p]);
            sacc[p] = fma2(kreg[p][2 * i + 1], q23, sacc[p]);
          }
      }
#pragma unroll
      for (int p = 0; p < NPL; ++p) s[p] = sum2(sacc[p]);
    }

    // (5) masked softmax over the nodes, warp-local: attention_model.py:559-565
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int p = 0; p < NPL; ++p) {
      s[p] = ((mw[p] >> lane) & 1u) ? -CUDART_INF_F : s[p];
      mx = fmaxf(mx, s[p]);
    }
    mx = fkey_inv(__reduce_max_sync(0xffffffffu, fkey(mx)));   // warp max as ONE integer REDUX on the order-preserving key
    float e[NPL], esum = 0.f;
#pragma unroll
    for (int p = 0; p < NPL; ++p) {
      e[p] = exp2f(s[p] - mx);           // exp2(-inf) = 0 for masked nodes; mx is finite (some node is feasible)
      esum += e[p];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, off);
    const float inv_esum = __frcp_rn(esum);          // off the critical path of the butterfly below
    if constexpr (DBG) {
      if (a.attn_dump != nullptr) {
#pragma unroll
        for (int p = 0; p < NPL; ++p)
          if (lane + 32 * p < n) a.attn_dump[((trow + t) * H + h) * n + lane + 32 * p] = e[p] * inv_esum;
      }
    }

    // (6) o_h = sum_j a_j V_jh, butterfly transpose-reduce over the 32 lanes
    float acc[HD];
    if constexpr (COMPACT) {
#pragma unroll
      for (int k = 0; k < HD; ++k) acc[k] = 0.f;
      for (int c0 = 0; c0 < nfeas; c0 += 32) {          // lane <-> compact slot; masked nodes have weight 0 and are skipped
        const bool valid = c0 + lane < nfeas;
        const int j = valid ? (int)s_idx[c0 + lane] : 0;
        const int src = j & 31, pj = j >> 5;
        float ev = 0.f;                                 // weight of node j: register e[pj] of lane src
  #pragma unroll
        for (int p = 0; p < NPL; ++p) {
          const float v = __shfl_sync(0xffffffffu, e[p], src);
          ev = (pj == p) ? v : ev;
        }
        if (valid) {
          const float4* vr = reinterpret_cast<const float4*>(Vm + (size_t)j * KS);
  #pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = vr[SWV ? swz(h * 4 + i, j) : h * 4 + i];
            acc[4 * i + 0] = fmaf(ev, v.x, acc[4 * i + 0]);
            acc[4 * i + 1] = fmaf(ev, v.y, acc[4 * i + 1]);
            acc[4 * i + 2] = fmaf(ev, v.z, acc[4 * i + 2]);
            acc[4 * i + 3] = fmaf(ev, v.w, acc[4 * i + 3]);
          }
        }
      }
    } else {
      uint64_t acc2[HD / 2];
  #pragma unroll
      for (int k = 0; k < HD / 2; ++k) acc2[k] = 0ull;
  #pragma unroll
      for (int p = 0; p < NPL; ++p) {
        const uint64_t ee = pack2(e[p], e[p]);
  #pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = (p == NPL - 1) ? *vb_last[i] : vb[i][p * 32 * (KS / 4)];
          acc2[2 * i] = fma2(ee, pack2(v.x, v.y), acc2[2 * i]);
          acc2[2 * i + 1] = fma2(ee, pack2(v.z, v.w), acc2[2 * i + 1]);
        }
      }
  #pragma unroll
      for (int k = 0; k < HD / 2; ++k) unpack2(acc2[k], acc[2 * k], acc[2 * k + 1]);
    }
    {
      float r8[8], r4[4], r2[2], r1;
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float send = b4 ? acc[k] : acc[k + 8], keep = b4 ? acc[k + 8] : acc[k];
        r8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float send = b3 ? r8[k] : r8[k + 4], keep = b3 ? r8[k + 4] : r8[k];
        r4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float send = b2 ? r4[k] : r4[k + 2], keep = b2 ? r4[k + 2] : r4[k];
        r2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      {
        const float send = b1 ? r2[0] : r2[1], keep = b1 ? r2[1] : r2[0];
        r1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
      // lane now owns element lane >> 1 of o_h
      if ((lane & 1) == 0) s_o[h * HD + (lane >> 1)] = r1 * inv_esum;
    }
    __syncthreads();   // (A) o complete
    if constexpr (DBG) {
      if (a.o_dump != nullptr && tid < D) a.o_dump[(trow + t) * D + tid] = s_o[tid];
    }

    if constexpr (COMPACT) {
      // (7) logits of the FEASIBLE nodes: 4 nodes per warp and pass, 8 lanes per node.  Lane fg of a node takes the
      //     16-byte chunks {fg, fg+8, fg+16, fg+24} of its LK' row: the 8 lanes of a quarter-warp read one whole
      //     swizzled 128-byte group (conflict-free) and the matching chunks of o are a broadcast across the 4 nodes.
      // selection score as an order-preserving integer key, so the arg-max is two warp REDUX operations
      int best_key = (int)0x80000000, second_key = (int)0x80000000;
      int best_j = 0x7fffffff;
      float wm = shift, ws = 0.f;
      for (int c0 = warp * 4; c0 < nfeas; c0 += 32) {
        const bool valid = c0 + ns < nfeas;
        const int j = valid ? (int)s_idx[c0 + ns] : 0;
        float u = 0.f;
        if (valid) {
          const float4* lr = reinterpret_cast<const float4*>(Lm + (size_t)j * KS);
          const float4* o4 = reinterpret_cast<const float4*>(s_o);
          const int sw = SWL ? ((fg ^ j) & 7) : fg;
          const float u0 = dot4(o4[fg], lr[sw], 0.f), u1 = dot4(o4[8 + fg], lr[8 + sw], 0.f);
          const float u2 = dot4(o4[16 + fg], lr[16 + sw], 0.f), u3 = dot4(o4[24 + fg], lr[24 + sw], 0.f);
          u = (u0 + u1) + (u2 + u3);
        }
        u += __shfl_xor_sync(0xffffffffu, u, 1);
        u += __shfl_xor_sync(0xffffffffu, u, 2);
        u += __shfl_xor_sync(0xffffffffu, u, 4);
        if (valid && fg == 0) {
          float z = 10.0f * tanhf(u);                               // attention_model.py:580
          if (DBG && a.inject_z != nullptr) z = __ldg(a.inject_z + (trow + t) * n + j);
          if (!unit_temp) z = __fdiv_rn(z, a.temp);                 // :447
          s_z[j] = z;
          float sc = z;
          if (DBG && a.inject_logp != nullptr) {
            sc = expf(__ldg(a.inject_logp + (trow + t) * n + j));   // argmax(exp(log_p)), :333,:377
          } else if (sampling) {
            sc = z + gumbel(a.seed, rng_id, (uint32_t)t, (uint32_t)j);
          }
          if (fixed_shift) ws += expf(z - shift);
          const int key = fkey(sc);
          // strict: the list is in index order, first index wins ties; the runner-up feeds the near-tie test
          if (key > best_key) { second_key = best_key; best_key = key; best_j = j; }
          else if (key > second_key) second_key = key;
        }
      }
      if (!fixed_shift) {                                           // cold temperatures: exact max-merging log-sum-exp
        __syncwarp();
        wm = -CUDART_INF_F;
        for (int c0 = warp * 4; c0 < nfeas; c0 += 32)
          if (c0 + ns < nfeas && fg == 0) wm = fmaxf(wm, s_z[s_idx[c0 + ns]]);
  #pragma unroll
        for (int off = 16; off >= 1; off >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, off));
        for (int c0 = warp * 4; c0 < nfeas; c0 += 32)
          if (c0 + ns < nfeas && fg == 0) ws += expf(s_z[s_idx[c0 + ns]] - wm);
      }
  #pragma unroll
      for (int off = 16; off >= 1; off >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, off);
      {
        const int kmax = __reduce_max_sync(0xffffffffu, best_key);
        int jmin = __reduce_min_sync(0xffffffffu, best_key == kmax ? best_j : 0x7fffffff);
        if (tie_check) {      // does this warp hold a second logit within NEAR_TIE of its best one?
          const int kthr = fkey(fkey_inv(kmax) - NEAR_TIE);
          const bool mine = best_key == kmax && best_j == jmin;
          if (__any_sync(0xffffffffu, (!mine && best_key >= kthr) || second_key >= kthr)) jmin |= 0x10000;
        }
        if (lane == 0)
          *reinterpret_cast<float4*>(s_red + warp * 4) = make_float4(ws, __int_as_float(kmax), __int_as_float(jmin), wm);
      }
    } else {
      // (7) logits, fixed mapping: 4 nodes per warp and 32-node pass, 8 lanes per node (see the row pointers above)
      constexpr int NP2 = NPL >= 3 ? 4 : NPL;        // passes padded to a power of two for the transpose-reduce
      uint64_t o01[4], o23[4];
      {
        const float4* o4 = reinterpret_cast<const float4*>(s_o) + fg;
  #pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 ov = o4[8 * k];
          o01[k] = pack2(ov.x, ov.y);
          o23[k] = pack2(ov.z, ov.w);
        }
      }
      float u[4] = {0.f, 0.f, 0.f, 0.f};
      int jn[4] = {0, 0, 0, 0};
  #pragma unroll
      for (int ps = 0; ps < NPL; ++ps) {
        if (ps * 32 < nfeas) {                                  // warp-uniform: passes beyond the list are skipped
          const int c = ps * 32 + warp * 4 + ns;
          jn[ps] = c < nfeas ? (int)s_idx[c] : 0;               // lanes past the end read row 0 (their result is dropped)
          const float4* lr = lbase + jn[ps] * (KS / 4);
          uint64_t ua = 0ull, ub = 0ull;
  #pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 lv = lr[8 * k];
            ua = fma2(o01[k], pack2(lv.x, lv.y), ua);
            ub = fma2(o23[k], pack2(lv.z, lv.w), ub);
          }
          u[ps] = sum2(ua) + sum2(ub);
        }
      }
      // transpose-reduce over the 8 lanes of a node group: lane (bit0, bit1 of fg) ends with the full dot product of
      // the node of pass 2 bit0 + bit1; lanes fg and fg ^ 4 hold the same value
      float uu;
      int pass_own;
      if constexpr (NP2 == 4) {
        const bool b0 = fg & 1, b1 = fg & 2;
        const float k0 = b0 ? u[2] : u[0], k1 = b0 ? u[3] : u[1], s0 = b0 ? u[0] : u[2], s1 = b0 ? u[1] : u[3];
        const float a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 1), a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 1);
        const float kk = b1 ? a1 : a0, ss = b1 ? a0 : a1;
        uu = kk + __shfl_xor_sync(0xffffffffu, ss, 2);
        pass_own = (b0 ? 2 : 0) + (b1 ? 1 : 0);
      } else if constexpr (NP2 == 2) {
        const bool b0 = fg & 1;
        const float kk = b0 ? u[1] : u[0], ss = b0 ? u[0] : u[1];
        uu = kk + __shfl_xor_sync(0xffffffffu, ss, 1);
        uu += __shfl_xor_sync(0xffffffffu, uu, 2);
        pass_own = b0 ? 1 : 0;
      } else {
        uu = u[0] + __shfl_xor_sync(0xffffffffu, u[0], 1);
        uu += __shfl_xor_sync(0xffffffffu, uu, 2);
        pass_own = 0;
      }
      uu += __shfl_xor_sync(0xffffffffu, uu, 4);
      const bool owner = NP2 == 4 ? (fg & 4) == 0 : (NP2 == 2 ? (fg & 6) == 0 : fg == 0);
      const int j = pass_own == 0 ? jn[0] : (pass_own == 1 ? jn[1] : (pass_own == 2 ? jn[2] : jn[3]));
      const bool valid = owner && pass_own * 32 + warp * 4 + ns < nfeas;      // every list entry is a feasible node
      const bool mj = !valid;
      float z = -CUDART_INF_F;
      if (!mj) {
        z = 10.0f * tanhf(uu);                                  // attention_model.py:580
        if (DBG && a.inject_z != nullptr) z = __ldg(a.inject_z + (trow + t) * n + j);
        if (!unit_temp) z = __fdiv_rn(z, a.temp);               // :447
      }
      if (valid) s_z[j] = z;
      float sc = z;
      if (DBG && a.inject_logp != nullptr) {
        sc = valid ? expf(__ldg(a.inject_logp + (trow + t) * n + j)) : -CUDART_INF_F;   // argmax(exp(log_p)), :333,:377
      } else if (sampling && !mj) {
        sc = z + gumbel(a.seed, rng_id, (uint32_t)t, (uint32_t)j);
      }
      // selection score as an order-preserving integer key, so the arg-max is two warp REDUX operations
      const int key = valid ? fkey(sc) : (int)0x80000000;
      float wm = shift, ws;
      if (!fixed_shift) {
        wm = z;
  #pragma unroll
        for (int off = 16; off >= 1; off >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, off));
      }
      ws = mj ? 0.f : expf(z - wm);
  #pragma unroll
      for (int off = 16; off >= 1; off >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, off);
      {
        const int kmax = __reduce_max_sync(0xffffffffu, key);
        int jmin = __reduce_min_sync(0xffffffffu, key == kmax ? j : 0x7fffffff);     // first index wins ties
        if (tie_check) {      // does this warp hold a second logit within NEAR_TIE of its best one?
          const float zthr = fkey_inv(kmax) - NEAR_TIE;
          if (__popc(__ballot_sync(0xffffffffu, !mj && z >= zthr)) > 1) jmin |= 0x10000;
        }
        if (lane == 0)
          *reinterpret_cast<float4*>(s_red + warp * 4) = make_float4(ws, __int_as_float(kmax), __int_as_float(jmin), wm);
      }
    }
    __syncthreads();   // (B) per-warp partials complete

    // (8) merge (lanes 0..7 take one warp's partial each), select, update; every warp does it redundantly so the
    //     state stays in registers: attention_model.py:446-447,372-390; state_agh.py:99-137
    float M = shift, S;
    int sel;
    {
      float4 pr = make_float4(0.f, __int_as_float((int)0x80000000), __int_as_float(0x7fffffff), -CUDART_INF_F);
      if (lane < 8) pr = *reinterpret_cast<const float4*>(s_red + lane * 4);
      const int key = __float_as_int(pr.y);
      const int kmax = __reduce_max_sync(0xffffffffu, key);
      const int jw = __float_as_int(pr.z) & 0xffff;
      sel = __reduce_min_sync(0xffffffffu, key == kmax ? jw : 0x7fffffff);
      if (tie_check) {
        // another warp's best logit, or a second one of the winning warp, within NEAR_TIE of the maximum
        const float zmax = fkey_inv(kmax);
        const int kthr = fkey(zmax - NEAR_TIE);
        if (__any_sync(0xffffffffu, lane < 8 && key >= kthr && (jw != sel || (__float_as_int(pr.z) & 0x10000)))) {
          // re-decide with the reference's arithmetic: log_p = (z - max) - log(sum exp(z - max)), p = exp(log_p),
          // first index of the maximum (attention_model.py:446-447,333,377).  Every warp computes the same values.
          float ssum = 0.f;
          for (int c = lane; c < nfeas; c += 32) ssum += expf(s_z[s_idx[c]] - zmax);     // the list holds the feasible nodes
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, off);
          const float lse = logf(ssum);
          int pk = (int)0x80000000, pj = 0x7fffffff;
          for (int c = lane; c < nfeas; c += 32) {                    // increasing slot = increasing node index
            const int j = s_idx[c];
            const float z = s_z[j];
            if (z >= zmax - NEAR_TIE) {
              const int k2 = __float_as_int(expf((z - zmax) - lse));      // p > 0: its bits order like the value
              if (k2 > pk) { pk = k2; pj = j; }
            }
          }
          const int pmax = __reduce_max_sync(0xffffffffu, pk);
          sel = __reduce_min_sync(0xffffffffu, pk == pmax ? pj : 0x7fffffff);
        }
      }
      float part = pr.x;
      if (!fixed_shift) {
        float wmax = pr.w;
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
        M = __shfl_sync(0xffffffffu, wmax, 0);
        part = (pr.w == -CUDART_INF_F) ? 0.f : pr.x * expf(pr.w - M);
      }
#pragma unroll
      for (int off = 4; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      S = __shfl_sync(0xffffffffu, part, 0);
    }
    if (replay) sel = (int)a.actions[trow + t];
    {   // next step's distance / travel-time row and P_sc component (see (1))
      const int cnext = s_crd[sel];
#pragma unroll
      for (int k = 0; k < MPT; ++k) {
        const bool ok = tid + 256 * k < n;
        ddm[k] = ok ? __ldg(a.distance + NODE_SIZE * cnext + cjm[k]) : 0.f;
        ttm[k] = ok ? __ldg(a.travel + NODE_SIZE * cnext + cjm[k]) : 0.f;
      }
      if (!(RES & RES_P) && qi >= 0) pv_r = __ldg(Pm + (size_t)sel * D + qi);
    }
    // log-likelihood of the step (attention_model.py:238-250): only thread 0 reports it, so only warp 0 pays for the log
    float lp = 0.f;
    if (DBG || warp == 0) {
      const float logS = logf(S);
      lp = (s_z[sel] - M) - logS;
      ll += lp;
      if constexpr (DBG) {
        if (a.logp_full != nullptr) {
          for (int j = tid; j < n; j += 256) a.logp_full[(trow + t) * n + j] = (s_z[j] - M) - logS;
        }
        if (a.lse_dump != nullptr && tid == 0) a.lse_dump[trow + t] = M + logS;
      }
    }

    const float d = dc[sel];
    len = __fadd_rn(len, d);
    if (sel != 0) {
      used = __fadd_rn(used, s_dem[sel]);
      cft = fin_t[sel];      // = fmax(cft + d/110, tw_left[sel]) + duration[sel], evaluated by the node's own thread above
      ++nvis;
    } else {
      used = 0.f;
      cft = -60.0;
      nvis += depot_seen ? 0 : 1;
      depot_seen = true;
    }
#pragma unroll
    for (int k = 0; k < MPT; ++k) vis[k] |= (sel == tid + 256 * k);
    if (tid == 0) {
      s_serve[sel] = (float)cft;
      a.pi[trow + t] = (int16_t)sel;
      if (DBG && a.logp_sel != nullptr) a.logp_sel[trow + t] = lp;
    }
    prev = sel;
    ++t;
  }

  // ---- epilogue: closed-tour length (problem_agh.py:49-50,66), outputs, zero padding ----
  __syncthreads();
  if (tid == 0) {
    const float back = __ldg(a.distance + NODE_SIZE * s_crd[prev]);
    a.cost[b] = __fadd_rn(len, back);
    a.ll[b] = ll;
    a.steps[b] = t;
    if (a.status != nullptr) a.status[b] = (!replay && nvis < n) ? AGH_DECODE_STUCK : 0;
  }
  for (int j = tid; j < n; j += 256) a.serve[(size_t)b * n + j] = s_serve[j];
  for (int k = t + tid; k < a.t_stride; k += 256) {
    a.pi[trow + k] = 0;
    if (DBG && a.logp_sel != nullptr) a.logp_sel[trow + k] = 0.f;
  }
}

template <int NPL, int RES, bool DBG>
static int launch_decode(const agh_decode_args& a, cudaStream_t st) {
  const int n = a.N + 1;
  const DecodeSmem L = decode_smem_layout(n, NPL, RES);
  static_assert(SmallOff<NPL>::mats % 128 == 0, "matrix region is 128-byte aligned");
  if (L.V != SmallOff<NPL>::mats || L.twl != SmallOff<NPL>::twl || L.fin != SmallOff<NPL>::fin) {
    set_error("agh_decode: shared-memory layout mismatch");
    return AGH_E_UNSUPPORTED;
  }
  auto kern = decode_kernel<NPL, RES, DBG>;
  AGH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  kern<<<a.B, 256, L.total, st>>>(a);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

template <int NPL, int RES>
static int launch_variant(const agh_decode_args& a, cudaStream_t st) {
  const bool dbg = a.mode == AGH_REPLAY || a.inject_logp || a.inject_z || a.logp_sel || a.logp_full || a.mask_dump ||
                   a.used_dump || a.cft_dump || a.q_dump || a.attn_dump || a.o_dump || a.lse_dump;
  return dbg ? launch_decode<NPL, RES, true>(a, st) : launch_decode<NPL, RES, false>(a, st);
}

// Pick the residency: the candidate with the most resident CTAs per SM wins (two instances per SM hide
// each other's step latency), ties go to the candidate that keeps more in shared memory.
template <int NPL>
static int dispatch_res(const agh_decode_args& a, cudaStream_t st) {
  const int n = a.N + 1;
  int dev = 0, max_smem = 0, sm_smem = 0;
  AGH_CUDA(cudaGetDevice(&dev));
  AGH_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  AGH_CUDA(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  const int cand[4] = {RES_V | RES_L | RES_P, RES_V | RES_L, RES_L, 0};
  int best = -1, best_ctas = 0;
  for (int c = 0; c < 4; ++c) {
    if (a.smem_mats > 0 && cand[c] != (a.smem_mats & 7)) continue;          // explicit override (bitmask)
    const int bytes = (int)decode_smem_layout(n, NPL, cand[c]).total;
    if (bytes > max_smem) continue;
    int ctas = sm_smem / (bytes + 1024);
    if (ctas > decode_min_blocks(NPL)) ctas = decode_min_blocks(NPL);
    if (ctas > best_ctas) { best_ctas = ctas; best = cand[c]; }
  }
  switch (best) {
    case RES_V | RES_L | RES_P: return launch_variant<NPL, RES_V | RES_L | RES_P>(a, st);
    case RES_V | RES_L: return launch_variant<NPL, RES_V | RES_L>(a, st);
    case RES_L: return launch_variant<NPL, RES_L>(a, st);
    case 0: return launch_variant<NPL, 0>(a, st);
    default: break;
  }
  set_error("agh_decode: N=%d does not fit shared memory (%d bytes available)", a.N, max_smem);
  return AGH_E_UNSUPPORTED;
}

int try_decode_shared(const agh_decode_args& a, cudaStream_t st, bool* taken);      // decode_shared.cu

}  // namespace agh

extern "C" int agh_decode(const agh_decode_args* args, void* stream) {
  using namespace agh;
  AGH_CHECK_ARG(args != nullptr);
  const agh_decode_args& a = *args;
  AGH_CHECK_ARG(a.B >= 0 && a.N >= 1 && a.N <= AGH_MAX_N - 1);
  if (a.B == 0) return AGH_OK;
  AGH_CHECK_ARG(a.keys && a.psc && a.qbase && a.coord && a.demand && a.tw_left && a.tw_right && a.duration);
  AGH_CHECK_ARG(a.distance && a.travel && a.wblob && a.pi && a.steps && a.ll && a.cost && a.serve);
  AGH_CHECK_ARG(a.mode == AGH_GREEDY || a.mode == AGH_SAMPLING || a.mode == AGH_REPLAY);
  AGH_CHECK_ARG(a.mode != AGH_REPLAY || (a.actions != nullptr && a.replay_steps >= 0 && a.replay_steps <= a.t_stride));
  AGH_CHECK_ARG(a.t_stride >= 2 * a.N - 1 && a.t_stride >= 2);      // N = 1 takes two steps: the node, then the depot
  AGH_CHECK_ARG(a.temp > 0.f);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {   // replicas that share one staged key set (sample_many): one warp per replica, decode_shared.cu
    bool taken = false;
    const int rc = try_decode_shared(a, st, &taken);
    if (rc != AGH_OK || taken) return rc;
  }
  const int npl = ceil_div(a.N + 1, 32);
  switch (npl) {
    case 1: return dispatch_res<1>(a, st);
    case 2: return dispatch_res<2>(a, st);
    case 3: return dispatch_res<3>(a, st);
    case 4: return dispatch_res<4>(a, st);
    case 5: return dispatch_res<5>(a, st);
    case 6: return dispatch_res<6>(a, st);
    case 7: return dispatch_res<7>(a, st);
    case 8: return dispatch_res<8>(a, st);
    case 9: return dispatch_res<9>(a, st);
    case 10: return dispatch_res<10>(a, st);
    default: break;
  }
  set_error("agh_decode: N=%d unsupported", a.N);
  return AGH_E_UNSUPPORTED;
}
