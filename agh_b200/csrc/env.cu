// AGH environment pieces outside the fused decode kernel: fleet-batch assembly (a1), cross-fleet
// time-window chaining, tour validation + cost (a13), single-step mask/update for the StateAGH API
// (a8/a12) and the best-of-R reduction of sample_many (a16).  Arithmetic follows the reference's
// dtypes exactly: f32 distance lookup and f32 IEEE division by SPEED, f64 time
// accumulation, f32 capacity.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace agh {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- a1: train.py:40-57 ----
__global__ void fleet_task_kernel(const int64_t* __restrict__ loc, const int64_t* __restrict__ type,
                                  const float* __restrict__ departure, const float* __restrict__ demand_all,
                                  const float* __restrict__ twl_level, int fleet, double d0, double d1, double d2,
                                  double nd0, double nd1, double nd2, int nd_is_f32, int B, int N,
                                  uint8_t* __restrict__ coord, float* __restrict__ demand, float* __restrict__ tw_left,
                                  double* __restrict__ tw_right, double* __restrict__ duration) {
  const int n = N + 1;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * n) return;
  const int j = (int)(idx % n);
  const size_t b = idx / n;
  if (j == 0) {
    coord[idx] = 0; tw_left[idx] = 0.f; tw_right[idx] = 1441.0; duration[idx] = 0.0;
    return;
  }
  const size_t src = b * N + (j - 1);
  const int ty = (int)type[src];
  coord[idx] = (uint8_t)loc[src];
  demand[src] = demand_all[(b * 10 + (fleet - 1)) * N + (j - 1)];
  tw_left[idx] = twl_level[src];
  const double nd = ty == 0 ? nd0 : (ty == 1 ? nd1 : nd2);
  // departure (f32) - next_duration: f64 arithmetic, or f32 when the level's list holds python floats
  tw_right[idx] = nd_is_f32 ? (double)__fsub_rn(departure[src], (float)nd) : __dsub_rn((double)departure[src], nd);
  duration[idx] = ty == 0 ? d0 : (ty == 1 ? d1 : d2);
}

// train.py:67
__global__ void chain_kernel(float* __restrict__ lvl, const float* __restrict__ serve, int B, int N) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * N) return;
  const size_t b = idx / N;
  const int j = (int)(idx % N);
  lvl[idx] = fmaxf(lvl[idx], serve[b * (N + 1) + j + 1]);
}

// ---- instance generator on the device (problem_agh.py:93-116 / generate_data.py:8-18 semantics) ----
// One thread per (instance, flight); Philox4x32-10 keyed by the seed, counter = (instance, flight, draw block).  Same
// distributions as the reference's numpy draws -- loc ~ U{1..91}, arrival = 60 Cat(arrival_prob) + U{0..59},
// type ~ U{0,1,2}, departure = arrival + (30,34,33)[type], demand ~ U{1..9} / capacity -- but NOT the same stream:
// the numpy generator (problems/agh/problem_agh.py) stays the parity generator.
__device__ __forceinline__ void philox4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ uint32_t bounded(uint32_t r, uint32_t n) { return __umulhi(r, n); }      // floor(r n / 2^32) in [0, n)

__global__ void generate_kernel(uint64_t seed, long long id_offset, int B, int N, int fleets, float capacity,
                                const float* __restrict__ arrival_cdf /*[24], cdf[23] = 1*/, int64_t* __restrict__ loc,
                                float* __restrict__ arrival, float* __restrict__ departure, int64_t* __restrict__ type,
                                float* __restrict__ demand /*[B][fleets][N]*/) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * N) return;
  const size_t b = idx / N;
  const int j = (int)(idx % N);
  const uint64_t gid = (uint64_t)(id_offset + (long long)b);
  uint32_t r[4];
  philox4(seed, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)j, 0u, r);
  loc[idx] = 1 + (int64_t)bounded(r[0], 91u);
  const float u = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  int hour = 0;
  while (hour < 23 && u > arrival_cdf[hour]) ++hour;
  const float arr = (float)(60 * hour + (int)bounded(r[2], 60u));
  const int ty = (int)bounded(r[3], 3u);
  arrival[idx] = arr;
  type[idx] = ty;
  departure[idx] = arr + (ty == 0 ? 30.0f : (ty == 1 ? 34.0f : 33.0f));
  for (int f0 = 0; f0 < fleets; f0 += 4) {
    philox4(seed, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)j, 1u + (uint32_t)(f0 >> 2), r);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (f0 + k < fleets) demand[(b * fleets + f0 + k) * N + j] = __fdiv_rn((float)(1 + bounded(r[k], 9u)), capacity);
  }
}

// ---- a13: problem_agh.py:18-66, one warp per instance ----
// Validity first (every flight exactly once: a per-warp visit counter in shared memory), then the replay of the tour: the
// per-step operands (demand, node location, travel distance, window, duration) are fetched by the lanes in parallel, 32 steps
// at a time, and the sequential recurrences (capacity, length, clock) are evaluated by every lane on shuffled copies -- the
// same operations in the same order as the reference's loop, without a chain of dependent loads per step.
constexpr int GC_WARPS = 4;
template <typename PiT>
__global__ void __launch_bounds__(GC_WARPS * 32) get_costs_kernel(const PiT* __restrict__ pi, int T, const uint8_t* __restrict__ coord,
                                 const float* __restrict__ demand, const float* __restrict__ tw_left,
                                 const double* __restrict__ tw_right, const double* __restrict__ duration,
                                 const float* __restrict__ dist, int B, int N, float* __restrict__ cost,
                                 int32_t* __restrict__ status) {
  __shared__ unsigned int s_cnt[GC_WARPS][AGH_MAX_N + 1];
  const int n = N + 1;
  const int b = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (b >= B) return;
  const PiT* p = pi + (size_t)b * T;
  unsigned int* cnt = s_cnt[wib];
  int st = 0;
  // (1) tour is zeros plus a permutation of 1..N (problem_agh.py:21-27): count visits per node
  int bad = 0;
  for (int j = lane; j <= N; j += 32) cnt[j] = 0u;
  __syncwarp();
  for (int t = lane; t < T; t += 32) {
    const int a = (int)p[t];
    if (a < 0 || a > N) bad = 1;
    else if (a > 0) atomicAdd(&cnt[a], 1u);
  }
  __syncwarp();
  for (int j = 1 + lane; j <= N; j += 32) bad |= (cnt[j] != 1u);
  if (__any_sync(0xffffffffu, bad)) st |= AGH_TOUR_INVALID;
  if (st) {                         // indices are not safe to replay
    if (lane == 0) { status[b] = st; cost[b] = 0.f; }
    return;
  }
  // (2)+(3) replay
  float used = 0.f, total = 0.f;
  double cur = -60.0;
  int carry_c = 0;                  // location of the node before the chunk (the depot's location index is 0 at t = 0)
  for (int base = 0; base < T; base += 32) {
    const int t = base + lane;
    const bool in = t < T;
    const int a = in ? (int)p[t] : 0;
    const size_t k = (size_t)b * n + a;
    const float dem = a == 0 ? -1.0f : demand[(size_t)b * N + a - 1];
    const int c = in ? (int)coord[k] : 0;
    const float twl = tw_left[k];
    const double dur = duration[k], twr = tw_right[k];
    int pc = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) pc = carry_c;
    const float d = dist[NODE_SIZE * pc + c];
    const float tt = __fdiv_rn(d, 110.0f);
    const int cnt_steps = min(32, T - base);
    for (int i = 0; i < cnt_steps; ++i) {
      const int ai = __shfl_sync(0xffffffffu, a, i);
      const float demi = __shfl_sync(0xffffffffu, dem, i), di = __shfl_sync(0xffffffffu, d, i);
      used = __fadd_rn(used, demi);
      if (used < 0.f) used = 0.f;
      if (!(used <= 1.0f + 1e-5f)) st |= AGH_CAPACITY_EXCEEDED;
      total = __fadd_rn(total, di);
      const double twri = __shfl_sync(0xffffffffu, twr, i);
      if (ai != 0) {
        // first step: cur_time is int64 -60 (problem_agh.py:58) so the sum is f32; identical result
        const float tti = __shfl_sync(0xffffffffu, tt, i), twli = __shfl_sync(0xffffffffu, twl, i);
        const double duri = __shfl_sync(0xffffffffu, dur, i);
        cur = __dadd_rn(fmax(__dadd_rn(cur, (double)tti), (double)twli), duri);
      } else {
        cur = -60.0;
      }
      if (!(cur <= twri + 1e-5)) st |= AGH_TW_VIOLATED;
    }
    carry_c = __shfl_sync(0xffffffffu, c, cnt_steps - 1);
  }
  if (lane == 0) {
    total = __fadd_rn(total, dist[NODE_SIZE * carry_c]);
    cost[b] = total;
    status[b] = st;
  }
}

// ---- a8: state_agh.py:148-174, one thread per (instance, node) + one block-free depot pass ----
__global__ void state_mask_kernel(const agh_state st, uint8_t* __restrict__ mask) {
  const int n = st.N + 1;
  const int b = blockIdx.x;
  __shared__ int any_feasible;
  if (threadIdx.x == 0) any_feasible = 0;
  __syncthreads();
  const int prev = (int)st.prev_a[b];
  const int cprev = st.coord[(size_t)b * n + prev];
  const float used = st.used_capacity[b];
  const double cft = st.cur_free_time[b];
  int feas = 0;
  for (int j = 1 + threadIdx.x; j < n; j += blockDim.x) {
    const size_t k = (size_t)b * n + j;
    const bool cap = __fadd_rn(st.demand[(size_t)b * st.N + j - 1], used) > 1.0f;
    const float tt = __fdiv_rn(st.distance[NODE_SIZE * cprev + st.coord[k]], 110.0f);
    const double fin = __dadd_rn(fmax(__dadd_rn(cft, (double)tt), (double)st.tw_left[k]), st.duration[k]);
    const bool m = st.visited[k] | cap | (fin > st.tw_right[k]);
    mask[k] = m;
    feas |= !m;
  }
  if (feas) atomicOr(&any_feasible, 1);
  __syncthreads();
  if (threadIdx.x == 0) mask[(size_t)b * n] = (prev == 0) && any_feasible;
}

// ---- a12: state_agh.py:99-137 ----
__global__ void state_update_kernel(const agh_state st, const int64_t* __restrict__ selected) {
  const int n = st.N + 1;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= st.B) return;
  const int sel = (int)selected[b];
  const int prev = (int)st.prev_a[b];
  const float d = st.distance[NODE_SIZE * st.coord[(size_t)b * n + prev] + st.coord[(size_t)b * n + sel]];
  st.lengths[b] = __fadd_rn(st.lengths[b], d);
  double cft;
  if (sel != 0) {
    st.used_capacity[b] = __fadd_rn(st.used_capacity[b], st.demand[(size_t)b * st.N + sel - 1]);
    const float tt = __fdiv_rn(d, 110.0f);
    cft = __dadd_rn(fmax(__dadd_rn(st.cur_free_time[b], (double)tt), (double)st.tw_left[(size_t)b * n + sel]),
                    st.duration[(size_t)b * n + sel]);
  } else {
    st.used_capacity[b] = 0.f;
    cft = -60.0;
  }
  st.cur_free_time[b] = cft;
  st.visited[(size_t)b * n + sel] = 1;
  st.serve_time[(size_t)b * n + sel] = (float)cft;
  st.prev_a[b] = sel;
}

// ---- a16: utils/functions.py:216-219, first minimum over replicas ----
__global__ void best_of_kernel(const float* __restrict__ costs, const int16_t* __restrict__ pi, const float* __restrict__ serve,
                               int R, int B, int n, int t_stride, float* __restrict__ best_cost,
                               int16_t* __restrict__ best_pi, float* __restrict__ best_serve, int32_t* __restrict__ best_rep) {
  const int b = blockIdx.x;
  __shared__ float s_c[32];
  __shared__ int s_r[32];
  float bc = INFINITY;
  int br = 0x7fffffff;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float c = costs[(size_t)r * B + b];
    if (c < bc) { bc = c; br = r; }
  }
  for (int off = 16; off >= 1; off >>= 1) {
    const float oc = __shfl_xor_sync(0xffffffffu, bc, off);
    const int orr = __shfl_xor_sync(0xffffffffu, br, off);
    if (oc < bc || (oc == bc && orr < br)) { bc = oc; br = orr; }
  }
  if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = bc; s_r[threadIdx.x >> 5] = br; }
  __syncthreads();
  bc = s_c[0]; br = s_r[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (s_c[w] < bc || (s_c[w] == bc && s_r[w] < br)) { bc = s_c[w]; br = s_r[w]; }
  if (br == 0x7fffffff) br = 0;
  const size_t src = (size_t)br * B + b;
  for (int k = threadIdx.x; k < t_stride; k += blockDim.x) best_pi[(size_t)b * t_stride + k] = pi[src * t_stride + k];
  for (int k = threadIdx.x; k < n; k += blockDim.x) best_serve[(size_t)b * n + k] = serve[src * n + k];
  if (threadIdx.x == 0) { best_cost[b] = bc; best_rep[b] = br; }
}

}  // namespace agh

using namespace agh;

extern "C" const char* agh_last_error(void) { return g_err; }
extern "C" int agh_version(void) { return 100; }
extern "C" unsigned long long agh_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" size_t agh_wblob_floats(void) { return W_TOTAL; }
extern "C" size_t agh_sizeof_decode_args(void) { return sizeof(agh_decode_args); }
extern "C" size_t agh_sizeof_state(void) { return sizeof(agh_state); }
extern "C" size_t agh_sizeof_gemm_desc(void) { return sizeof(agh_gemm_desc); }

extern "C" int agh_fleet_task(const int64_t* loc, const int64_t* type, const float* departure, const float* demand_all,
                              const float* tw_left_levels, int level, int fleet, const double* duration3,
                              const double* next_duration3, int nd_is_f32, int B, int N, uint8_t* coord, float* demand,
                              float* tw_left, double* tw_right, double* duration, void* stream) {
  AGH_CHECK_ARG(loc && type && departure && demand_all && tw_left_levels && duration3 && next_duration3);
  AGH_CHECK_ARG(coord && demand && tw_left && tw_right && duration);
  AGH_CHECK_ARG(level >= 0 && level <= 4 && fleet >= 1 && fleet <= 10 && B >= 0 && N >= 1);
  if (B == 0) return AGH_OK;
  const size_t total = (size_t)B * (N + 1);
  fleet_task_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      loc, type, departure, demand_all, tw_left_levels + (size_t)level * B * N, fleet, duration3[0], duration3[1],
      duration3[2], next_duration3[0], next_duration3[1], next_duration3[2], nd_is_f32, B, N, coord, demand, tw_left,
      tw_right, duration);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_chain_tw_left(float* tw_left_levels, int level, const float* serve, int B, int N, void* stream) {
  AGH_CHECK_ARG(tw_left_levels && serve && level >= 0 && level <= 4 && B >= 0 && N >= 1);
  if (B == 0) return AGH_OK;
  const size_t total = (size_t)B * N;
  chain_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tw_left_levels + (size_t)(level + 1) * B * N,
                                                                                 serve, B, N);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_get_costs(const void* pi, int pi_elem_bytes, int T, const uint8_t* coord, const float* demand, const float* tw_left,
                             const double* tw_right, const double* duration, const float* distance, int B, int N,
                             float* cost, int32_t* status, void* stream) {
  AGH_CHECK_ARG(pi && coord && demand && tw_left && tw_right && duration && distance && cost && status);
  AGH_CHECK_ARG(B >= 0 && N >= 1 && T >= 1 && (pi_elem_bytes == 8 || pi_elem_bytes == 2));
  if (B == 0) return AGH_OK;
  const size_t threads = (size_t)B * 32;
  const unsigned grid = (unsigned)((threads + 127) / 128);      // GC_WARPS warps per block
  if (pi_elem_bytes == 8)
    get_costs_kernel<int64_t><<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int64_t*>(pi), T, coord, demand, tw_left,
                                                                      tw_right, duration, distance, B, N, cost, status);
  else
    get_costs_kernel<int16_t><<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int16_t*>(pi), T, coord, demand, tw_left,
                                                                      tw_right, duration, distance, B, N, cost, status);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_state_get_mask(const agh_state* st, uint8_t* mask, void* stream) {
  AGH_CHECK_ARG(st && mask && st->coord && st->demand && st->tw_left && st->tw_right && st->duration && st->distance);
  AGH_CHECK_ARG(st->prev_a && st->used_capacity && st->visited && st->cur_free_time && st->B >= 0 && st->N >= 1);
  if (st->B == 0) return AGH_OK;
  state_mask_kernel<<<st->B, 128, 0, (cudaStream_t)stream>>>(*st, mask);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_state_update(const agh_state* st, const int64_t* selected, void* stream) {
  AGH_CHECK_ARG(st && selected && st->coord && st->demand && st->tw_left && st->duration && st->distance);
  AGH_CHECK_ARG(st->prev_a && st->used_capacity && st->visited && st->lengths && st->cur_free_time && st->serve_time);
  if (st->B == 0) return AGH_OK;
  state_update_kernel<<<(st->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*st, selected);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_best_of(const float* costs, const int16_t* pi, const float* serve, int R, int B, int N, int t_stride,
                           float* best_cost, int16_t* best_pi, float* best_serve, int32_t* best_rep, void* stream) {
  AGH_CHECK_ARG(costs && pi && serve && best_cost && best_pi && best_serve && best_rep && R >= 1 && B >= 0 && N >= 1);
  if (B == 0) return AGH_OK;
  best_of_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(costs, pi, serve, R, B, N + 1, t_stride, best_cost, best_pi, best_serve,
                                                      best_rep);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}

extern "C" int agh_generate(uint64_t seed, long long id_offset, int B, int N, int fleets, float capacity, const float* arrival_cdf,
                            int64_t* loc, float* arrival, float* departure, int64_t* type, float* demand, void* stream) {
  AGH_CHECK_ARG(arrival_cdf && loc && arrival && departure && type && demand && B >= 0 && N >= 1 && fleets >= 1 && capacity > 0.f);
  if (B == 0) return AGH_OK;
  const size_t total = (size_t)B * N;
  generate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, id_offset, B, N, fleets, capacity,
                                                                                   arrival_cdf, loc, arrival, departure, type, demand);
  AGH_LAUNCH_CHECK();
  return AGH_OK;
}
