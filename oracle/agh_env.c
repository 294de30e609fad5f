/* Plain-C restatement of the AGH environment -- TEST INFRASTRUCTURE ONLY (see oracle/agh_oracle.py).
 *
 * Scalar, one instance at a time, no vectorisation: an independent second statement of the integer /
 * mixed f32-f64 contract the CUDA kernels must reproduce bit-for-bit:
 *   mask    CB/problems/agh/state_agh.py:148-174
 *   update  CB/problems/agh/state_agh.py:99-137
 *   cost    CB/problems/agh/problem_agh.py:18-66   (returns a status instead of asserting)
 *   nearest-neighbour driver  CB/agh_baseline.py:80-101  (weight-free KAT: 147099.953125 on agh20)
 * Parity: pinned by tests/test_oracle_golden.py::test_c_env_matches_golden against fixtures from the
 * unmodified reference.  Build: make -C oracle  (gcc -O2 -ffp-contract=off, no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NODE_SIZE 92
#define SPEED 110.0f

typedef struct {
  int N;
  const int64_t* loc;      /* [N] 1..91 */
  const float* demand;     /* [N] */
  const float* tw_left;    /* [N+1] */
  const double* tw_right;  /* [N+1] */
  const double* duration;  /* [N] */
  const float* dist;       /* [92*92] */
  /* state */
  int prev;
  float used, lengths;
  double cft;
  uint8_t visited[512];
  float serve[512];
  int steps;
} env_t;

static int coord(const env_t* e, int j) { return j == 0 ? 0 : (int)e->loc[j - 1]; }
static double dur(const env_t* e, int j) { return j == 0 ? 0.0 : e->duration[j - 1]; }

static void env_init(env_t* e) {
  e->prev = 0; e->used = 0.f; e->lengths = 0.f; e->cft = -60.0; e->steps = 0;
  memset(e->visited, 0, sizeof(e->visited));
  memset(e->serve, 0, sizeof(e->serve));
}

/* mask[j] = 1 infeasible */
static void env_mask(const env_t* e, uint8_t* mask) {
  int any = 0;
  const int cp = coord(e, e->prev);
  for (int j = 1; j <= e->N; ++j) {
    const int cap = (e->demand[j - 1] + e->used) > 1.0f;
    const float tt = e->dist[NODE_SIZE * cp + coord(e, j)] / SPEED;           /* f32 division */
    const double a = e->cft + (double)tt;
    const double b = (double)e->tw_left[j];
    const double fin = (a > b ? a : b) + dur(e, j);
    const int tw = fin > e->tw_right[j];
    mask[j] = (uint8_t)(e->visited[j] | cap | tw);
    any |= !mask[j];
  }
  mask[0] = (uint8_t)(e->prev == 0 && any);
}

static void env_update(env_t* e, int sel) {
  const float d = e->dist[NODE_SIZE * coord(e, e->prev) + coord(e, sel)];
  e->lengths = e->lengths + d;
  if (sel != 0) {
    e->used = e->used + e->demand[sel - 1];
    const float tt = d / SPEED;
    const double a = e->cft + (double)tt, b = (double)e->tw_left[sel];
    e->cft = (a > b ? a : b) + dur(e, sel);
  } else {
    e->used = 0.f;
    e->cft = -60.0;
  }
  e->visited[sel] = 1;
  e->serve[sel] = (float)e->cft;
  e->prev = sel;
  e->steps++;
}

static int env_finished(const env_t* e) {
  if (e->steps < e->N) return 0;
  for (int j = 0; j <= e->N; ++j) if (!e->visited[j]) return 0;
  return 1;
}

/* status bits: 1 invalid tour, 2 capacity, 4 time window */
int agh_c_get_cost(int N, int T, const int64_t* pi, const int64_t* loc, const float* demand, const float* tw_left,
                   const double* tw_right, const double* duration, const float* dist, float* cost) {
  int cnt[512] = {0};
  int st = 0;
  for (int t = 0; t < T; ++t) { if (pi[t] < 0 || pi[t] > N) return 1; cnt[pi[t]]++; }
  for (int j = 1; j <= N; ++j) if (cnt[j] != 1) return 1;
  float used = 0.f, total = 0.f;
  double cur = -60.0;
  int pc = 0;
  for (int t = 0; t < T; ++t) {
    const int a = (int)pi[t];
    used = used + (a == 0 ? -1.0f : demand[a - 1]);
    if (used < 0.f) used = 0.f;
    if (!(used <= 1.0f + 1e-5f)) st |= 2;
    const int c = a == 0 ? 0 : (int)loc[a - 1];
    const float d = dist[NODE_SIZE * pc + c];
    total = total + d;
    if (a != 0) {
      const double x = cur + (double)(d / SPEED), y = (double)tw_left[a];
      cur = (x > y ? x : y) + duration[a - 1];
    } else {
      cur = -60.0;
    }
    if (!(cur <= tw_right[a] + 1e-5)) st |= 4;
    pc = c;
  }
  *cost = total + dist[NODE_SIZE * pc];
  return st;
}

/* nearest-neighbour heuristic over a batch; pi [B][2N] (zero padded), steps [B], serve [B][N+1], cost [B].
 * Also dumps the mask of every step when mask_out != NULL ([B][2N][N+1]). */
void agh_c_nearest_neighbor(int B, int N, const int64_t* loc, const float* demand, const float* tw_left,
                            const double* tw_right, const double* duration, const float* dist, int64_t* pi, int32_t* steps,
                            float* serve, float* cost, uint8_t* mask_out) {
  for (int b = 0; b < B; ++b) {
    env_t e;
    e.N = N; e.loc = loc + (size_t)b * N; e.demand = demand + (size_t)b * N; e.tw_left = tw_left + (size_t)b * (N + 1);
    e.tw_right = tw_right + (size_t)b * (N + 1); e.duration = duration + (size_t)b * N; e.dist = dist;
    env_init(&e);
    int64_t* p = pi + (size_t)b * 2 * N;
    memset(p, 0, sizeof(int64_t) * 2 * N);
    uint8_t mask[512];
    while (!env_finished(&e)) {
      env_mask(&e, mask);
      if (mask_out) memcpy(mask_out + ((size_t)b * 2 * N + e.steps) * (N + 1), mask, N + 1);
      int best = 0; float bd = INFINITY;
      const int cp = coord(&e, e.prev);
      for (int j = 0; j <= N; ++j) {
        const float d = mask[j] ? 10000.f : dist[NODE_SIZE * cp + coord(&e, j)];
        if (d < bd) { bd = d; best = j; }                 /* first minimum, like torch.min */
      }
      p[e.steps] = best;
      env_update(&e, best);
    }
    steps[b] = e.steps;
    memcpy(serve + (size_t)b * (N + 1), e.serve, sizeof(float) * (N + 1));
    agh_c_get_cost(N, e.steps, p, e.loc, e.demand, e.tw_left, e.tw_right, e.duration, dist, cost + b);
  }
}
