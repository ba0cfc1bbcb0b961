// sy_env.cu -- B200 (sm_100a) kernels + C ABI of the batched Scotland Yard environment.
//
// Reference behaviour restated here (paths relative to /root/reference):
//   src/environment/yard.py:80-142   reset          -> sy_reset_kernel
//   src/environment/yard.py:144-269  step           -> sy_step_kernel (phases 1a-1c)
//   src/environment/yard.py:271-335  observations   -> assemble_tile (phase 2)
//   src/environment/yard.py:420-472  possible moves -> dense weight table + move-count table
//   src/environment/action_mask.py:30-83            -> mask scatter in assemble_tile, sy_mask_dense_kernel
//   src/environment/reward_calculator.py:26-266     -> phases 1a (status) and 1b (rewards)
//   src/environment/pathfinding.py:34-137           -> sy_apsp_kernel (all-pairs table, built once per graph)
//   src/environment/belief_module.py:69-111         -> belief propagation in assemble_tile (expectation)
//
// Kernel shape: one CTA of 256 threads owns a tile of 32 consecutive envs.
//   phase 0  cooperative, coalesced staging of the tile's AoS state ([B,A] int32/int64) into smem
//   phase 1a warp 0, lane = env: the order-dependent MrX -> Police0..P-1 move rule, capture/timeout/no-money
//   phase 1b warp a, lane = env: agent a's reward (fp64 or no-FMA fp32) + visit-count RMW  (no divergence)
//   phase 1c warp 0, lane = env: timestep, reveal schedule, same-step auto-reset (Philox), statistics
//   phase 2  all warps: coalesced write-back of state/results, 16-byte zero-fill of the tile's contiguous
//            mask / node_feature regions followed by a sparse scatter of the few ones, warp-per-env belief.
// The kernel is bound by the HBM writes of phase 2; graph tables are a few hundred KB
// and are read through L1/L2 with ld.global.nc.
#include "../../include/sy_env.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(SY_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

constexpr int TILE = 32;       // envs per CTA (one warp-width, so phase 1 is lane = env)
constexpr int THREADS = 256;   // 8 warps
constexpr int NWARPS = THREADS / 32;
constexpr int AS = SY_MAX_AGENTS + 1;  // odd smem row stride -> conflict-free lane = env access
constexpr unsigned FULL = 0xffffffffu;
enum { DIST_INF = 0xFFFF };

enum { RNG_RESET_POS = 0, RNG_RESET_GRAPH = 1, RNG_ACTION = 2 };
enum { ST_RUNNING = 0, ST_CAPTURE = 1, ST_TIMEOUT = 2, ST_NO_MONEY = 3 };
enum { BEL_KEEP = 0, BEL_UNIFORM = 1, BEL_DELTA = 2, BEL_PROPAGATE = 3 };

struct Tables {
  const uint8_t* W;      // [G, N, Ns] dense edge weight, 0 = no edge
  const uint16_t* D;     // [G, N, N]  all-pairs shortest path, 0xFFFF = unreachable
  const int32_t* row_ptr;  // [G, N+1]
  const uint16_t* col;   // [G, nnz_stride]
  const uint8_t* wgt;    // [G, nnz_stride]
  const uint8_t* cnt;    // [G, N, wcap+1] #neighbours with weight <= c
  const float* inv_deg;  // [G, N]
  const double* exp_neg; // [n_exp]
  const double* coverage;  // [n_cov]
  int n_exp, n_cov, G, Ns, nnz_stride, wcap;
};

struct Params {
  int B, N, P, A;
  int agent_money, mrx_money, max_t, reveal, toll, belief_on, auto_reset, resample_graph, reward_mode;
  unsigned long long env_offset;
  unsigned seed_lo, seed_hi;
  double w64[SY_NUM_REWARD_WEIGHTS];
  float w32[SY_NUM_REWARD_WEIGHTS];
  Tables tb;
  SyState st;
  SyObs ob;
  SyOut out;
  const long long* actions;
  // reset-only inputs
  const uint8_t* reset_mask;
  const int32_t* init_pos;
  const int32_t* init_gid;
  int restart;
};

struct TileSmem {
  int act[TILE * AS];
  int pos[TILE * AS];
  int money[TILE * AS];
  float rew[TILE * AS];
  double rew64[TILE * AS];
  int t[TILE];
  int gid[TILE];
  int episode[TILE];
  int status[TILE];
  int frozen[TILE];
  int done[TILE];
  int clear_visits[TILE];
  int bel_op[TILE];
  int revealed[TILE];
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (same constants / counter layout as oracle/sy_oracle.py:philox4x32)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ unsigned word_of(const uint4& r, int i) {
  return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w));
}

// A distinct uniform nodes (distribution of np.random.choice(N, A, replace=False), yard.py:112-116)
__device__ void philox_start_positions(const Params& p, unsigned env, unsigned episode, int* out /*stride 1*/) {
  int chosen[SY_MAX_AGENTS];
  uint4 r = make_uint4(0, 0, 0, 0);
  const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
  for (int a = 0; a < p.A; ++a) {
    if ((a & 3) == 0) r = philox4x32(make_uint4(env, episode, RNG_RESET_POS, (unsigned)(a >> 2)), key);
    int x = (int)__umulhi(word_of(r, a & 3), (unsigned)(p.N - a));
    for (int i = 0; i < a; ++i)
      if (x >= chosen[i]) ++x;
    int j = a;
    while (j > 0 && chosen[j - 1] > x) {
      chosen[j] = chosen[j - 1];
      --j;
    }
    chosen[j] = x;
    out[a] = x;
  }
}

__device__ __forceinline__ int philox_graph_choice(const Params& p, unsigned env, unsigned episode) {
  uint4 r = philox4x32(make_uint4(env, episode, RNG_RESET_GRAPH, 0u), make_uint2(p.seed_lo, p.seed_hi));
  return (int)__umulhi(r.x, (unsigned)p.tb.G);
}

// ---------------------------------------------------------------------------------------------
// table lookups (read-only path; tables are tiny and L1/L2 resident)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int edge_weight(const Tables& tb, int N, int g, int u, int v) {
  return __ldg(tb.W + ((size_t)g * N + u) * tb.Ns + v);
}
__device__ __forceinline__ int dist_of(const Tables& tb, int N, int g, int u, int v) {
  return __ldg(tb.D + ((size_t)g * N + u) * N + v);
}
// len(_get_possible_moves(pos, .)[0]) for a budget (yard.py:420-472), with the toll extension
__device__ __forceinline__ int move_count(const Tables& tb, int N, int g, int u, int money, int toll) {
  int c = money - toll;
  if (c <= 0) return 0;
  c = min(c, tb.wcap);
  return __ldg(tb.cnt + ((size_t)g * N + u) * (tb.wcap + 1) + c);
}
__device__ __forceinline__ double exp_neg(const Tables& tb, int d) {
  return (d < tb.n_exp) ? __ldg(tb.exp_neg + d) : 0.0;  // d == 0xFFFF (unreachable) -> exp(-inf) = 0
}

// ---------------------------------------------------------------------------------------------
// phase 2: everything that is written per env and is a pure function of the tile state in smem
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void zero_fill_bytes(uint8_t* base, size_t nbytes, int tid) {
  // base is 16-byte aligned (tile starts are multiples of 32 envs); the tail of the last tile is bytewise
  const size_t nvec = nbytes >> 4;
  uint4* v = reinterpret_cast<uint4*>(base);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = tid; i < nvec; i += THREADS) __stcs(v + i, z);
  for (size_t i = (nvec << 4) + tid; i < nbytes; i += THREADS) base[i] = 0;
}

__device__ void assemble_tile(const Params& p, TileSmem& s, float* sbel_all, int tile0, int nEnv) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, A = p.A;
  const Tables& tb = p.tb;

  // ---- state write-back + small per-agent observations (coalesced from smem)
  for (int i = tid; i < nEnv * A; i += THREADS) {
    const int e = i / A, a = i - e * A;
    const size_t o = (size_t)tile0 * A + i;
    const int m = s.money[e * AS + a];
    p.st.pos[o] = s.pos[e * AS + a];
    p.st.money[o] = m;
    p.ob.agent_budget[o] = (float)m;
  }
  if (tid < nEnv) {
    const int b = tile0 + tid;
    p.st.timestep[b] = s.t[tid];
    p.st.graph_id[b] = s.gid[tid];
    p.st.episode[b] = s.episode[tid];
    p.st.done[b] = (uint8_t)s.done[tid];
    p.ob.mrx_revealed[b] = s.revealed[tid];
  }

  // ---- dense observations: zero the tile's contiguous regions, then scatter the ones
  uint8_t* mask_base = p.ob.action_mask + (size_t)tile0 * A * N;
  float* nf_base = p.ob.node_features + (size_t)tile0 * N * A;
  zero_fill_bytes(mask_base, (size_t)nEnv * A * N, tid);
  zero_fill_bytes(reinterpret_cast<uint8_t*>(nf_base), (size_t)nEnv * N * A * sizeof(float), tid);
  __syncthreads();  // orders the zero stores before the ones below (same CTA)
  for (int i = tid; i < nEnv * A; i += THREADS) {
    const int e = i / A, a = i - e * A;
    const int g = s.gid[e], u = s.pos[e * AS + a], m = s.money[e * AS + a];
    // action_mask.py:65-76: adjacent and weight + toll <= budget
    const int r0 = __ldg(tb.row_ptr + (size_t)g * (N + 1) + u), r1 = __ldg(tb.row_ptr + (size_t)g * (N + 1) + u + 1);
    uint8_t* row = mask_base + (size_t)i * N;
    for (int k = r0; k < r1; ++k) {
      const int v = __ldg(tb.col + (size_t)g * tb.nnz_stride + k);
      const int w = __ldg(tb.wgt + (size_t)g * tb.nnz_stride + k);
      if (w + p.toll <= m) row[v] = 1;
    }
    // yard.py:279-290: one-hot positions; the MrX column stays blank while he is hidden
    if (a > 0 || s.revealed[e] >= 0) nf_base[((size_t)e * N + u) * A + a] = 1.0f;
  }

  // ---- visit counts are cleared on reset (yard.py:85)
  for (int e = warp; e < nEnv; e += NWARPS) {
    if (s.clear_visits[e]) {
      uint16_t* v = p.st.visits + (size_t)(tile0 + e) * N;
      for (int j = lane; j < N; j += 32) v[j] = 0;
    }
  }

  // ---- belief_map (belief_module.py:69-111 in expectation), one warp per env
  if (p.belief_on) {
    float* sb = sbel_all + (size_t)warp * 2 * N;  // [0,N): b[i]/deg(i)   [N,2N): un-normalised result
    for (int e = warp; e < nEnv; e += NWARPS) {
      const int op = s.bel_op[e];
      float* bel = p.st.belief + (size_t)(tile0 + e) * N;
      if (op == BEL_UNIFORM) {
        const float u = 1.0f / (float)N;
        for (int j = lane; j < N; j += 32) bel[j] = u;
      } else if (op == BEL_DELTA) {
        const int x = s.pos[e * AS + 0];
        for (int j = lane; j < N; j += 32) bel[j] = (j == x) ? 1.0f : 0.0f;
      } else if (op == BEL_PROPAGATE) {
        const int g = s.gid[e];
        const float* idg = tb.inv_deg + (size_t)g * N;
        const int32_t* rp = tb.row_ptr + (size_t)g * (N + 1);
        const uint16_t* cl = tb.col + (size_t)g * tb.nnz_stride;
        __syncwarp();
        for (int j = lane; j < N; j += 32) sb[j] = bel[j] * __ldg(idg + j);
        __syncwarp();
        float part = 0.0f;
        for (int j = lane; j < N; j += 32) {
          const int r0 = __ldg(rp + j), r1 = __ldg(rp + j + 1);
          float acc = 0.0f;
          for (int k = r0; k < r1; ++k) acc += sb[__ldg(cl + k)];
          if (r1 == r0) acc = bel[j];  // isolated node keeps its mass (belief_module.py:93-97)
          sb[N + j] = acc;
          part += acc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
        if (part == 0.0f) {  // belief_module.py:36-37
          const float u = 1.0f / (float)N;
          for (int j = lane; j < N; j += 32) bel[j] = u;
        } else {
          for (int j = lane; j < N; j += 32) bel[j] = sb[N + j] / part;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// step kernel
// ---------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void agent_reward(const Params& p, TileSmem& s, int e, int b, int a) {
  // lane = env e, this warp = agent a.  reward_calculator.py:26-92 (status decided in phase 1a).
  const Tables& tb = p.tb;
  const int N = p.N, P = p.P;
  const int g = s.gid[e];
  const int status = s.status[e];
  const int u = s.pos[e * AS + a];
  int visits_here = 0;
  if (a > 0) {  // yard.py:244-245: node_visit_counts[pos] += 1 for every police, before the rewards
    uint16_t* vp = p.st.visits + (size_t)b * N + u;
    visits_here = (int)(*vp) + 1;
    *vp = (uint16_t)min(visits_here, 0xFFFF);
  }
  double r64;
  float r32;
  if (status == ST_CAPTURE) {
    r64 = (a == 0) ? -1.0 : 1.0;
    r32 = (float)r64;
  } else if (status != ST_RUNNING) {
    r64 = (a == 0) ? 1.0 : 0.0;
    r32 = (float)r64;
  } else {
    const double tt = (double)s.t[e];
    const int x = s.pos[e * AS + 0];
    if (a == 0) {
      // reward_calculator.py:126-148
      int dmin = DIST_INF;
      long long dsum = 0;
      bool any_inf = false;
      for (int i = 1; i <= P; ++i) {
        const int d = dist_of(tb, N, g, x, s.pos[e * AS + i]);
        dmin = min(dmin, d);
        any_inf |= (d == DIST_INF);
        dsum += d;
      }
      const double inf = __longlong_as_double(0x7ff0000000000000LL);
      const double closest = (dmin == DIST_INF) ? inf : (double)dmin;
      const double avg = any_inf ? inf : __ddiv_rn((double)dsum, (double)P);  // np.mean: exact sum / P
      const double x1 = __ddiv_rn(-1.0, __dadd_rn(closest, 1.0));
      const double x2 = __ddiv_rn(-1.0, __dadd_rn(avg, 1.0));
      const double x3 = (double)move_count(tb, N, g, x, s.money[e * AS + 0], p.toll);
      const double x4 = __dmul_rn(0.1, tt);
      if (MODE == SY_REWARD_FP64) {
        const double t1 = __dmul_rn(p.w64[4], x1), t2 = __dmul_rn(p.w64[5], x2), t3 = __dmul_rn(p.w64[6], x3);
        const double t4 = __dmul_rn(__dsub_rn(1.0, p.w64[7]), x4);
        r64 = __dadd_rn(__dadd_rn(__dadd_rn(t1, t2), t3), t4);
        r32 = (float)r64;
      } else {
        const float t1 = __fmul_rn(p.w32[4], (float)x1), t2 = __fmul_rn(p.w32[5], (float)x2);
        const float t3 = __fmul_rn(p.w32[6], (float)x3);
        const float t4 = __fmul_rn(__fsub_rn(1.0f, p.w32[7]), (float)x4);
        r32 = __fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), t4);
        r64 = (double)r32;
      }
    } else {
      // reward_calculator.py:182-227
      const int k = a - 1;
      const double dx = exp_neg(tb, dist_of(tb, N, g, u, x));
      double grp = 0.0, ov = 0.0, prox = 0.0;
      for (int j = 0; j < P; ++j) {
        if (j == k) continue;
        const int d = dist_of(tb, N, g, u, s.pos[e * AS + 1 + j]);
        const double ex = exp_neg(tb, d);
        grp = __dadd_rn(grp, ex);
        if (d <= 1)
          ov = __dadd_rn(ov, 1.0);
        else
          prox = __dadd_rn(prox, ex);
      }
      // QUIRK reward_calculator.py:190: the mobility term uses the budget of agent index k (not k+1)
      const double mob = (double)move_count(tb, N, g, u, s.money[e * AS + k], p.toll);
      const double cov = __ldg(tb.coverage + min(visits_here, tb.n_cov - 1));
      const double x4 = __dmul_rn(0.05, tt);
      if (MODE == SY_REWARD_FP64) {
        const double u1 = __dmul_rn(p.w64[0], dx), u2 = __dmul_rn(p.w64[1], grp), u3 = __dmul_rn(p.w64[2], mob);
        const double u4 = __dmul_rn(__dsub_rn(1.0, p.w64[3]), x4);
        const double u5 = __dmul_rn(p.w64[9], prox), u6 = __dmul_rn(p.w64[10], ov), u7 = __dmul_rn(p.w64[8], cov);
        r64 = __dadd_rn(__dsub_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(u1, u2), u3), u4), u5), u6), u7);
        r32 = (float)r64;
      } else {
        const float u1 = __fmul_rn(p.w32[0], (float)dx), u2 = __fmul_rn(p.w32[1], (float)grp);
        const float u3 = __fmul_rn(p.w32[2], (float)mob);
        const float u4 = __fmul_rn(__fsub_rn(1.0f, p.w32[3]), (float)x4);
        const float u5 = __fmul_rn(p.w32[9], (float)prox), u6 = __fmul_rn(p.w32[10], (float)ov);
        const float u7 = __fmul_rn(p.w32[8], (float)cov);
        r32 = __fadd_rn(__fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(u1, u2), u3), u4), u5), u6), u7);
        r64 = (double)r32;
      }
    }
  }
  s.rew[e * AS + a] = r32;
  s.rew64[e * AS + a] = r64;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS) sy_step_kernel(const Params p) {
  __shared__ TileSmem s;
  extern __shared__ float sbel[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, P = p.P, A = p.A;
  const int tile0 = blockIdx.x * TILE;
  const int nEnv = min(TILE, p.B - tile0);
  const Tables& tb = p.tb;

  // ---- phase 0: stage the tile (coalesced)
  for (int i = tid; i < nEnv * A; i += THREADS) {
    const int e = i / A, a = i - e * A;
    const size_t o = (size_t)tile0 * A + i;
    const long long a64 = p.actions[o];
    s.act[e * AS + a] = (a64 >= 0 && a64 < N) ? (int)a64 : (a64 == -1 ? -1 : -2);
    s.pos[e * AS + a] = p.st.pos[o];
    s.money[e * AS + a] = p.st.money[o];
  }
  if (tid < TILE) {
    const bool live = tid < nEnv;
    const int b = tile0 + tid;
    s.t[tid] = live ? p.st.timestep[b] : 0;
    s.gid[tid] = live ? p.st.graph_id[b] : 0;
    s.episode[tid] = live ? p.st.episode[b] : 0;
    s.frozen[tid] = live ? (int)p.st.done[b] : 1;
    s.done[tid] = s.frozen[tid];
    s.status[tid] = ST_RUNNING;
    s.clear_visits[tid] = 0;
    s.bel_op[tid] = BEL_KEEP;
    s.revealed[tid] = -1;
  }
  __syncthreads();

  // ---- phase 1a: sequential move rule, lane = env (yard.py:155-243)
  int spent = 0;
  if (warp == 0 && lane < nEnv && !s.frozen[lane]) {
    const int e = lane, g = s.gid[e];
    int* pos = s.pos + e * AS;
    int* money = s.money + e * AS;
    const int* act = s.act + e * AS;
    {  // MrX: legal target or stay; may not step onto a police node (yard.py:161-188)
      const int a0 = act[0], u = pos[0];
      int tgt = u;
      if (a0 >= 0) {
        const int w = edge_weight(tb, N, g, u, a0);
        if (w > 0 && w + p.toll <= money[0]) tgt = a0;
      }
      bool occupied = false;
      for (int i = 1; i <= P; ++i) occupied |= (pos[i] == tgt);
      if (!occupied) pos[0] = tgt;
    }
    bool no_money = true;
    for (int i = 1; i <= P; ++i) {  // police in order; later police see earlier moves (yard.py:192-243)
      const int ai = act[i], m = money[i], u = pos[i];
      if (ai == -1 || m == 0) continue;  // None / DEFAULT_ACTION / broke: skipped (yard.py:210-215)
      no_money = false;
      if (ai < 0 || ai == u) continue;
      const int w = edge_weight(tb, N, g, u, ai);
      if (w == 0 || w + p.toll > m) continue;  // not a possible move -> stay
      bool occupied = false;
      for (int j = 1; j <= P; ++j) occupied |= (pos[j] == ai);
      if (occupied) continue;  // may step onto MrX (capture) but not onto police (yard.py:231)
      pos[i] = ai;
      money[i] = m - (w + p.toll);
      spent += w + p.toll;
    }
    bool capture = false;
    for (int i = 1; i <= P; ++i) capture |= (pos[i] == pos[0]);
    // reward_calculator.py:63-79; `timestep` is the pre-increment value (yard.py:345,355)
    s.status[e] = capture ? ST_CAPTURE : (s.t[e] > p.max_t ? ST_TIMEOUT : (no_money ? ST_NO_MONEY : ST_RUNNING));
  }
  __syncthreads();

  // ---- phase 1b: rewards, warp = agent, lane = env
  for (int a = warp; a < A; a += NWARPS) {
    if (lane < nEnv) {
      if (s.frozen[lane]) {
        s.rew[lane * AS + a] = 0.0f;
        s.rew64[lane * AS + a] = 0.0;
      } else {
        agent_reward<MODE>(p, s, lane, tile0 + lane, a);
      }
    }
  }
  __syncthreads();

  // ---- results of this step (coalesced), before the auto-reset rewrites the tile state
  for (int i = tid; i < nEnv * A; i += THREADS) {
    const int e = i / A, a = i - e * A;
    const size_t o = (size_t)tile0 * A + i;
    const int st = s.status[e];
    const bool term = (st == ST_CAPTURE) || (st == ST_NO_MONEY), trunc = (st == ST_TIMEOUT);
    p.out.reward[o] = s.rew[e * AS + a];
    if (p.out.reward64) p.out.reward64[o] = s.rew64[e * AS + a];
    p.out.terminated[o] = term;
    p.out.truncated[o] = trunc;
    p.out.done[o] = term || trunc || s.frozen[e];
  }
  __syncthreads();

  // ---- phase 1c: timestep, reveal, auto-reset, statistics; lane = env
  if (warp == 0) {
    int n_step = 0, n_ep = 0, n_mrx = 0, n_pol = 0, n_trunc = 0, n_broke = 0, len_sum = 0;
    if (lane < nEnv) {
      const int e = lane, b = tile0 + e;
      const int st = s.status[e];
      if (!s.frozen[e]) {
        n_step = 1;
        int t_new = s.t[e] + 1;  // yard.py:355
        p.out.winner[b] = (int8_t)(st == ST_CAPTURE ? SY_WINNER_POLICE : (st == ST_RUNNING ? SY_WINNER_NONE : SY_WINNER_MRX));
        int bel = BEL_PROPAGATE;
        if (st != ST_RUNNING) {
          n_ep = 1;
          n_pol = (st == ST_CAPTURE);
          n_mrx = (st != ST_CAPTURE);
          n_trunc = (st == ST_TIMEOUT);
          n_broke = (st == ST_NO_MONEY);
          len_sum = t_new;
          if (p.auto_reset) {  // same-step auto-reset: the observation describes the fresh episode
            const unsigned env_id = (unsigned)(p.env_offset + (unsigned long long)b);
            const int ep = s.episode[e] + 1;
            s.episode[e] = ep;
            if (p.resample_graph) s.gid[e] = philox_graph_choice(p, env_id, (unsigned)ep);
            philox_start_positions(p, env_id, (unsigned)ep, s.pos + e * AS);
            s.money[e * AS] = p.mrx_money;
            for (int i = 1; i <= P; ++i) s.money[e * AS + i] = p.agent_money;
            t_new = 0;
            s.clear_visits[e] = 1;
            bel = BEL_UNIFORM;
          } else {
            s.done[e] = 1;
          }
        }
        s.t[e] = t_new;
        // reveal schedule (src/eval/run_ablations.py:225-229) on the new timestep
        const bool rev = p.reveal > 0 && t_new > 0 && (t_new % p.reveal) == 0;
        if (rev && bel == BEL_PROPAGATE) bel = BEL_DELTA;
        s.bel_op[e] = bel;
        s.revealed[e] = (p.reveal <= 0 || rev) ? s.pos[e * AS] : -1;
      } else {
        p.out.winner[b] = SY_WINNER_NONE;
        const int t_cur = s.t[e];
        const bool rev = p.reveal > 0 && t_cur > 0 && (t_cur % p.reveal) == 0;
        s.revealed[e] = (p.reveal <= 0 || rev) ? s.pos[e * AS] : -1;
      }
    }
    if (p.out.stats) {
      n_step = __reduce_add_sync(FULL, n_step);
      n_ep = __reduce_add_sync(FULL, n_ep);
      n_mrx = __reduce_add_sync(FULL, n_mrx);
      n_pol = __reduce_add_sync(FULL, n_pol);
      n_trunc = __reduce_add_sync(FULL, n_trunc);
      n_broke = __reduce_add_sync(FULL, n_broke);
      len_sum = __reduce_add_sync(FULL, len_sum);
      spent = __reduce_add_sync(FULL, spent);
      if (lane == 0) {
        unsigned long long* st = reinterpret_cast<unsigned long long*>(p.out.stats);
        atomicAdd(st + SY_STAT_ENV_STEPS, (unsigned long long)n_step);
        if (n_ep) {
          atomicAdd(st + SY_STAT_EPISODES, (unsigned long long)n_ep);
          atomicAdd(st + SY_STAT_MRX_WINS, (unsigned long long)n_mrx);
          atomicAdd(st + SY_STAT_POLICE_WINS, (unsigned long long)n_pol);
          atomicAdd(st + SY_STAT_TRUNCATIONS, (unsigned long long)n_trunc);
          atomicAdd(st + SY_STAT_OUT_OF_MONEY, (unsigned long long)n_broke);
          atomicAdd(st + SY_STAT_SUM_EPISODE_LENGTH, (unsigned long long)len_sum);
        }
        if (spent) atomicAdd(st + SY_STAT_SUM_BUDGET_SPENT, (unsigned long long)spent);
      }
    }
  }
  __syncthreads();

  // ---- phase 2
  assemble_tile(p, s, sbel, tile0, nEnv);
}

// ---------------------------------------------------------------------------------------------
// reset kernel (yard.py:80-142): re-initialise the masked envs, rewrite every env's observations
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) sy_reset_kernel(const Params p) {
  __shared__ TileSmem s;
  extern __shared__ float sbel[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int A = p.A, P = p.P;
  const int tile0 = blockIdx.x * TILE;
  const int nEnv = min(TILE, p.B - tile0);

  if (tid < TILE) {
    const bool live = tid < nEnv;
    const int b = tile0 + tid;
    const bool rst = live && (p.reset_mask == nullptr || p.reset_mask[b] != 0);
    s.clear_visits[tid] = rst;
    s.bel_op[tid] = rst ? BEL_UNIFORM : BEL_KEEP;
    s.status[tid] = ST_RUNNING;
    s.frozen[tid] = 0;
    if (rst) {
      s.t[tid] = 0;
      s.episode[tid] = p.restart ? 0 : p.st.episode[b] + 1;
      s.done[tid] = 0;
      s.gid[tid] = p.init_gid ? p.init_gid[b] : p.st.graph_id[b];
    } else {
      s.t[tid] = live ? p.st.timestep[b] : 0;
      s.episode[tid] = live ? p.st.episode[b] : 0;
      s.done[tid] = live ? (int)p.st.done[b] : 1;
      s.gid[tid] = live ? p.st.graph_id[b] : 0;
    }
  }
  __syncthreads();
  for (int i = tid; i < nEnv * A; i += THREADS) {
    const int e = i / A, a = i - e * A;
    const size_t o = (size_t)tile0 * A + i;
    if (s.clear_visits[e]) {
      s.money[e * AS + a] = (a == 0) ? p.mrx_money : p.agent_money;  // yard.py:117-119
      s.pos[e * AS + a] = p.init_pos ? p.init_pos[o] : 0;
    } else {
      s.money[e * AS + a] = p.st.money[o];
      s.pos[e * AS + a] = p.st.pos[o];
    }
  }
  __syncthreads();
  if (warp == 0 && lane < nEnv) {
    const int e = lane, b = tile0 + e;
    if (s.clear_visits[e] && p.init_pos == nullptr) {
      const unsigned env_id = (unsigned)(p.env_offset + (unsigned long long)b);
      if (p.resample_graph && p.init_gid == nullptr) s.gid[e] = philox_graph_choice(p, env_id, (unsigned)s.episode[e]);
      philox_start_positions(p, env_id, (unsigned)s.episode[e], s.pos + e * AS);
    }
    const int t_cur = s.t[e];
    const bool rev = p.reveal > 0 && t_cur > 0 && (t_cur % p.reveal) == 0;
    s.revealed[e] = (p.reveal <= 0 || rev) ? s.pos[e * AS] : -1;
  }
  (void)P;
  __syncthreads();
  assemble_tile(p, s, sbel, tile0, nEnv);
}

// ---------------------------------------------------------------------------------------------
// random valid policy: thread per (env, agent)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sy_sample_actions_kernel(const Params p, unsigned step_counter, long long* actions) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)p.B * p.A) return;
  const int b = (int)(i / p.A), a = (int)(i - (size_t)b * p.A);
  const Tables& tb = p.tb;
  const int N = p.N;
  const int g = p.st.graph_id[b], u = p.st.pos[i], m = p.st.money[i];
  const int r0 = __ldg(tb.row_ptr + (size_t)g * (N + 1) + u), r1 = __ldg(tb.row_ptr + (size_t)g * (N + 1) + u + 1);
  const uint8_t* wg = tb.wgt + (size_t)g * tb.nnz_stride;
  int nvalid = 0;
  for (int k = r0; k < r1; ++k) nvalid += (__ldg(wg + k) + p.toll <= m);
  long long act = -1;
  if (nvalid > 0) {
    const unsigned env_id = (unsigned)(p.env_offset + (unsigned long long)b);
    const uint4 r = philox4x32(make_uint4(env_id, step_counter, RNG_ACTION, (unsigned)a), make_uint2(p.seed_lo, p.seed_hi));
    int pick = (int)__umulhi(r.x, (unsigned)nvalid);
    for (int k = r0; k < r1; ++k) {
      if (__ldg(wg + k) + p.toll <= m) {
        if (pick == 0) {
          act = __ldg(tb.col + (size_t)g * tb.nnz_stride + k);
          break;
        }
        --pick;
      }
    }
  }
  actions[i] = act;
}

// ---------------------------------------------------------------------------------------------
// graph table construction
// ---------------------------------------------------------------------------------------------
// one CTA per (graph, node): dense weight row, move-count row, 1/deg
__global__ void sy_build_rows_kernel(int N, int Ns, int nnz_stride, int wcap, const int32_t* row_ptr,
                                     const uint16_t* col, const uint8_t* wgt, uint8_t* W, uint8_t* cnt,
                                     float* inv_deg) {
  const int g = blockIdx.y, u = blockIdx.x;
  const int r0 = row_ptr[(size_t)g * (N + 1) + u], r1 = row_ptr[(size_t)g * (N + 1) + u + 1];
  uint8_t* row = W + ((size_t)g * N + u) * Ns;
  for (int j = threadIdx.x; j < Ns; j += blockDim.x) row[j] = 0;
  __syncthreads();
  for (int k = r0 + threadIdx.x; k < r1; k += blockDim.x) row[col[(size_t)g * nnz_stride + k]] = wgt[(size_t)g * nnz_stride + k];
  for (int c = threadIdx.x; c <= wcap; c += blockDim.x) {
    int n = 0;
    for (int k = r0; k < r1; ++k) n += (wgt[(size_t)g * nnz_stride + k] <= c);
    cnt[((size_t)g * N + u) * (wcap + 1) + c] = (uint8_t)min(n, 255);
  }
  if (threadIdx.x == 0) inv_deg[(size_t)g * N + u] = (r1 > r0) ? 1.0f / (float)(r1 - r0) : 0.0f;
}

// all-pairs shortest paths, one warp per (graph, source): in-place Bellman-Ford relaxation to a fixed point
// (replaces the per-query heap Dijkstra of pathfinding.py:34-137; weights are positive integers)
__global__ void sy_apsp_kernel(int G, int N, int nnz_stride, const int32_t* row_ptr, const uint16_t* col,
                               const uint8_t* wgt, uint16_t* D, int warps_per_block) {
  extern __shared__ int sdist_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wg = (long long)blockIdx.x * warps_per_block + warp;
  if (wg >= (long long)G * N) return;
  const int g = (int)(wg / N), src = (int)(wg - (long long)g * N);
  volatile int* dist = sdist_all + (size_t)warp * N;
  const int32_t* rp = row_ptr + (size_t)g * (N + 1);
  const uint16_t* cl = col + (size_t)g * nnz_stride;
  const uint8_t* wt = wgt + (size_t)g * nnz_stride;
  const int BIG = 0x3fffffff;
  for (int v = lane; v < N; v += 32) dist[v] = (v == src) ? 0 : BIG;
  __syncwarp();
  bool changed = true;
  while (changed) {
    changed = false;
    for (int v = lane; v < N; v += 32) {
      int best = dist[v];
      for (int k = rp[v]; k < rp[v + 1]; ++k) best = min(best, dist[cl[k]] + (int)wt[k]);
      if (best < dist[v]) {
        dist[v] = best;
        changed = true;
      }
    }
    changed = __any_sync(FULL, changed);
    __syncwarp();
  }
  uint16_t* out = D + ((size_t)g * N + src) * N;
  for (int v = lane; v < N; v += 32) out[v] = (uint16_t)(dist[v] >= DIST_INF ? DIST_INF : dist[v]);
}

// ---------------------------------------------------------------------------------------------
// dense float64 action mask == compute_action_mask (action_mask.py:30-83); one CTA per query
// ---------------------------------------------------------------------------------------------
__global__ void sy_mask_dense_kernel(int N, const double* adj, const double* w, const double* toll_m, double toll_s,
                                     const int32_t* cur, const double* budget, uint8_t* out) {
  const int q = blockIdx.x;
  const int u = cur[q];
  const double bud = budget[q];
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const double a = adj[(size_t)u * N + j];
    const double wt = w ? w[(size_t)u * N + j] : a;            // action_mask.py:110-112
    const double tl = toll_m ? toll_m[(size_t)u * N + j] : toll_s;
    const double cost = __dadd_rn(wt, tl);                     // action_mask.py:72
    out[(size_t)q * N + j] = (j != u) && (a != 0.0) && (cost <= bud);
  }
}

}  // namespace

// =============================================================================================
// host side
// =============================================================================================
struct SyEnv {
  SyConfig cfg;
  int A = 0;
  bool graphs_loaded = false;
  bool tables_set = false;
  Tables tb{};
  // device allocations owned by the handle
  void* d_W = nullptr;
  void* d_D = nullptr;
  void* d_row_ptr = nullptr;
  void* d_col = nullptr;
  void* d_wgt = nullptr;
  void* d_cnt = nullptr;
  void* d_inv_deg = nullptr;
  void* d_exp = nullptr;
  void* d_cov = nullptr;
  size_t bel_smem = 0;
};

namespace {

void free_graph_tables(SyEnv* e) {
  for (void** ptr : {&e->d_W, &e->d_D, &e->d_row_ptr, &e->d_col, &e->d_wgt, &e->d_cnt, &e->d_inv_deg}) {
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
  }
  e->graphs_loaded = false;
}

int fill_params(const SyEnv* env, const SyState* st, const SyObs* ob, const SyOut* out, Params& p) {
  const SyConfig& c = env->cfg;
  if (!env->graphs_loaded) return fail(SY_ERR_STATE, "sy_load_graphs must be called first");
  if (!env->tables_set) return fail(SY_ERR_STATE, "sy_set_reward_tables must be called first");
  if (!st || !st->pos || !st->money || !st->timestep || !st->graph_id || !st->episode || !st->done || !st->visits)
    return fail(SY_ERR_INVALID_ARGUMENT, "SyState has NULL members");
  if (c.belief && !st->belief) return fail(SY_ERR_INVALID_ARGUMENT, "SyState.belief is NULL but config.belief is on");
  std::memset(&p, 0, sizeof(p));
  p.B = c.num_envs;
  p.N = c.num_nodes;
  p.P = c.num_police;
  p.A = env->A;
  p.agent_money = c.agent_money;
  p.mrx_money = c.mrx_money;
  p.max_t = c.max_timestep;
  p.reveal = c.reveal_interval;
  p.toll = c.toll;
  p.belief_on = c.belief;
  p.auto_reset = c.auto_reset;
  p.resample_graph = c.resample_graph;
  p.reward_mode = c.reward_mode;
  p.env_offset = (unsigned long long)c.env_offset;
  p.seed_lo = (unsigned)(c.seed & 0xffffffffu);
  p.seed_hi = (unsigned)(c.seed >> 32);
  for (int i = 0; i < SY_NUM_REWARD_WEIGHTS; ++i) {
    p.w64[i] = c.reward_weights[i];
    p.w32[i] = (float)c.reward_weights[i];
  }
  p.tb = env->tb;
  p.st = *st;
  if (ob) p.ob = *ob;
  if (out) p.out = *out;
  return SY_OK;
}

bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

int check_obs(const SyObs* ob) {
  if (!ob || !ob->action_mask || !ob->node_features || !ob->agent_budget || !ob->mrx_revealed)
    return fail(SY_ERR_INVALID_ARGUMENT, "SyObs has NULL members");
  if (!aligned16(ob->action_mask) || !aligned16(ob->node_features))
    return fail(SY_ERR_INVALID_ARGUMENT, "action_mask / node_features must be 16-byte aligned");
  return SY_OK;
}

}  // namespace

extern "C" {

int sy_abi_version(void) { return SY_ABI_VERSION; }
const char* sy_last_error(void) { return g_err.c_str(); }
int64_t sy_launch_count(void) { return (int64_t)g_launches.load(); }

int sy_create(const SyConfig* c, SyEnv** out_env) {
  if (!c || !out_env) return fail(SY_ERR_INVALID_ARGUMENT, "NULL config / out_env");
  if (c->struct_bytes != (int32_t)sizeof(SyConfig))
    return fail(SY_ERR_INVALID_ARGUMENT, "SyConfig size mismatch: got %d, library expects %zu", c->struct_bytes, sizeof(SyConfig));
  if (c->num_envs <= 0) return fail(SY_ERR_INVALID_ARGUMENT, "num_envs must be positive");
  if (c->num_police < 1 || c->num_police + 1 > SY_MAX_AGENTS)
    return fail(SY_ERR_INVALID_ARGUMENT, "num_police must be in [1, %d]", SY_MAX_AGENTS - 1);
  if (c->num_nodes < c->num_police + 1 || c->num_nodes > 65534)
    return fail(SY_ERR_INVALID_ARGUMENT, "num_nodes must be in [num_police + 1, 65534]");
  if (c->agent_money < 0 || c->mrx_money < 0 || c->toll < 0 || c->reveal_interval < 0 || c->max_timestep < 0)
    return fail(SY_ERR_INVALID_ARGUMENT, "negative money / toll / reveal_interval / max_timestep");
  if (c->reward_mode != SY_REWARD_FP64 && c->reward_mode != SY_REWARD_FP32)
    return fail(SY_ERR_INVALID_ARGUMENT, "unknown reward_mode %d", c->reward_mode);
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (c->device < 0 || c->device >= ndev) return fail(SY_ERR_INVALID_ARGUMENT, "device %d out of range (%d visible)", c->device, ndev);
  CUDA_TRY(cudaSetDevice(c->device));
  SyEnv* e = new SyEnv();
  e->cfg = *c;
  e->A = c->num_police + 1;
  e->bel_smem = c->belief ? (size_t)NWARPS * 2 * c->num_nodes * sizeof(float) : 0;
  if (e->bel_smem > 160 * 1024) {
    delete e;
    return fail(SY_ERR_INVALID_ARGUMENT, "num_nodes too large for the belief kernel's shared memory");
  }
  if (e->bel_smem > 32 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(sy_step_kernel<SY_REWARD_FP64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->bel_smem);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(sy_step_kernel<SY_REWARD_FP32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->bel_smem);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(sy_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->bel_smem);
    if (err != cudaSuccess) {
      delete e;
      return fail(SY_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(err));
    }
  }
  *out_env = e;
  return SY_OK;
}

void sy_destroy(SyEnv* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  free_graph_tables(e);
  if (e->d_exp) cudaFree(e->d_exp);
  if (e->d_cov) cudaFree(e->d_cov);
  delete e;
}

int sy_set_seed(SyEnv* e, uint64_t seed) {
  if (!e) return fail(SY_ERR_INVALID_ARGUMENT, "NULL env");
  e->cfg.seed = seed;
  return SY_OK;
}

int sy_set_reward_tables(SyEnv* e, const double* exp_neg_h, int32_t n_exp, const double* cov_h, int32_t n_cov, sy_stream_t stream) {
  if (!e || !exp_neg_h || !cov_h || n_exp < 1 || n_cov < 2) return fail(SY_ERR_INVALID_ARGUMENT, "bad reward tables");
  if (n_exp > DIST_INF) n_exp = DIST_INF;  // index 0xFFFF must read as exp(-inf) = 0
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (e->d_exp) cudaFree(e->d_exp);
  if (e->d_cov) cudaFree(e->d_cov);
  e->d_exp = e->d_cov = nullptr;
  CUDA_TRY(cudaMalloc(&e->d_exp, (size_t)n_exp * sizeof(double)));
  CUDA_TRY(cudaMalloc(&e->d_cov, (size_t)n_cov * sizeof(double)));
  CUDA_TRY(cudaMemcpyAsync(e->d_exp, exp_neg_h, (size_t)n_exp * sizeof(double), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(e->d_cov, cov_h, (size_t)n_cov * sizeof(double), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  e->tb.exp_neg = (const double*)e->d_exp;
  e->tb.coverage = (const double*)e->d_cov;
  e->tb.n_exp = n_exp;
  e->tb.n_cov = n_cov;
  e->tables_set = true;
  return SY_OK;
}

int sy_load_graphs(SyEnv* e, int32_t G, const int32_t* row_ptr, const int32_t* col, const int32_t* w, int32_t nnz_stride, sy_stream_t stream) {
  if (G > 65535) return fail(SY_ERR_INVALID_ARGUMENT, "at most 65535 graphs in the pool");
  if (!e || G <= 0 || !row_ptr || !col || !w || nnz_stride <= 0) return fail(SY_ERR_INVALID_ARGUMENT, "bad graph arguments");
  const int N = e->cfg.num_nodes;
  // validate + narrow on the host (setup path)
  std::vector<uint16_t> col16((size_t)G * nnz_stride, 0);
  std::vector<uint8_t> w8((size_t)G * nnz_stride, 0);
  int wcap = 1;
  for (int g = 0; g < G; ++g) {
    const int32_t* rp = row_ptr + (size_t)g * (N + 1);
    if (rp[0] != 0) return fail(SY_ERR_INVALID_ARGUMENT, "graph %d: row_ptr[0] != 0", g);
    for (int u = 0; u < N; ++u) {
      if (rp[u + 1] < rp[u] || rp[u + 1] > nnz_stride) return fail(SY_ERR_INVALID_ARGUMENT, "graph %d: bad row_ptr at node %d", g, u);
      if (rp[u + 1] - rp[u] > 255) return fail(SY_ERR_INVALID_ARGUMENT, "graph %d: node %d has more than 255 neighbours", g, u);
      for (int k = rp[u]; k < rp[u + 1]; ++k) {
        const int v = col[(size_t)g * nnz_stride + k], wt = w[(size_t)g * nnz_stride + k];
        if (v < 0 || v >= N || v == u) return fail(SY_ERR_INVALID_ARGUMENT, "graph %d: bad neighbour %d of node %d", g, v, u);
        if (k > rp[u] && v <= col[(size_t)g * nnz_stride + k - 1]) return fail(SY_ERR_INVALID_ARGUMENT, "graph %d: neighbours of node %d not strictly ascending", g, u);
        if (wt < 1 || wt > 255) return fail(SY_ERR_INVALID_ARGUMENT, "graph %d: edge weight %d outside 1..255", g, wt);
        col16[(size_t)g * nnz_stride + k] = (uint16_t)v;
        w8[(size_t)g * nnz_stride + k] = (uint8_t)wt;
        if (wt > wcap) wcap = wt;
      }
    }
  }
  if ((long long)wcap * (N - 1) >= DIST_INF) return fail(SY_ERR_INVALID_ARGUMENT, "max path length does not fit u16");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  free_graph_tables(e);
  const int Ns = (N + 15) & ~15;
  const size_t nW = (size_t)G * N * Ns, nD = (size_t)G * N * N;
  CUDA_TRY(cudaMalloc(&e->d_W, nW));
  CUDA_TRY(cudaMalloc(&e->d_D, nD * sizeof(uint16_t)));
  CUDA_TRY(cudaMalloc(&e->d_row_ptr, (size_t)G * (N + 1) * sizeof(int32_t)));
  CUDA_TRY(cudaMalloc(&e->d_col, (size_t)G * nnz_stride * sizeof(uint16_t)));
  CUDA_TRY(cudaMalloc(&e->d_wgt, (size_t)G * nnz_stride));
  CUDA_TRY(cudaMalloc(&e->d_cnt, (size_t)G * N * (wcap + 1)));
  CUDA_TRY(cudaMalloc(&e->d_inv_deg, (size_t)G * N * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(e->d_row_ptr, row_ptr, (size_t)G * (N + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(e->d_col, col16.data(), col16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(e->d_wgt, w8.data(), w8.size(), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));  // the staging vectors die at return
  sy_build_rows_kernel<<<dim3(N, G), 64, 0, s>>>(N, Ns, nnz_stride, wcap, (const int32_t*)e->d_row_ptr, (const uint16_t*)e->d_col,
                                                 (const uint8_t*)e->d_wgt, (uint8_t*)e->d_W, (uint8_t*)e->d_cnt, (float*)e->d_inv_deg);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  int wpb = (int)((96 * 1024) / ((size_t)N * sizeof(int)));
  if (wpb < 1) wpb = 1;
  if (wpb > 8) wpb = 8;
  const size_t smem = (size_t)wpb * N * sizeof(int);
  if (smem > 200 * 1024) return fail(SY_ERR_INVALID_ARGUMENT, "num_nodes too large for the APSP kernel");
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(sy_apsp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nwarps = (long long)G * N;
  const unsigned nblocks = (unsigned)((nwarps + wpb - 1) / wpb);
  sy_apsp_kernel<<<nblocks, wpb * 32, smem, s>>>(G, N, nnz_stride, (const int32_t*)e->d_row_ptr, (const uint16_t*)e->d_col,
                                                 (const uint8_t*)e->d_wgt, (uint16_t*)e->d_D, wpb);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  e->tb.W = (const uint8_t*)e->d_W;
  e->tb.D = (const uint16_t*)e->d_D;
  e->tb.row_ptr = (const int32_t*)e->d_row_ptr;
  e->tb.col = (const uint16_t*)e->d_col;
  e->tb.wgt = (const uint8_t*)e->d_wgt;
  e->tb.cnt = (const uint8_t*)e->d_cnt;
  e->tb.inv_deg = (const float*)e->d_inv_deg;
  e->tb.G = G;
  e->tb.Ns = Ns;
  e->tb.nnz_stride = nnz_stride;
  e->tb.wcap = wcap;
  e->graphs_loaded = true;
  return SY_OK;
}

int sy_read_graph_tables(SyEnv* e, int32_t g, uint8_t* weights, uint16_t* apsp, sy_stream_t stream) {
  if (!e || !e->graphs_loaded) return fail(SY_ERR_STATE, "graphs not loaded");
  if (g < 0 || g >= e->tb.G) return fail(SY_ERR_INVALID_ARGUMENT, "graph index out of range");
  const int N = e->cfg.num_nodes;
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  if (weights)
    CUDA_TRY(cudaMemcpy2DAsync(weights, N, e->tb.W + (size_t)g * N * e->tb.Ns, e->tb.Ns, N, N, cudaMemcpyDeviceToHost, s));
  if (apsp) CUDA_TRY(cudaMemcpyAsync(apsp, e->tb.D + (size_t)g * N * N, (size_t)N * N * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return SY_OK;
}

int sy_reset(SyEnv* e, const uint8_t* reset_mask, const int32_t* init_pos, const int32_t* init_gid, int32_t restart,
             const SyState* st, const SyObs* ob, sy_stream_t stream) {
  if (!e) return fail(SY_ERR_INVALID_ARGUMENT, "NULL env");
  Params p;
  int rc = fill_params(e, st, ob, nullptr, p);
  if (rc) return rc;
  if ((rc = check_obs(ob))) return rc;
  p.reset_mask = reset_mask;
  p.init_pos = init_pos;
  p.init_gid = init_gid;
  p.restart = restart;
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const unsigned grid = (unsigned)((p.B + TILE - 1) / TILE);
  sy_reset_kernel<<<grid, THREADS, e->bel_smem, (cudaStream_t)stream>>>(p);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_OK;
}

int sy_step(SyEnv* e, const int64_t* actions, const SyState* st, const SyObs* ob, const SyOut* out, sy_stream_t stream) {
  if (!e || !actions) return fail(SY_ERR_INVALID_ARGUMENT, "NULL env / actions");
  if (!out || !out->reward || !out->terminated || !out->truncated || !out->done || !out->winner)
    return fail(SY_ERR_INVALID_ARGUMENT, "SyOut has NULL members");
  Params p;
  int rc = fill_params(e, st, ob, out, p);
  if (rc) return rc;
  if ((rc = check_obs(ob))) return rc;
  p.actions = reinterpret_cast<const long long*>(actions);
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const unsigned grid = (unsigned)((p.B + TILE - 1) / TILE);
  cudaStream_t s = (cudaStream_t)stream;
  if (e->cfg.reward_mode == SY_REWARD_FP64)
    sy_step_kernel<SY_REWARD_FP64><<<grid, THREADS, e->bel_smem, s>>>(p);
  else
    sy_step_kernel<SY_REWARD_FP32><<<grid, THREADS, e->bel_smem, s>>>(p);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_OK;
}

int sy_step_host(SyEnv* e, const int64_t* actions_host, int64_t* actions_dev, const SyState* st, const SyObs* ob,
                 const SyOut* out, const SyHostOut* ho, sy_stream_t stream) {
  if (!e || !actions_host || !actions_dev || !ho) return fail(SY_ERR_INVALID_ARGUMENT, "NULL env / actions / host_out");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)e->cfg.num_envs * e->A;
  CUDA_TRY(cudaMemcpyAsync(actions_dev, actions_host, n * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  const int rc = sy_step(e, actions_dev, st, ob, out, stream);
  if (rc) return rc;
  if (ho->reward) CUDA_TRY(cudaMemcpyAsync(ho->reward, out->reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (ho->terminated) CUDA_TRY(cudaMemcpyAsync(ho->terminated, out->terminated, n, cudaMemcpyDeviceToHost, s));
  if (ho->truncated) CUDA_TRY(cudaMemcpyAsync(ho->truncated, out->truncated, n, cudaMemcpyDeviceToHost, s));
  if (ho->done) CUDA_TRY(cudaMemcpyAsync(ho->done, out->done, n, cudaMemcpyDeviceToHost, s));
  if (ho->winner) CUDA_TRY(cudaMemcpyAsync(ho->winner, out->winner, (size_t)e->cfg.num_envs, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return SY_OK;
}

int sy_sample_actions(SyEnv* e, const SyState* st, uint32_t step_counter, int64_t* actions, sy_stream_t stream) {
  if (!e || !actions) return fail(SY_ERR_INVALID_ARGUMENT, "NULL env / actions");
  Params p;
  int rc = fill_params(e, st, nullptr, nullptr, p);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const size_t n = (size_t)p.B * p.A;
  sy_sample_actions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, step_counter, reinterpret_cast<long long*>(actions));
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_OK;
}

int sy_action_mask_dense(int32_t Q, int32_t N, const double* adj, const double* w, const double* toll_m, double toll_s,
                         const int32_t* cur, const double* budget, uint8_t* out, sy_stream_t stream) {
  if (Q < 0 || N <= 0 || !adj || !cur || !budget || !out) return fail(SY_ERR_INVALID_ARGUMENT, "bad dense mask arguments");
  if (Q == 0) return SY_OK;
  sy_mask_dense_kernel<<<(unsigned)Q, 128, 0, (cudaStream_t)stream>>>(N, adj, w, toll_m, toll_s, cur, budget, out);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_OK;
}

}  // extern "C"
