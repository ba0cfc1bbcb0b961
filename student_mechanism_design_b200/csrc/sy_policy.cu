// sy_policy.cu -- batched policy forward for the reference's two shipped agents, sm_100a.
// C ABI: include/sy_policy.h.  Stand-alone: reads the env's state buffers and the graph pool through plain pointers.
//
// Reference code paths replaced (file:line under /root/reference):
//   src/agent/gnn_agent.py:45-82     GNNAgent.select_action      -> sy_gnn_act_kernel (epsilon-greedy over valid moves)
//   src/agent/gnn_agent.py:230-257   GNNModel.forward            -> q_at_node (2 x AntiSymmetricConv + ReLU, Linear)
//   src/training/utils.py:151-211    create_graph_data           -> feature columns (SY_FEATURES_REFERENCE) / the env's
//                                                                   node_features (SY_FEATURES_ENV), never materialised
//   src/agent/mappo_agent.py:6-29    AgentPolicy.forward         -> sy_mappo_act_kernel (MLP + softmax)
//   src/agent/mappo_agent.py:87-142  MappoAgent.select_action    -> sy_mappo_act_kernel (mask, renormalise, sample, log-prob)
//   src/agent/mappo_agent.py:32-44   CentralCritic.forward       -> sy_mappo_values_kernel
//
// GNN design: the input features are one-hot columns (at most A non-zero rows of x), the graph is shared, and
// epsilon-greedy only ever compares the Q values of an agent's VALID MOVES (its affordable neighbours).  So instead of
// two dense message-passing layers over all N nodes per env and agent, every candidate move gets one lane that
// evaluates Q at that node from its 2-hop in-neighbourhood: layer-1 rows are recomputed on the fly (a node nobody
// "touches" has the constant row relu(eps * tanh(bias1))), layer 2 aggregates the in-neighbours' rows first and applies
// the K x K matrices once (A (X Theta) = (A X) Theta).  Weights live in shared memory, K is padded to 4/8/16 so rows are
// float4 broadcasts, everything is fp32 like the reference.  The dense [B, 2, N] Q tensor (training targets) comes from
// the same per-node routine with lane = node.
//
// MAPPO design: the logits GEMM [rows x H] x [H x N] runs on the tensor cores (sy_mappo_act_tc_kernel: tcgen05.mma
// kind::tf32 as 3xTF32, accumulator tiles in TMEM, thread-per-row epilogue from tcgen05.ld; details above that kernel).
// sy_mappo_act_kernel is the CUDA-core form for shapes the tensor path does not take (N > 256, H > 64, obs > 16,
// degree > 16): one CTA per (64-env tile, agent), hidden layer in shared memory, thread = action node with its W2 row
// chunk in registers walking four rows at a time, one warp per row for softmax / mask / renormalise / sampling / log-prob.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/sy_policy.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int GNN_WARPS = 8;
enum { RNG_GNN_POLICY = 3, RNG_MAPPO_POLICY = 4 };

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
int g_mappo_tc = 1;        // sy_policy_set_option("mappo_tensor_cores", 0 | 1)
int* g_tc_flag = nullptr;  // device flag the tensor-core kernel raises instead of hanging (barrier timeout / degree overflow)

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) return fail(SY_POLICY_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));     \
  } while (0)

// Philox4x32-10, same counter layout as the env library (oracle/sy_oracle.py:philox4x32)
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

__device__ __forceinline__ float u01(unsigned w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }  // [0, 1), 24 bits

// ---------------------------------------------------------------------------------------------
// GNN
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int gnn_kp(int K) { return K <= 4 ? 4 : (K <= 8 ? 8 : 16); }
__host__ __device__ constexpr int gnn_model_floats(int KP) { return 2 * (2 * KP * KP + KP) + KP + 4; }

struct GnnParams {
  SyPolicyGraphs g;
  SyPolicyState st;
  const float* params;  // [2][gnn_model_floats(KP)]
  int K, feature_mode;
  float conv_eps, eps_mrx, eps_police;
  unsigned seed_lo, seed_hi, step;
};

// model layout in shared memory (floats): conv c at c * CONV: WasT [KP][KP], ThT [KP][KP], bias [KP]; then out_w [KP], out_b
template <int KP>
struct Lay {
  static constexpr int CONV = 2 * KP * KP + KP;
  static constexpr int WAS = 0, TH = KP * KP, BIAS = 2 * KP * KP;
  static constexpr int OUT_W = 2 * CONV, OUT_B = 2 * CONV + KP;
  static constexpr int MF = gnn_model_floats(KP);
};

struct GraphView {
  const int32_t* in_ptr;
  const int32_t* in_src;
  const float* in_coef;
  const float* self_coef;
};

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the fast exp / divide units: |error| ~1e-6, an order below the fp32 noise the
// reference itself has between its CPU and CUDA runs of the same model, 30x fewer instructions than tanhf
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));  // exp(2x) = 2^(2x log2 e)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

template <int KP>
__device__ __forceinline__ void row_acc(float (&acc)[KP], float xk, const float* __restrict__ row) {
#pragma unroll
  for (int c4 = 0; c4 < KP / 4; ++c4) {
    const float4 m = reinterpret_cast<const float4*>(row)[c4];
    acc[4 * c4 + 0] = fmaf(xk, m.x, acc[4 * c4 + 0]);
    acc[4 * c4 + 1] = fmaf(xk, m.y, acc[4 * c4 + 1]);
    acc[4 * c4 + 2] = fmaf(xk, m.z, acc[4 * c4 + 2]);
    acc[4 * c4 + 3] = fmaf(xk, m.w, acc[4 * c4 + 3]);
  }
}

template <int KP>
__device__ __forceinline__ void matvec_acc(float (&acc)[KP], const float (&x)[KP], const float* __restrict__ MT) {
#pragma unroll
  for (int k = 0; k < KP; ++k) row_acc<KP>(acc, x[k], MT + k * KP);
}

// layer-1 output row of node m: relu(x0 + eps * tanh(x0 Was^T + (A_hat x0) Theta^T + b)).  x0 is given as a per-node
// bitmask of feature columns (feat[node], bit k = column k is 1 there): almost every node and in-neighbour has none,
// and a node nothing touches has the constant row c1; for the others only the rows of the non-zero entries are added.
template <int KP>
__device__ __forceinline__ void x1_at(const float* __restrict__ ms, const float* __restrict__ c1, const GraphView& gv,
                                      const unsigned* __restrict__ feat, int m, float eps, float (&x1)[KP]) {
  const unsigned fm = feat[m];
  float acc[KP];
  bool touched = fm != 0;
  if (fm) {
    const float sc = __ldg(gv.self_coef + m);
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = ms[Lay<KP>::BIAS + k];
    for (unsigned bits = fm; bits; bits &= bits - 1) {  // x0[m] has column k: its Was row, and its Theta row times the self-loop weight
      const int k = __ffs(bits) - 1;
      row_acc<KP>(acc, 1.0f, ms + Lay<KP>::WAS + k * KP);
      row_acc<KP>(acc, sc, ms + Lay<KP>::TH + k * KP);
    }
  }
  const int e0 = __ldg(gv.in_ptr + m), e1 = __ldg(gv.in_ptr + m + 1);
  for (int e = e0; e < e1; ++e) {
    const unsigned fs = feat[__ldg(gv.in_src + e)];
    if (fs) {
      const float c = __ldg(gv.in_coef + e);
      if (!touched) {
        touched = true;
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = ms[Lay<KP>::BIAS + k];
      }
      for (unsigned bits = fs; bits; bits &= bits - 1) row_acc<KP>(acc, c, ms + Lay<KP>::TH + (__ffs(bits) - 1) * KP);
    }
  }
  if (!touched) {
#pragma unroll
    for (int k = 0; k < KP; ++k) x1[k] = c1[k];
    return;
  }
#pragma unroll
  for (int k = 0; k < KP; ++k) x1[k] = fmaxf((((fm >> k) & 1u) ? 1.0f : 0.0f) + eps * tanh_fast(acc[k]), 0.0f);
}

// Q value of node n (gnn_agent.py:249-257)
template <int KP>
__device__ __forceinline__ float q_at_node(const float* __restrict__ ms, const float* __restrict__ c1, const GraphView& gv,
                                           const unsigned* __restrict__ feat, int n, float eps) {
  float x1n[KP], z[KP], x1s[KP];
  x1_at<KP>(ms, c1, gv, feat, n, eps, x1n);
  const float sc = __ldg(gv.self_coef + n);
#pragma unroll
  for (int k = 0; k < KP; ++k) z[k] = sc * x1n[k];
  const int e0 = __ldg(gv.in_ptr + n), e1 = __ldg(gv.in_ptr + n + 1);
  for (int e = e0; e < e1; ++e) {
    const float c = __ldg(gv.in_coef + e);
    x1_at<KP>(ms, c1, gv, feat, __ldg(gv.in_src + e), eps, x1s);
#pragma unroll
    for (int k = 0; k < KP; ++k) z[k] = fmaf(c, x1s[k], z[k]);
  }
  const float* ms2 = ms + Lay<KP>::CONV;
  float acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) acc[k] = ms2[Lay<KP>::BIAS + k];
  matvec_acc<KP>(acc, x1n, ms2 + Lay<KP>::WAS);
  matvec_acc<KP>(acc, z, ms2 + Lay<KP>::TH);
  float q = ms[Lay<KP>::OUT_B];
#pragma unroll
  for (int k = 0; k < KP; ++k) q = fmaf(fmaxf(x1n[k] + eps * tanh_fast(acc[k]), 0.0f), ms[Lay<KP>::OUT_W + k], q);
  return q;
}

// static shared memory: both models + their c1 rows, per-warp candidate bookkeeping; dynamic: per-warp feat[N] bitmasks
template <int KP>
struct GnnSmem {
  float model[2][gnn_model_floats(KP)];
  float c1[2][KP];
  int cstart[GNN_WARPS][SY_POLICY_MAX_FEATURES + 1];
  int apos[GNN_WARPS][SY_POLICY_MAX_FEATURES];
  int amoney[GNN_WARPS][SY_POLICY_MAX_FEATURES];
};

template <int KP>
__device__ __forceinline__ void gnn_load_models(const GnnParams& p, GnnSmem<KP>& sm, unsigned* feat_all, int n_feat) {
  for (int i = threadIdx.x; i < 2 * gnn_model_floats(KP); i += blockDim.x) (&sm.model[0][0])[i] = __ldg(p.params + i);
  for (int i = threadIdx.x; i < n_feat; i += blockDim.x) feat_all[i] = 0;
  __syncthreads();
  if (threadIdx.x < 2 * KP) {
    const int m = threadIdx.x / KP, k = threadIdx.x % KP;
    sm.c1[m][k] = fmaxf(0.0f + p.conv_eps * tanh_fast(sm.model[m][Lay<KP>::BIAS + k]), 0.0f);
  }
  __syncthreads();
}

// feature column k of env b sits at node ... (-2 = nowhere); see SY_FEATURES_*
__device__ __forceinline__ int gnn_feature_node(const GnnParams& p, int b, int k) {
  const int A = p.st.num_agents, N = p.g.num_nodes;
  if (k >= p.K) return -2;
  const int rev = p.st.mrx_revealed ? p.st.mrx_revealed[b] : 0;
  if (p.feature_mode == SY_FEATURES_ENV) return (k < A && !(k == 0 && rev < 0)) ? p.st.pos[(size_t)b * A + k] : -2;
  if (k == 0) return rev < 0 ? N - 1 : p.st.pos[(size_t)b * A];  // utils.py:176-199: numpy index -1 while MrX is hidden
  return k < A - 1 ? p.st.pos[(size_t)b * A + 1] : -2;           // range(number_of_agents - 1), always Polices_pos[0]
}

// set (on = true) or clear this env's feature bits in the warp's feat[]; lane k < KP owns column k (several columns
// may share a node, hence the shared-memory atomicOr)
template <int KP>
__device__ __forceinline__ void gnn_mark_features(unsigned* feat, int my_node, int lane, bool on) {
  __syncwarp();
  if (my_node >= 0) {
    if (on) atomicOr(feat + my_node, 1u << lane);
    else feat[my_node] = 0u;
  }
  __syncwarp();
}

// layer 2 + read-out of node n from the layer-1 rows of the whole graph (rows [N][KP] in shared memory)
template <int KP>
__device__ __forceinline__ float q_from_rows(const float* __restrict__ ms, const GraphView& gv, const float* __restrict__ rows, int n,
                                             float eps) {
  float x1n[KP], z[KP];
  const float sc = __ldg(gv.self_coef + n);
#pragma unroll
  for (int c4 = 0; c4 < KP / 4; ++c4) {
    const float4 r = reinterpret_cast<const float4*>(rows + n * KP)[c4];
    x1n[4 * c4] = r.x; x1n[4 * c4 + 1] = r.y; x1n[4 * c4 + 2] = r.z; x1n[4 * c4 + 3] = r.w;
  }
#pragma unroll
  for (int k = 0; k < KP; ++k) z[k] = sc * x1n[k];
  const int e0 = __ldg(gv.in_ptr + n), e1 = __ldg(gv.in_ptr + n + 1);
  for (int e = e0; e < e1; ++e) {
    const float c = __ldg(gv.in_coef + e);
    const float* rs = rows + __ldg(gv.in_src + e) * KP;
#pragma unroll
    for (int c4 = 0; c4 < KP / 4; ++c4) {
      const float4 r = reinterpret_cast<const float4*>(rs)[c4];
      z[4 * c4] = fmaf(c, r.x, z[4 * c4]);
      z[4 * c4 + 1] = fmaf(c, r.y, z[4 * c4 + 1]);
      z[4 * c4 + 2] = fmaf(c, r.z, z[4 * c4 + 2]);
      z[4 * c4 + 3] = fmaf(c, r.w, z[4 * c4 + 3]);
    }
  }
  const float* ms2 = ms + Lay<KP>::CONV;
  float acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) acc[k] = ms2[Lay<KP>::BIAS + k];
  matvec_acc<KP>(acc, x1n, ms2 + Lay<KP>::WAS);
  matvec_acc<KP>(acc, z, ms2 + Lay<KP>::TH);
  float q = ms[Lay<KP>::OUT_B];
#pragma unroll
  for (int k = 0; k < KP; ++k) q = fmaf(fmaxf(x1n[k] + eps * tanh_fast(acc[k]), 0.0f), ms[Lay<KP>::OUT_W + k], q);
  return q;
}

// Dense Q [B, 2, N] (training targets).  With room for one [N][KP] row block per warp the two layers run as two passes
// over the nodes (every layer-1 row computed once, same arithmetic and order as q_at_node); otherwise every node
// recomputes its neighbourhood (rows_per_warp_floats == 0).
template <int KP>
__global__ void __launch_bounds__(GNN_WARPS * 32) sy_gnn_q_dense_kernel(const GnnParams p, float* __restrict__ q, int rows_per_warp_floats) {
  __shared__ __align__(16) GnnSmem<KP> sm;
  extern __shared__ __align__(16) unsigned feat_all[];
  const int N = p.g.num_nodes, NF = (N + 7) & ~7;
  gnn_load_models<KP>(p, sm, feat_all, GNN_WARPS * NF);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned* feat = feat_all + w * NF;
  float* rows = reinterpret_cast<float*>(feat_all + GNN_WARPS * NF) + (size_t)w * rows_per_warp_floats;
  for (int b = blockIdx.x * GNN_WARPS + w; b < p.st.num_envs; b += gridDim.x * GNN_WARPS) {
    const int my_node = lane < KP ? gnn_feature_node(p, b, lane) : -2;
    gnn_mark_features<KP>(feat, my_node, lane, true);
    const int g = p.st.graph_id[b];
    const GraphView gv{p.g.in_ptr + (size_t)g * (N + 1), p.g.in_src + (size_t)g * p.g.in_stride, p.g.in_coef + (size_t)g * p.g.in_stride,
                       p.g.self_coef + (size_t)g * N};
    for (int m = 0; m < 2; ++m) {
      if (rows_per_warp_floats) {
        for (int n = lane; n < N; n += 32) {
          float x1[KP];
          x1_at<KP>(sm.model[m], sm.c1[m], gv, feat, n, p.conv_eps, x1);
#pragma unroll
          for (int c4 = 0; c4 < KP / 4; ++c4)
            reinterpret_cast<float4*>(rows + n * KP)[c4] = make_float4(x1[4 * c4], x1[4 * c4 + 1], x1[4 * c4 + 2], x1[4 * c4 + 3]);
        }
        __syncwarp();
        for (int n = lane; n < N; n += 32) q[((size_t)b * 2 + m) * N + n] = q_from_rows<KP>(sm.model[m], gv, rows, n, p.conv_eps);
        __syncwarp();
      } else {
        for (int n = lane; n < N; n += 32) q[((size_t)b * 2 + m) * N + n] = q_at_node<KP>(sm.model[m], sm.c1[m], gv, feat, n, p.conv_eps);
      }
    }
    gnn_mark_features<KP>(feat, my_node, lane, false);
  }
}

__device__ __forceinline__ unsigned lane_range_mask(int lo, int hi) {  // bits [lo, hi), 0 <= lo < 32, hi <= 32
  return hi <= lo ? 0u : ((hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u));
}

__device__ __forceinline__ unsigned ordered_key(float f) {  // monotone float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int KP>
__global__ void __launch_bounds__(GNN_WARPS * 32, 3) sy_gnn_act_kernel(const GnnParams p, int64_t* __restrict__ actions, float* __restrict__ q_taken) {
  __shared__ __align__(16) GnnSmem<KP> sm;
  extern __shared__ __align__(16) unsigned feat_all[];
  const int N = p.g.num_nodes, A = p.st.num_agents, NF = (N + 7) & ~7;
  gnn_load_models<KP>(p, sm, feat_all, GNN_WARPS * NF);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned* feat = feat_all + w * NF;
  // persistent: the models are loaded once per CTA, every warp walks its share of the envs
  for (int b = blockIdx.x * GNN_WARPS + w; b < p.st.num_envs; b += gridDim.x * GNN_WARPS) {
    const int my_node = lane < KP ? gnn_feature_node(p, b, lane) : -2;
    gnn_mark_features<KP>(feat, my_node, lane, true);
    const int g = p.st.graph_id[b];
    const int32_t* rp = p.g.row_ptr + (size_t)g * (N + 1);
    const int32_t* col = p.g.col + (size_t)g * p.g.nnz_stride;
    const int32_t* wt = p.g.w + (size_t)g * p.g.nnz_stride;
    const GraphView gv{p.g.in_ptr + (size_t)g * (N + 1), p.g.in_src + (size_t)g * p.g.in_stride, p.g.in_coef + (size_t)g * p.g.in_stride,
                       p.g.self_coef + (size_t)g * N};
    // lane a < A owns agent a: its node, budget, neighbour range, number of valid moves, explore decision
    int my_money = 0, r0 = 0, deg = 0, n_valid = 0;
    if (lane < A) {
      const int my_pos = p.st.pos[(size_t)b * A + lane];
      my_money = p.st.money[(size_t)b * A + lane];
      r0 = __ldg(rp + my_pos);
      deg = __ldg(rp + my_pos + 1) - r0;
      for (int k = 0; k < deg; ++k) n_valid += (__ldg(wt + r0 + k) + p.st.toll <= my_money);
      sm.apos[w][lane] = r0;
      sm.amoney[w][lane] = my_money;
    }
    int incl = deg;  // inclusive prefix of the degrees over lanes -> candidate ranges
#pragma unroll
    for (int o = 1; o < SY_POLICY_MAX_FEATURES; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane < A) sm.cstart[w][lane + 1] = incl;
    if (lane == 0) sm.cstart[w][0] = 0;
    const int C = __shfl_sync(FULL, incl, A - 1);
    bool explore = false;
    int target = 0;
    if (lane < A) {
      const uint4 r = philox4x32(make_uint4((unsigned)(p.st.env_offset + b), p.step, RNG_GNN_POLICY, (unsigned)lane), make_uint2(p.seed_lo, p.seed_hi));
      explore = u01(r.x) < (lane == 0 ? p.eps_mrx : p.eps_police);  // np.random.rand() <= epsilon (gnn_agent.py:69)
      target = (int)__umulhi(r.y, (unsigned)n_valid);                // np.random.choice(valid_actions)
    }
    const unsigned explore_mask = __ballot_sync(FULL, explore);
    __syncwarp();
    int seen = 0;        // lane a: valid candidates of agent a in earlier chunks
    int best_node = -1;  // lane a: running choice
    unsigned best_key = 0;
    float best_q = CUDART_NAN_F;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      int a = -1, node = -1;
      bool valid = false;
      if (c < C) {
        a = 0;
        while (c >= sm.cstart[w][a + 1]) ++a;
        const int k = sm.apos[w][a] + (c - sm.cstart[w][a]);
        node = __ldg(col + k);
        valid = __ldg(wt + k) + p.st.toll <= sm.amoney[w][a];
      }
      float qv = 0.0f;
      if (valid && !((explore_mask >> a) & 1u)) {
        const int mi = a == 0 ? 0 : 1;  // MrX's agent / the police agent
        qv = q_at_node<KP>(&sm.model[0][0] + mi * Lay<KP>::MF, &sm.c1[0][0] + mi * KP, gv, feat, node, p.conv_eps);
      }
      const unsigned key = valid ? ordered_key(qv) : 0u;  // every valid key is > 0
      __syncwarp();
      // per-agent choice without a loop over the agents: the candidates of one agent are contiguous lanes, so a
      // segmented shuffle reduction leaves each segment's first maximum (np.argmax, gnn_agent.py:75) in its first lane,
      // and one ballot marks every exploring agent's target-th valid candidate; lane a (< A) then collects its own.
      const unsigned vb = __ballot_sync(FULL, valid);
      const int seg_hi = a >= 0 ? min(sm.cstart[w][a + 1] - c0, 32) : 0;
      const int seg_lo = a >= 0 ? max(sm.cstart[w][a] - c0, 0) : 0;
      unsigned mykey = key;
      int mybest = lane;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned k2 = __shfl_down_sync(FULL, mykey, o);
        const int b2 = __shfl_down_sync(FULL, mybest, o);
        if (lane + o < seg_hi && k2 > mykey) {
          mykey = k2;
          mybest = b2;
        }
      }
      const int seen_a = __shfl_sync(FULL, seen, max(a, 0)), target_a = __shfl_sync(FULL, target, max(a, 0));
      const unsigned seg_mask = lane_range_mask(seg_lo, seg_hi);
      const int rank = seen_a + __popc(vb & seg_mask & ((1u << lane) - 1u));
      const unsigned hb = __ballot_sync(FULL, valid && ((explore_mask >> a) & 1u) && rank == target_a);
      int s0 = 32, s1 = 0;
      if (lane < A) {
        s0 = sm.cstart[w][lane] - c0;
        s1 = sm.cstart[w][lane + 1] - c0;
      }
      const bool mine = s0 < 32 && s1 > 0 && s1 > s0;  // agent `lane` has candidates in this chunk
      const unsigned own = mine ? lane_range_mask(max(s0, 0), min(s1, 32)) : 0u;
      const int head = mine ? max(s0, 0) : 0;
      const unsigned kmax = __shfl_sync(FULL, mykey, head);
      const int exploit_lane = __shfl_sync(FULL, mybest, head);
      int pick_lane = -1;
      if (explore) {
        if (hb & own) pick_lane = __ffs(hb & own) - 1;
      } else if (vb & own) {
        pick_lane = exploit_lane;
      }
      const int pn = __shfl_sync(FULL, node, max(pick_lane, 0));
      const float pq = __shfl_sync(FULL, qv, max(pick_lane, 0));
      if (mine) {
        if (pick_lane >= 0 && (explore || best_node < 0 || kmax > best_key)) {
          best_node = pn;
          best_key = kmax;
          best_q = explore ? CUDART_NAN_F : pq;
        }
        seen += __popc(vb & own);
      }
    }
    if (lane < A) {
      actions[(size_t)b * A + lane] = best_node;  // -1: no valid move (DEFAULT_ACTION)
      if (q_taken) q_taken[(size_t)b * A + lane] = best_q;
    }
    gnn_mark_features<KP>(feat, my_node, lane, false);
  }
}

// ---------------------------------------------------------------------------------------------
// MAPPO
// ---------------------------------------------------------------------------------------------
constexpr int MP_ROWS = 64;      // envs per CTA (69 KB of shared memory at N = 200, H = 64: three CTAs per SM)
constexpr int MP_THREADS = 256;
constexpr int MP_HC = 32;        // hidden units per register chunk
constexpr int MP_RU = 4;         // rows in flight per thread in the logits loop

struct MappoParams {
  SyPolicyGraphs g;
  SyPolicyState st;
  const float* obs;
  const float* params;
  int obs_size, H, HP, policy_floats;
  int policy_of_agent[SY_POLICY_MAX_FEATURES];
  unsigned seed_lo, seed_hi, step;
};


// MappoTrainer's observation of an agent built from the env state (mappo_trainer.py:171-199) when the caller passes
// obs == NULL: MrX sees his own node (obs["MrX_pos"], one float), police officer i the nodes of all officers
// (obs["Polices_pos"].sum(dim=1), P floats) -- raw node ids as float32, zero-padded to obs_size columns.
__device__ __forceinline__ float mappo_obs_value(const MappoParams& p, int b, int a, int o) {
  if (p.obs) return __ldg(p.obs + ((size_t)b * p.st.num_agents + a) * p.obs_size + o);
  const int A = p.st.num_agents;
  if (a == 0) return o == 0 ? (float)__ldg(p.st.pos + (size_t)b * A) : 0.0f;
  return o < A - 1 ? (float)__ldg(p.st.pos + (size_t)b * A + 1 + o) : 0.0f;
}

__global__ void __launch_bounds__(MP_THREADS, 3) sy_mappo_act_kernel(const MappoParams p, int64_t* __restrict__ actions,
                                                                      float* __restrict__ log_probs, float* __restrict__ probs_out) {
  extern __shared__ __align__(16) float mp_smem[];
  const int N = p.g.num_nodes, A = p.st.num_agents, H = p.H, HP = p.HP, D = p.obs_size;
  float* hid = mp_smem;                    // [MP_ROWS][HP]
  float* logits = hid + MP_ROWS * HP;      // [MP_ROWS][N]
  float* obs_s = logits + MP_ROWS * N;     // [MP_ROWS][D]
  const int a = blockIdx.y, row0 = blockIdx.x * MP_ROWS;
  const int nrows = min(MP_ROWS, p.st.num_envs - row0);
  const float* pol = p.params + (size_t)p.policy_of_agent[a] * p.policy_floats;
  const float* W1 = pol;                     // [H, D]
  const float* b1 = W1 + (size_t)H * D;      // [H]
  const float* W2 = b1 + H;                  // [N, H]
  const float* b2 = W2 + (size_t)N * H;      // [N]
  const int tid = threadIdx.x;
  for (int i = tid; i < MP_ROWS * D; i += MP_THREADS) {
    const int r = i / D, o = i - r * D;
    obs_s[i] = r < nrows ? mappo_obs_value(p, row0 + r, a, o) : 0.0f;
  }
  __syncthreads();
  // hidden layer: relu(W1 obs + b1); with HP dividing the block every thread keeps one hidden unit
  if (MP_THREADS % HP == 0) {
    const int h = tid % HP;
    const float bias = h < H ? __ldg(b1 + h) : 0.0f;
    for (int r = tid / HP; r < MP_ROWS; r += MP_THREADS / HP) {
      float v = bias;
      if (h < H)
        for (int o = 0; o < D; ++o) v = fmaf(__ldg(W1 + (size_t)h * D + o), obs_s[r * D + o], v);
      hid[r * HP + h] = h < H ? fmaxf(v, 0.0f) : 0.0f;
    }
  } else {
    for (int i = tid; i < MP_ROWS * HP; i += MP_THREADS) {
      const int r = i / HP, h = i - r * HP;
      float v = 0.0f;
      if (h < H) {
        v = __ldg(b1 + h);
        for (int o = 0; o < D; ++o) v = fmaf(__ldg(W1 + (size_t)h * D + o), obs_s[r * D + o], v);
        v = fmaxf(v, 0.0f);
      }
      hid[i] = v;
    }
  }
  __syncthreads();
  // logits: thread = action node, W2 row chunk in registers, MP_RU rows of the tile in flight
  for (int n = tid; n < N; n += MP_THREADS) {
    const float bias = __ldg(b2 + n);
    for (int hc = 0; hc < HP; hc += MP_HC) {
      float wreg[MP_HC];
#pragma unroll
      for (int j = 0; j < MP_HC; ++j) wreg[j] = (hc + j < H) ? __ldg(W2 + (size_t)n * H + hc + j) : 0.0f;
      for (int r = 0; r < MP_ROWS; r += MP_RU) {
        float acc[MP_RU];
#pragma unroll
        for (int u = 0; u < MP_RU; ++u) acc[u] = hc == 0 ? bias : logits[(r + u) * N + n];
#pragma unroll
        for (int j4 = 0; j4 < MP_HC / 4; ++j4) {
#pragma unroll
          for (int u = 0; u < MP_RU; ++u) {
            const float4 h4 = reinterpret_cast<const float4*>(hid + (r + u) * HP + hc)[j4];
            acc[u] = fmaf(wreg[4 * j4 + 0], h4.x, acc[u]);
            acc[u] = fmaf(wreg[4 * j4 + 1], h4.y, acc[u]);
            acc[u] = fmaf(wreg[4 * j4 + 2], h4.z, acc[u]);
            acc[u] = fmaf(wreg[4 * j4 + 3], h4.w, acc[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < MP_RU; ++u) logits[(r + u) * N + n] = acc[u];
      }
    }
  }
  __syncthreads();
  // softmax, mask, renormalise, sample, log-prob: warp per row, lane = neighbour slot (mappo_agent.py:102-142)
  const int lane = tid & 31, w = tid >> 5;
  for (int r = w; r < nrows; r += MP_THREADS / 32) {
    const int b = row0 + r;
    float* lr = logits + r * N;
    float mx = -CUDART_INF_F;
    for (int n = lane; n < N; n += 32) mx = fmaxf(mx, lr[n]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    float Z = 0.0f;
    for (int n = lane; n < N; n += 32) {
      const float e = expf(lr[n] - mx);
      lr[n] = e;
      Z += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(FULL, Z, o);
    __syncwarp();
    const int g = p.st.graph_id[b];
    const int pos = p.st.pos[(size_t)b * A + a], money = p.st.money[(size_t)b * A + a];
    const int32_t* rp = p.g.row_ptr + (size_t)g * (N + 1);
    const int r0 = __ldg(rp + pos), deg = __ldg(rp + pos + 1) - r0;
    const int32_t* col = p.g.col + (size_t)g * p.g.nnz_stride + r0;
    const int32_t* wt = p.g.w + (size_t)g * p.g.nnz_stride + r0;
    // masked softmax mass over the valid moves
    float Sv = 0.0f;
    int nv = 0;
    for (int k = lane; k < deg; k += 32) {
      if (__ldg(wt + k) + p.st.toll <= money) {
        Sv += lr[__ldg(col + k)] / Z;
        nv += 1;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Sv += __shfl_xor_sync(FULL, Sv, o);
      nv += __shfl_xor_sync(FULL, nv, o);
    }
    // mode 0: softmax * mask / (sum + 1e-8); 1: uniform over the mask; 2: uniform over all nodes
    const int mode = Sv > 1e-8f ? 0 : (nv > 0 ? 1 : 2);
    const uint4 rnd = philox4x32(make_uint4((unsigned)(p.st.env_offset + b), p.step, RNG_MAPPO_POLICY, (unsigned)a), make_uint2(p.seed_lo, p.seed_hi));
    const float u = u01(rnd.x);
    int action = -1;
    float pa = 0.0f, total = 1.0f;
    if (mode == 2) {
      action = min((int)(u * (float)N), N - 1);
      pa = 1.0f / (float)N;  // total: N * (1/N) up to rounding; Categorical renormalises
      if (probs_out)
        for (int n = lane; n < N; n += 32) probs_out[((size_t)b * A + a) * N + n] = pa;
    } else {
      if (probs_out) {
        for (int n = lane; n < N; n += 32) probs_out[((size_t)b * A + a) * N + n] = 0.0f;
        __syncwarp();
      }
      // current_probs over the valid moves (ascending); Categorical's own normalisation
      const float inv_nv = 1.0f / (float)max(nv, 1), den = Sv + 1e-8f;
      total = 0.0f;
      for (int k = lane; k < deg; k += 32) {
        if (__ldg(wt + k) + p.st.toll <= money) {
          const int node = __ldg(col + k);
          const float pk = mode == 0 ? (lr[node] / Z) / den : inv_nv;
          total += pk;
          if (probs_out) probs_out[((size_t)b * A + a) * N + node] = pk;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
      // inverse CDF in ascending node order: each lane holds one slot's probability, the walk is a shuffle loop
      const float thr = u * total;
      float cum = 0.0f, plast = 0.0f;
      int last = -1;
      for (int c0 = 0; c0 < deg; c0 += 32) {
        const int k = c0 + lane;
        int node = -1;
        float pk = 0.0f;
        if (k < deg && __ldg(wt + k) + p.st.toll <= money) {
          node = __ldg(col + k);
          pk = mode == 0 ? (lr[node] / Z) / den : inv_nv;
        }
        const unsigned vm = __ballot_sync(FULL, node >= 0);
        for (unsigned bits = vm; bits; bits &= bits - 1) {
          const int j = __ffs(bits) - 1;
          const float pj = __shfl_sync(FULL, pk, j);
          cum += pj;
          last = j + c0;
          plast = pj;
          if (action < 0 && thr < cum) {
            action = __shfl_sync(FULL, node, j);
            pa = pj;
          }
        }
      }
      if (action < 0) {  // rounding at the top of the CDF
        action = __ldg(col + last);
        pa = plast;
      }
    }
    if (lane == 0) {
      actions[(size_t)b * A + a] = action;
      // Categorical(probs).log_prob: log(clamp(p / sum p, eps, 1 - eps)) (torch.distributions.utils.probs_to_logits)
      const float eps = 1.1920928955078125e-07f;
      log_probs[(size_t)b * A + a] = logf(fminf(fmaxf(pa / total, eps), 1.0f - eps));
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// MAPPO on the tensor cores (tcgen05 + TMEM): the logits GEMM [128 envs x H] x [H x N] of one agent's policy as
// 3xTF32 (A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32-level accuracy) with the accumulator tile in tensor memory.
// One persistent CTA per SM slice of (agent, row tiles): W2 is split and laid out once, every tile then costs 24 MMA
// instructions issued by one thread; the epilogue reads TMEM with thread = row (lane = TMEM lane), so the softmax
// statistics, the masked mass, the inverse-CDF sample and the log-prob need no cross-lane traffic at all.
// Operands sit in shared memory in the canonical K-major no-swizzle layout of the UMMA descriptors: 16-byte chunks of
// 4 consecutive k for one row, rows contiguous (8 rows = one 128-byte core matrix), K-chunk panels LBO apart.
// ---------------------------------------------------------------------------------------------
constexpr int TC_ROWS = 128;
constexpr int TC_THREADS = 256;
constexpr int TC_MAXV = 16;  // valid moves per agent the epilogue keeps per column half (host checks the pool's max degree)
constexpr int TC_DMAX = 16;  // observation size the hidden-layer registers hold

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ uint64_t umma_desc_k_major(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); base offset 0, SWIZZLE_NONE
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcShape {
  int Kp, Np;                       // hidden padded to 8, nodes padded to 16
  int off_alo, off_bhi, off_blo;    // float offsets inside the dynamic shared memory (A_hi at 0)
  int off_obs, off_b2, off_vals, off_nodes, off_mask, off_red, off_w1, total_bytes;
};

__host__ __device__ inline TcShape tc_shape(int H, int N, int D) {
  TcShape t;
  t.Kp = (H + 7) & ~7;
  t.Np = (N + 15) & ~15;
  int o = TC_ROWS * t.Kp;
  t.off_alo = o; o += TC_ROWS * t.Kp;
  t.off_bhi = o; o += t.Np * t.Kp;
  t.off_blo = o; o += t.Np * t.Kp;
  t.off_obs = o; o += 2 * TC_ROWS * D;              // two tiles in flight
  t.off_b2 = o; o += t.Np;
  t.off_vals = o; o += TC_ROWS * 2 * TC_MAXV;
  t.off_nodes = o; o += TC_ROWS * 2 * TC_MAXV / 2;  // u16
  t.off_mask = o; o += 2 * TC_ROWS * 8;             // 256-bit valid-move mask per row, two tiles in flight
  t.off_red = o; o += 2 * TC_ROWS * 4;              // per (column half, row): max, Z, #valid, spare
  t.off_w1 = o; o += t.Kp * (D + 1);                // W1 rows + b1, padded hidden units zero
  t.total_bytes = o * 4 + 64;
  return t;
}

__global__ void __launch_bounds__(TC_THREADS, 1) sy_mappo_act_tc_kernel(const MappoParams p, int64_t* __restrict__ actions,
                                                                         float* __restrict__ log_probs, float* __restrict__ probs_out,
                                                                         int* __restrict__ error_flag) {
  extern __shared__ __align__(1024) float tc_smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar[2];  // MMA completion, one per accumulator buffer
  const int N = p.g.num_nodes, A = p.st.num_agents, H = p.H, D = p.obs_size;
  const TcShape ts = tc_shape(H, N, D);
  const int Kp = ts.Kp, Np = ts.Np;
  float* a_hi = tc_smem;
  float* a_lo = tc_smem + ts.off_alo;
  float* b_hi = tc_smem + ts.off_bhi;
  float* b_lo = tc_smem + ts.off_blo;
  float* obs_all = tc_smem + ts.off_obs;
  float* b2_s = tc_smem + ts.off_b2;
  float* vals = tc_smem + ts.off_vals;
  uint16_t* vnodes = reinterpret_cast<uint16_t*>(tc_smem + ts.off_nodes);
  unsigned* vmask_all = reinterpret_cast<unsigned*>(tc_smem + ts.off_mask);
  float* red = tc_smem + ts.off_red;
  float* w1_s = tc_smem + ts.off_w1;  // [Kp][D + 1]: W1 row then b1
  const int tid = threadIdx.x, warp = tid >> 5, a = blockIdx.y;
  const float* pol = p.params + (size_t)p.policy_of_agent[a] * p.policy_floats;
  const float* W1 = pol;
  const float* b1 = W1 + (size_t)H * D;
  const float* W2 = b1 + H;
  const float* b2 = W2 + (size_t)N * H;

  // ---- one-time setup: TMEM columns, the MMA-completion barrier, W2 split into tf32 hi / lo in the UMMA layout
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&mbar[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&mbar[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = tid; i < Np * (Kp / 4); i += TC_THREADS) {
    const int kc = i / Np, n = i - kc * Np;
    float w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc * 4 + j;
      w[j] = (n < N && k < H) ? __ldg(W2 + (size_t)n * H + k) : 0.0f;
    }
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(w[0]) & 0xFFFFE000u); lo.x = w[0] - hi.x;
    hi.y = __uint_as_float(__float_as_uint(w[1]) & 0xFFFFE000u); lo.y = w[1] - hi.y;
    hi.z = __uint_as_float(__float_as_uint(w[2]) & 0xFFFFE000u); lo.z = w[2] - hi.z;
    hi.w = __uint_as_float(__float_as_uint(w[3]) & 0xFFFFE000u); lo.w = w[3] - hi.w;
    reinterpret_cast<float4*>(b_hi)[i] = hi;  // float4 index (k / 4) * Np + n
    reinterpret_cast<float4*>(b_lo)[i] = lo;
  }
  // bias pre-scaled for the exp2 of the epilogue; the padded columns get -inf-like so they drop out of max and sum
  for (int n = tid; n < Np; n += TC_THREADS) b2_s[n] = n < N ? __ldg(b2 + n) : -1e30f;
  for (int i = tid; i < Kp * (D + 1); i += TC_THREADS) {
    const int h = i / (D + 1), o = i - h * (D + 1);
    w1_s[i] = h < H ? (o < D ? __ldg(W1 + (size_t)h * D + o) : __ldg(b1 + h)) : 0.0f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Np >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
  const uint32_t lbo_a = TC_ROWS * 16, lbo_b = (uint32_t)Np * 16, sbo = 128;
  // hidden layer: this thread's four hidden units (one 16-byte K chunk) stay in registers for the whole kernel
  const int my_kc = tid >> 4;  // Kp / 4 <= 16 chunks; threads 16 apart in a chunk take rows 16 apart
  float w1r[4][TC_DMAX], b1r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int h = min(my_kc * 4 + j, Kp - 1);
    const bool live = my_kc * 4 + j < Kp;
    b1r[j] = live ? w1_s[h * (D + 1) + D] : 0.0f;
#pragma unroll
    for (int o = 0; o < TC_DMAX; ++o) w1r[j][o] = (live && o < D) ? w1_s[h * (D + 1) + o] : 0.0f;
  }
  // epilogue split: warps w and w + 4 read the same TMEM lanes (rows), each one half of the columns
  const int half = warp >> 2, erow = tid & 127;
  const int ncb = Np / 16, cb0 = half ? (ncb + 1) / 2 : 0, cb1 = half ? ncb : (ncb + 1) / 2;

  const int ntiles = (p.st.num_envs + TC_ROWS - 1) / TC_ROWS;

  // wait for the MMAs of accumulator buffer `buf` (bounded: a mistake must not hang the GPU)
  auto wait_mma = [&](int buf, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; ++spin)
      asm volatile(
          "{\n\t.reg .pred P1;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, P1;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar[buf])), "r"(parity)
          : "memory");
    if (!done && error_flag) atomicExch(error_flag, 1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  };

  // front half of a tile: obs + valid-move masks + hidden layer into the A operand, then its MMAs are issued into
  // accumulator buffer `buf` (256 TMEM columns each); they run while the previous tile's epilogue executes
  auto front = [&](int tile, int buf) {
    const int row0 = tile * TC_ROWS;
    const int nrows = min(TC_ROWS, p.st.num_envs - row0);
    float* obs_s = obs_all + buf * TC_ROWS * D;
    unsigned* vmask = vmask_all + buf * TC_ROWS * 8;
    for (int i = tid; i < TC_ROWS * D; i += TC_THREADS) {
      const int r = i / D, o = i - r * D;
      obs_s[i] = r < nrows ? mappo_obs_value(p, row0 + r, a, o) : 0.0f;
    }
    for (int i = tid; i < TC_ROWS * 8; i += TC_THREADS) vmask[i] = 0u;
    __syncthreads();
    if (tid >= 128 && tid - 128 < nrows) {  // the upper warps build the masks while the lower ones start on the hidden layer
      const int r = tid - 128, b = row0 + r;
      const int g = p.st.graph_id[b];
      const int pos = p.st.pos[(size_t)b * A + a], money = p.st.money[(size_t)b * A + a];
      const int32_t* rp = p.g.row_ptr + (size_t)g * (N + 1);
      const int r0 = __ldg(rp + pos), deg = __ldg(rp + pos + 1) - r0;
      const int32_t* col = p.g.col + (size_t)g * p.g.nnz_stride + r0;
      const int32_t* wt = p.g.w + (size_t)g * p.g.nnz_stride + r0;
      for (int k = 0; k < deg; ++k)
        if (__ldg(wt + k) + p.st.toll <= money) {
          const int n = __ldg(col + k);
          vmask[r * 8 + (n >> 5)] |= 1u << (n & 31);
        }
    }
    if (my_kc < Kp / 4) {  // hidden layer relu(W1 obs + b1) split into tf32 hi / lo, written in the UMMA layout
      for (int r = tid & 15; r < TC_ROWS; r += 16) {
        float v[4] = {b1r[0], b1r[1], b1r[2], b1r[3]};
#pragma unroll
        for (int o = 0; o < TC_DMAX; ++o) {
          if (o < D) {
            const float x = obs_s[r * D + o];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = fmaf(w1r[j][o], x, v[j]);
          }
        }
        float4 hi, lo;
        v[0] = fmaxf(v[0], 0.0f); v[1] = fmaxf(v[1], 0.0f); v[2] = fmaxf(v[2], 0.0f); v[3] = fmaxf(v[3], 0.0f);
        hi.x = __uint_as_float(__float_as_uint(v[0]) & 0xFFFFE000u); lo.x = v[0] - hi.x;
        hi.y = __uint_as_float(__float_as_uint(v[1]) & 0xFFFFE000u); lo.y = v[1] - hi.y;
        hi.z = __uint_as_float(__float_as_uint(v[2]) & 0xFFFFE000u); lo.z = v[2] - hi.z;
        hi.w = __uint_as_float(__float_as_uint(v[3]) & 0xFFFFE000u); lo.w = v[3] - hi.w;
        reinterpret_cast<float4*>(a_hi)[my_kc * TC_ROWS + r] = hi;  // float4 index (k / 4) * 128 + row
        reinterpret_cast<float4*>(a_lo)[my_kc * TC_ROWS + r] = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (tid == 0) {  // one thread issues the tile's MMAs; completion arrives on the buffer's mbarrier
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
      const uint32_t td = tmem_base + (uint32_t)buf * 256u;
      for (int ks = 0; ks < Kp / 8; ++ks) {  // one MMA covers K = 8 = two 16-byte chunks, LBO apart
        const uint64_t dah = umma_desc_k_major(ah + ks * 2 * lbo_a, lbo_a, sbo), dal = umma_desc_k_major(al + ks * 2 * lbo_a, lbo_a, sbo);
        const uint64_t dbh = umma_desc_k_major(bh + ks * 2 * lbo_b, lbo_b, sbo), dbl = umma_desc_k_major(bl + ks * 2 * lbo_b, lbo_b, sbo);
        umma_tf32(td, dah, dbh, idesc, ks > 0 ? 1u : 0u);
        umma_tf32(td, dal, dbh, idesc, 1u);
        umma_tf32(td, dah, dbl, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar[buf])) : "memory");
    }
  };

  if ((int)blockIdx.x < ntiles) front(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int row0 = tile * TC_ROWS;
    const int nrows = min(TC_ROWS, p.st.num_envs - row0);
    const unsigned* vmask = vmask_all + buf * TC_ROWS * 8;
    wait_mma(buf, (uint32_t)(it >> 1) & 1u);  // this tile's accumulators are complete, the A operand is free again
    if (tile + (int)gridDim.x < ntiles) front(tile + gridDim.x, buf ^ 1);  // next tile's MMAs overlap this epilogue
    // ---- (4) epilogue: thread = (row = TMEM lane, column half)
    {
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * 256u;
      const int r = erow;
      float mx = -CUDART_INF_F;
      for (int cb = cb0; cb < cb1; ++cb) {
        float v[16];
        tmem_ld16(taddr + cb * 16, v);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 bb = reinterpret_cast<const float4*>(b2_s + cb * 16)[j4];
          mx = fmaxf(mx, fmaxf(fmaxf(v[4 * j4] + bb.x, v[4 * j4 + 1] + bb.y), fmaxf(v[4 * j4 + 2] + bb.z, v[4 * j4 + 3] + bb.w)));
        }
      }
      red[(half * TC_ROWS + r) * 4] = mx;
      __syncthreads();
      mx = fmaxf(red[r * 4], red[(TC_ROWS + r) * 4]);
      float Zp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      int nv = 0;
      float* myvals = vals + (r * 2 + half) * TC_MAXV;
      uint16_t* mynodes = vnodes + (r * 2 + half) * TC_MAXV;
      for (int cb = cb0; cb < cb1; ++cb) {
        float v[16];
        tmem_ld16(taddr + cb * 16, v);
        const int c0 = cb * 16;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {  // e = exp(logit - max), logit = accumulator + bias
          const float4 bb = reinterpret_cast<const float4*>(b2_s + c0)[j4];
          v[4 * j4 + 0] = __expf(v[4 * j4 + 0] + bb.x - mx);
          v[4 * j4 + 1] = __expf(v[4 * j4 + 1] + bb.y - mx);
          v[4 * j4 + 2] = __expf(v[4 * j4 + 2] + bb.z - mx);
          v[4 * j4 + 3] = __expf(v[4 * j4 + 3] + bb.w - mx);
          Zp[0] += v[4 * j4 + 0];
          Zp[1] += v[4 * j4 + 1];
          Zp[2] += v[4 * j4 + 2];
          Zp[3] += v[4 * j4 + 3];
        }
        const unsigned bits = (vmask[r * 8 + (c0 >> 5)] >> (c0 & 31)) & 0xFFFFu;
        // valid moves among these 16 columns (uncommon): keep their e and node, ascending.  One pass per set bit with
        // a select chain over the register tile instead of sixteen predicated tests per chunk.
        for (unsigned bb = bits; bb; bb &= bb - 1u) {
          const int j = __ffs(bb) - 1;
          float e = v[0];
#pragma unroll
          for (int q = 1; q < 16; ++q) e = (j == q) ? v[q] : e;
          if (nv < TC_MAXV) {
            myvals[nv] = e;
            mynodes[nv] = (uint16_t)(c0 + j);
          }
          ++nv;
        }
      }
      float Z = (Zp[0] + Zp[1]) + (Zp[2] + Zp[3]);
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      red[(half * TC_ROWS + r) * 4 + 1] = Z;
      red[(half * TC_ROWS + r) * 4 + 2] = __int_as_float(nv);
      __syncthreads();
      if (half == 0 && r < nrows) {  // the lower half's thread finishes the row: valid moves of both halves, ascending
        const int b = row0 + r;
        const int nv0 = nv, nv1 = __float_as_int(red[(TC_ROWS + r) * 4 + 2]);
        if ((nv0 > TC_MAXV || nv1 > TC_MAXV) && error_flag) atomicExch(error_flag, 2);
        const int n0 = min(nv0, TC_MAXV), n1 = min(nv1, TC_MAXV), nvt = n0 + n1;
        Z += red[(TC_ROWS + r) * 4 + 1];
        // entry k of the row's list: first the lower half's, then the upper half's (two adjacent TC_MAXV blocks)
        float* lv = vals + r * 2 * TC_MAXV;
        uint16_t* ln = vnodes + r * 2 * TC_MAXV;
#define ROW_SLOT(k) ((k) < n0 ? (k) : TC_MAXV + (k) - n0)
        const float rZ = 1.0f / Z;  // softmax = e / Z as e * (1 / Z): well inside the parity tolerance, a third of the divides
        float Sv = 0.0f;
        for (int k = 0; k < nvt; ++k) Sv += lv[ROW_SLOT(k)] * rZ;
        const int mode = Sv > 1e-8f ? 0 : (nvt > 0 ? 1 : 2);  // mappo_agent.py:121-133
        const uint4 rnd = philox4x32(make_uint4((unsigned)(p.st.env_offset + b), p.step, RNG_MAPPO_POLICY, (unsigned)a), make_uint2(p.seed_lo, p.seed_hi));
        const float u = u01(rnd.x);
        int action = -1;
        float pa = 0.0f, total = 1.0f;
        float* po = probs_out ? probs_out + ((size_t)b * A + a) * N : nullptr;
        if (mode == 2) {
          action = min((int)(u * (float)N), N - 1);
          pa = 1.0f / (float)N;
          if (po)
            for (int n = 0; n < N; ++n) po[n] = pa;
        } else {
          if (po)
            for (int n = 0; n < N; ++n) po[n] = 0.0f;
          const float inv_nv = 1.0f / (float)max(nvt, 1), scale = rZ / (Sv + 1e-8f);
          total = 0.0f;
          for (int k = 0; k < nvt; ++k) {
            const float pk = mode == 0 ? lv[ROW_SLOT(k)] * scale : inv_nv;
            lv[ROW_SLOT(k)] = pk;
            total += pk;
            if (po) po[ln[ROW_SLOT(k)]] = pk;
          }
          const float thr = u * total;
          float cum = 0.0f;
          for (int k = 0; k < nvt; ++k) {
            const float pk = lv[ROW_SLOT(k)];
            cum += pk;
            if (thr < cum) {
              action = ln[ROW_SLOT(k)];
              pa = pk;
              break;
            }
          }
          if (action < 0) {  // rounding at the top of the CDF
            action = ln[ROW_SLOT(nvt - 1)];
            pa = lv[ROW_SLOT(nvt - 1)];
          }
        }
#undef ROW_SLOT
        actions[(size_t)b * A + a] = action;
        const float eps = 1.1920928955078125e-07f;  // Categorical: log(clamp(p / sum p, eps, 1 - eps))
        log_probs[(size_t)b * A + a] = logf(fminf(fmaxf(pa / total, eps), 1.0f - eps));
      }
    }
    __syncthreads();  // this accumulator buffer and the vals / red scratch are free again
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(512));
}

// CentralCritic: thread per row
__global__ void sy_mappo_values_kernel(const float* __restrict__ x, int M, int D, int H, const float* __restrict__ params, float* __restrict__ out) {
  extern __shared__ float cw[];  // W1 [H, D], b1 [H], W2 [H], b2
  const int total = H * D + 2 * H + 1;
  for (int i = threadIdx.x; i < total; i += blockDim.x) cw[i] = __ldg(params + i);
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const float* W1 = cw;
  const float* b1 = cw + H * D;
  const float* W2 = b1 + H;
  float v = W2[H];
  const float* xr = x + (size_t)r * D;
  for (int h = 0; h < H; ++h) {
    float s = b1[h];
    for (int o = 0; o < D; ++o) s = fmaf(W1[h * D + o], __ldg(xr + o), s);
    v = fmaf(fmaxf(s, 0.0f), W2[h], v);
  }
  out[r] = v;
}

int check_common(const SyPolicyGraphs* g, const SyPolicyState* st) {
  if (!g || !st) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "NULL graphs / state");
  if (g->num_graphs <= 0 || g->num_nodes <= 1 || !g->row_ptr || !g->col || !g->w) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "bad graph pool");
  if (st->num_envs <= 0 || st->num_agents < 2 || st->num_agents > SY_POLICY_MAX_FEATURES) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "bad batch shape");
  if (!st->pos || !st->money || !st->graph_id) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "NULL state members");
  return SY_POLICY_OK;
}

int fill_gnn(const SyPolicyGraphs* g, const SyPolicyState* st, const float* params, int K, float conv_eps, int mode, GnnParams& p) {
  if (int rc = check_common(g, st)) return rc;
  if (!g->in_ptr || !g->in_src || !g->in_coef || !g->self_coef) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "NULL in-edge lists");
  if (!params || K < 1 || K > SY_POLICY_MAX_FEATURES) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "bad GNN parameters (K = %d)", K);
  if (mode != SY_FEATURES_ENV && mode != SY_FEATURES_REFERENCE) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "bad feature mode");
  std::memset(&p, 0, sizeof(p));
  p.g = *g;
  p.st = *st;
  p.params = params;
  p.K = K;
  p.feature_mode = mode;
  p.conv_eps = conv_eps;
  return SY_POLICY_OK;
}

// persistent grid (a few CTAs per SM, each warp walks its share of the envs) + the per-warp feature bitmasks
int gnn_launch_shape(const SyPolicyGraphs* g, const SyPolicyState* st, unsigned& grid, size_t& dyn) {
  dyn = (size_t)GNN_WARPS * ((g->num_nodes + 7) & ~7) * sizeof(unsigned);
  if (dyn > 160 * 1024) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "num_nodes too large for the GNN kernels' shared memory");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned need = (unsigned)((st->num_envs + GNN_WARPS - 1) / GNN_WARPS);
  grid = need < (unsigned)(sms * 6) ? need : (unsigned)(sms * 6);
  return SY_POLICY_OK;
}


// ---------------------------------------------------------------------------------------------
// sy_masked_sample: one node per (env, agent) from arbitrary policy logits restricted to the env's action mask -- the
// batched form of the trainers' "pick among the valid moves" (gnn_trainer.py:221-229, mappo_agent.py:121-140) for a
// policy that is NOT one of the two built-in agents (e.g. a torch module producing logits [B, A, N]).
// greedy: first argmax over the legal nodes.  Otherwise Gumbel-max: argmax_j (logit_j + g_j), g_j = -log(-log u_j),
// u_j from Philox(seed; env, step, RNG_MASKED_SAMPLE + 16 * agent, j / 4) -- an exact sample of softmax(logits | mask)
// without normalising or a second pass.  Rows without a legal node get -1 (DEFAULT_ACTION).  One warp per row.
// ---------------------------------------------------------------------------------------------
enum { RNG_MASKED_SAMPLE = 5 };
__global__ void __launch_bounds__(256) sy_masked_sample_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask, int rows,
                                                               int A, int N, int env_offset, unsigned seed_lo, unsigned seed_hi,
                                                               unsigned step, int greedy, int64_t* __restrict__ actions) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / A, a = row - b * A;
  const float* lg = logits + (size_t)row * N;
  const uint8_t* mk = mask + (size_t)row * N;
  float best = -INFINITY;
  int arg = -1;
  for (int j0 = lane * 4; j0 < N; j0 += 128) {  // a lane takes 4 consecutive nodes = one Philox block
    uint4 r = make_uint4(0, 0, 0, 0);
    if (!greedy) r = philox4x32(make_uint4((unsigned)(env_offset + b), step, (unsigned)(RNG_MASKED_SAMPLE + 16 * a), (unsigned)(j0 >> 2)), make_uint2(seed_lo, seed_hi));
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + q;
      if (j < N && mk[j]) {
        float key = lg[j];
        if (!greedy) {
          const float u = ((float)(w[q] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
          key -= __logf(-__logf(u));
        }
        if (key > best || arg < 0) {  // ascending j inside the lane: the first maximum wins
          best = key;
          arg = j;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (oa >= 0 && (arg < 0 || ob > best || (ob == best && oa < arg))) {
      best = ob;
      arg = oa;
    }
  }
  if (lane == 0) actions[row] = (int64_t)arg;
}

}  // namespace

extern "C" {

int sy_policy_abi_version(void) { return SY_POLICY_ABI_VERSION; }
const char* sy_policy_last_error(void) { return g_err; }
long long sy_policy_launch_count(void) { return g_launches.load(); }

int32_t sy_gnn_param_count(int32_t K) { return (K < 1 || K > SY_POLICY_MAX_FEATURES) ? 0 : gnn_model_floats(gnn_kp(K)); }

int sy_gnn_q_values(const SyPolicyGraphs* graphs, const SyPolicyState* state, const float* params, int32_t K, float conv_epsilon,
                    int32_t feature_mode, float* q, sy_policy_stream_t stream) {
  GnnParams p;
  if (int rc = fill_gnn(graphs, state, params, K, conv_epsilon, feature_mode, p)) return rc;
  if (!q) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "NULL q");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned grid = 0;
  size_t dyn = 0;
  if (int rc = gnn_launch_shape(graphs, state, grid, dyn)) return rc;
  const int KP = gnn_kp(K);
  int rows = ((graphs->num_nodes * KP + 3) & ~3);  // floats per warp for the layer-1 rows; 0 = recompute per node
  if (dyn + (size_t)GNN_WARPS * rows * sizeof(float) > 150 * 1024) rows = 0;
  dyn += (size_t)GNN_WARPS * rows * sizeof(float);
  switch (KP) {
    case 4: CUDA_TRY(cudaFuncSetAttribute(sy_gnn_q_dense_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            sy_gnn_q_dense_kernel<4><<<grid, GNN_WARPS * 32, dyn, s>>>(p, q, rows); break;
    case 8: CUDA_TRY(cudaFuncSetAttribute(sy_gnn_q_dense_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            sy_gnn_q_dense_kernel<8><<<grid, GNN_WARPS * 32, dyn, s>>>(p, q, rows); break;
    default: CUDA_TRY(cudaFuncSetAttribute(sy_gnn_q_dense_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
             sy_gnn_q_dense_kernel<16><<<grid, GNN_WARPS * 32, dyn, s>>>(p, q, rows); break;
  }
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_POLICY_OK;
}

int sy_gnn_act(const SyPolicyGraphs* graphs, const SyPolicyState* state, const float* params, int32_t K, float conv_epsilon,
               int32_t feature_mode, float epsilon_mrx, float epsilon_police, uint64_t seed, uint32_t step_counter, int64_t* actions,
               float* q_taken, sy_policy_stream_t stream) {
  GnnParams p;
  if (int rc = fill_gnn(graphs, state, params, K, conv_epsilon, feature_mode, p)) return rc;
  if (!actions) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "NULL actions");
  p.eps_mrx = epsilon_mrx;
  p.eps_police = epsilon_police;
  p.seed_lo = (unsigned)(seed & 0xFFFFFFFFu);
  p.seed_hi = (unsigned)(seed >> 32);
  p.step = step_counter;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned grid = 0;
  size_t dyn = 0;
  if (int rc = gnn_launch_shape(graphs, state, grid, dyn)) return rc;
  switch (gnn_kp(K)) {
    case 4: CUDA_TRY(cudaFuncSetAttribute(sy_gnn_act_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            sy_gnn_act_kernel<4><<<grid, GNN_WARPS * 32, dyn, s>>>(p, actions, q_taken); break;
    case 8: CUDA_TRY(cudaFuncSetAttribute(sy_gnn_act_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            sy_gnn_act_kernel<8><<<grid, GNN_WARPS * 32, dyn, s>>>(p, actions, q_taken); break;
    default: CUDA_TRY(cudaFuncSetAttribute(sy_gnn_act_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
             sy_gnn_act_kernel<16><<<grid, GNN_WARPS * 32, dyn, s>>>(p, actions, q_taken); break;
  }
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_POLICY_OK;
}

int sy_policy_set_option(const char* name, int32_t value) {
  if (name && std::strcmp(name, "mappo_tensor_cores") == 0) {
    g_mappo_tc = value;
    return SY_POLICY_OK;
  }
  return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "unknown option");
}

int sy_policy_check(sy_policy_stream_t stream) {
  if (!g_tc_flag) return SY_POLICY_OK;
  int flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, g_tc_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  if (flag == 1) return fail(SY_POLICY_ERR_CUDA, "tensor-core MAPPO kernel: MMA completion barrier timed out");
  if (flag == 2) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "tensor-core MAPPO kernel: a node has more valid moves than max_degree promised");
  return SY_POLICY_OK;
}

int32_t sy_mappo_param_count(int32_t obs_size, int32_t hidden, int32_t num_nodes) {
  if (obs_size < 1 || hidden < 1 || num_nodes < 1) return 0;
  return hidden * obs_size + hidden + num_nodes * hidden + num_nodes;
}

int sy_mappo_act(const SyPolicyGraphs* graphs, const SyPolicyState* state, const float* obs, int32_t obs_size, int32_t hidden,
                 const float* params, const int32_t* policy_of_agent, int32_t max_degree, uint64_t seed, uint32_t step_counter,
                 int64_t* actions, float* log_probs, float* probs, sy_policy_stream_t stream) {
  if (int rc = check_common(graphs, state)) return rc;
  if (!params || !policy_of_agent || !actions || !log_probs) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "NULL MAPPO arguments");
  if (obs_size < 1 || hidden < 1 || hidden > 1024) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "bad MAPPO layer sizes");
  if (!obs && obs_size < state->num_agents - 1) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "obs == NULL (trainer features) needs obs_size >= number of police");
  MappoParams p;
  std::memset(&p, 0, sizeof(p));
  p.g = *graphs;
  p.st = *state;
  p.obs = obs;
  p.params = params;
  p.obs_size = obs_size;
  p.H = hidden;
  p.HP = (hidden + MP_HC - 1) / MP_HC * MP_HC;
  p.policy_floats = sy_mappo_param_count(obs_size, hidden, graphs->num_nodes);
  for (int a = 0; a < state->num_agents; ++a) {
    if (policy_of_agent[a] < 0) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "negative policy index");
    p.policy_of_agent[a] = policy_of_agent[a];
  }
  p.seed_lo = (unsigned)(seed & 0xFFFFFFFFu);
  p.seed_hi = (unsigned)(seed >> 32);
  p.step = step_counter;
  // tensor-core path (tcgen05): N <= 256 nodes, operands + epilogue scratch within one SM's shared memory, and at most
  // TC_MAXV neighbours per node (max_degree from the host; 0 = unknown -> CUDA-core kernel)
  const TcShape ts = tc_shape(hidden, graphs->num_nodes, obs_size);
  const bool want_tc = g_mappo_tc != 0 && max_degree > 0 && max_degree <= TC_MAXV && graphs->num_nodes <= 256 &&
                       graphs->num_nodes >= 16 && hidden <= 64 && obs_size <= TC_DMAX && ts.total_bytes <= 220 * 1024;
  if (want_tc) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!g_tc_flag) CUDA_TRY(cudaMalloc(&g_tc_flag, sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(g_tc_flag, 0, sizeof(int), (cudaStream_t)stream));
    CUDA_TRY(cudaFuncSetAttribute(sy_mappo_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ts.total_bytes));
    const int ntiles = (state->num_envs + TC_ROWS - 1) / TC_ROWS;
    int gx = sms / state->num_agents;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    sy_mappo_act_tc_kernel<<<dim3((unsigned)gx, (unsigned)state->num_agents), TC_THREADS, ts.total_bytes, (cudaStream_t)stream>>>(
        p, actions, log_probs, probs, g_tc_flag);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return SY_POLICY_OK;
  }
  const size_t smem = (size_t)MP_ROWS * (p.HP + graphs->num_nodes + obs_size) * sizeof(float);
  if (smem > 220 * 1024) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "hidden + num_nodes too large for the MAPPO kernel's shared memory");
  CUDA_TRY(cudaFuncSetAttribute(sy_mappo_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((unsigned)((state->num_envs + MP_ROWS - 1) / MP_ROWS), (unsigned)state->num_agents);
  sy_mappo_act_kernel<<<grid, MP_THREADS, smem, (cudaStream_t)stream>>>(p, actions, log_probs, probs);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_POLICY_OK;
}

int sy_masked_sample(const float* logits, const uint8_t* mask, int32_t num_envs, int32_t num_agents, int32_t num_nodes,
                     int32_t env_offset, uint64_t seed, uint32_t step_counter, int32_t greedy, int64_t* actions,
                     sy_policy_stream_t stream) {
  if (!logits || !mask || !actions || num_envs < 1 || num_agents < 1 || num_nodes < 1)
    return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "bad sy_masked_sample arguments");
  const int rows = num_envs * num_agents;
  sy_masked_sample_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(logits, mask, rows, num_agents, num_nodes, env_offset,
                                                                                       (unsigned)(seed & 0xFFFFFFFFu), (unsigned)(seed >> 32),
                                                                                       step_counter, greedy, actions);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_POLICY_OK;
}

int sy_mappo_values(const float* global_obs, int32_t num_rows, int32_t obs_size, int32_t hidden, const float* params, float* values,
                    sy_policy_stream_t stream) {
  if (!global_obs || !params || !values || num_rows <= 0 || obs_size < 1 || hidden < 1)
    return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "bad critic arguments");
  const size_t smem = (size_t)(hidden * obs_size + 2 * hidden + 1) * sizeof(float);
  if (smem > 200 * 1024) return fail(SY_POLICY_ERR_INVALID_ARGUMENT, "critic too large for shared memory");
  CUDA_TRY(cudaFuncSetAttribute(sy_mappo_values_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sy_mappo_values_kernel<<<(unsigned)((num_rows + 255) / 256), 256, smem, (cudaStream_t)stream>>>(global_obs, num_rows, obs_size, hidden, params, values);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return SY_POLICY_OK;
}

}  // extern "C"
