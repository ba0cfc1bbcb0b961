/* sy_policy.h -- C ABI of the batched policy forward for the reference's two shipped agents,
 * the step either side of the env in BASELINE config 5.  Stand-alone library (libsy_policy.so): every argument is a
 * plain pointer / size, all pointers are DEVICE pointers owned by the caller unless marked HOST, calls are
 * asynchronous on `stream` and return SY_POLICY_OK or an error code (message: sy_policy_last_error()).
 *
 * Replaces, for all B envs and all agents in one launch:
 *   GNNAgent.select_action + GNNModel.forward      /root/reference/src/agent/gnn_agent.py:45-82, 230-257
 *     (2 x torch_geometric AntiSymmetricConv(phi = GCNConv(K, K, bias=False), num_iters=1, epsilon=0.1, gamma=0.1,
 *      act=tanh) + ReLU, Linear(K, 1); epsilon-greedy over the valid moves) as driven by the trainer's loop
 *     src/training/gnn_trainer.py:201-232 with create_graph_data (src/training/utils.py:151-211)
 *   MappoAgent.select_action + AgentPolicy.forward /root/reference/src/agent/mappo_agent.py:6-29, 87-142
 *     (Linear-ReLU-Linear-softmax, mask, renormalise with the reference's fall-backs, Categorical sample, log-prob)
 *   CentralCritic.forward                          /root/reference/src/agent/mappo_agent.py:32-44
 * torch_geometric is a requirements.txt dependency of the reference (unpinned, not vendored): the GNN follows its
 * published algorithm (AntiSymmetricConv / GCNConv / gcn_norm, PyG 2.x).
 */
#ifndef SY_POLICY_H
#define SY_POLICY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SY_POLICY_ABI_VERSION 2
#define SY_POLICY_MAX_FEATURES 16 /* K = node_feature_size = agents (utils.py:176) */

enum { SY_POLICY_OK = 0, SY_POLICY_ERR_INVALID_ARGUMENT = 1, SY_POLICY_ERR_CUDA = 2 };

/* what the GNN sees as node features x [N, K] */
enum {
  SY_FEATURES_ENV = 0,      /* the env's node_features observation (yard.py:283-291): column k is one-hot at agent k's
                               node, MrX's column blank while he is hidden */
  SY_FEATURES_REFERENCE = 1 /* create_graph_data as written (utils.py:176-199): column 0 at MrX_pos (index -1 = last
                               node while hidden), columns 1..P-1 ALL at Police0's node, column P empty */
};

typedef void* sy_policy_stream_t;

/* graph pool: the env's undirected CSR (as given to sy_load_graphs) for the valid-move rule, plus the DIRECTED
 * in-edge lists GCNConv aggregates over: create_graph_data passes env.board.edge_links.T unsymmetrised
 * (utils.py:169), so node v receives from u for every stored edge (u, v); in_coef = deg^-1/2[u] * deg^-1/2[v],
 * self_coef = 1/deg[v] with deg = 1 + in-degree (gcn_norm with self-loops). */
typedef struct SyPolicyGraphs {
  int32_t num_graphs, num_nodes, nnz_stride, in_stride;
  const int32_t* row_ptr; /* [G, N+1] */
  const int32_t* col;     /* [G, nnz_stride] neighbours ascending */
  const int32_t* w;       /* [G, nnz_stride] edge weights */
  const int32_t* in_ptr;  /* [G, N+1] */
  const int32_t* in_src;  /* [G, in_stride] */
  const float* in_coef;   /* [G, in_stride] */
  const float* self_coef; /* [G, N] */
} SyPolicyGraphs;

/* the slice of the env state the policies read (the buffers of SyState / SyObs, include/sy_env.h) */
typedef struct SyPolicyState {
  int32_t num_envs, num_agents; /* B, A = police + 1 (agent 0 = MrX) */
  int32_t toll;                 /* a move is valid iff w + toll <= budget (yard.py:420-472 + tolls) */
  int32_t env_offset;           /* global index of env 0 (Philox streams of a sharded batch) */
  const int32_t* pos;           /* [B, A] */
  const int32_t* money;         /* [B, A] */
  const int32_t* graph_id;      /* [B] */
  const int32_t* mrx_revealed;  /* [B] MrX's node if visible else -1; NULL = always visible */
} SyPolicyState;

int sy_policy_abi_version(void);
const char* sy_policy_last_error(void);
long long sy_policy_launch_count(void);

/* Packed GNNModel parameters, floats, K padded to KP = 4 | 8 | 16 columns (sy_gnn_param_count(K) floats per model):
 *   conv1: WasT [KP, KP] (WasT[k][c] = (W - W^T - gamma I)[c][k]), ThT [KP, KP] (ThT[k][c] = phi.lin.weight[c][k]),
 *          bias [KP];  conv2: the same;  out_w [KP];  out_b, 0, 0, 0.
 * params holds two models: [0] MrX's agent, [1] the police agent (gnn_trainer.py:147-165). */
int32_t sy_gnn_param_count(int32_t K);

/* q [B, 2, N] float32: GNNModel.forward of both agents on every env's graph (gnn_agent.py:249-257) */
int sy_gnn_q_values(const SyPolicyGraphs* graphs, const SyPolicyState* state, const float* params, int32_t K,
                    float conv_epsilon, int32_t feature_mode, float* q, sy_policy_stream_t stream);

/* actions int64 [B, A]: GNNAgent.select_action for every agent (gnn_agent.py:45-82): with probability epsilon a
 * uniform valid move, else the valid move with the largest Q (first one on ties); -1 (DEFAULT_ACTION) without a valid
 * move (gnn_trainer.py:227-229).  Randomness: Philox(seed; env_offset + env, step_counter, 3, agent).  Only the Q
 * values of the valid moves of exploiting agents are evaluated.  q_taken [B, A] (may be NULL): Q of the chosen move,
 * NaN for explored / missing moves. */
int sy_gnn_act(const SyPolicyGraphs* graphs, const SyPolicyState* state, const float* params, int32_t K,
               float conv_epsilon, int32_t feature_mode, float epsilon_mrx, float epsilon_police, uint64_t seed,
               uint32_t step_counter, int64_t* actions, float* q_taken, sy_policy_stream_t stream);

/* MappoAgent.select_action for every (env, agent) (mappo_agent.py:87-142).
 * obs [B, A, obs_size] float32, or NULL = the kernel builds MappoTrainer's own observations from the state
 * (mappo_trainer.py:171-199: MrX sees obs["MrX_pos"], officer i the officers' nodes obs["Polices_pos"].sum(dim=1); raw
 * node ids as float32, zero-padded to obs_size >= P columns) -- no observation tensor is materialised at all;
 * params: per policy  W1 [H, obs_size], b1 [H], W2 [N, H], b2 [N]  (nn.Linear layout),
 * sy_mappo_param_count floats each; policy_of_agent [A] (HOST) = index of the AgentPolicy each agent uses
 * (mappo_trainer.py:124-147: MrX's agent has one policy, the police agent one per officer).
 * probs = softmax(logits) * mask; renormalised by (sum + 1e-8); sum <= 1e-8 -> uniform over the mask, empty mask ->
 * uniform over all N (mappo_agent.py:121-133).  Sampling: inverse CDF with Philox(seed; env, step_counter, 4, agent).
 * actions int64 [B, A], log_probs float32 [B, A], probs (may be NULL) float32 [B, A, N] the detached distribution. */
int32_t sy_mappo_param_count(int32_t obs_size, int32_t hidden, int32_t num_nodes);
int sy_mappo_act(const SyPolicyGraphs* graphs, const SyPolicyState* state, const float* obs, int32_t obs_size,
                 int32_t hidden, const float* params, const int32_t* policy_of_agent, int32_t max_degree,
                 uint64_t seed, uint32_t step_counter, int64_t* actions, float* log_probs, float* probs,
                 sy_policy_stream_t stream);
/* max_degree: largest neighbour count in the pool (0 = unknown).  With num_nodes <= 256, hidden <= 64, obs_size <= 16,
 * max_degree <= 16 and the operands fitting one SM's shared memory the logits GEMM runs on the tensor cores (tcgen05, 3xTF32, accumulators in
 * tensor memory); otherwise on the CUDA cores.  sy_policy_set_option("mappo_tensor_cores", 0) forces the latter.
 * sy_policy_check synchronises the stream and reports a (never expected) failure of the tensor-core kernel. */
int sy_policy_set_option(const char* name, int32_t value);
int sy_policy_check(sy_policy_stream_t stream);

/* One node per (env, agent) from ANY policy's logits [B, A, N] (float32) restricted to the env's action_mask [B, A, N]
 * (bool / uint8): the batched form of the trainers' choice among the valid moves (gnn_trainer.py:221-229,
 * mappo_agent.py:121-140) for policies that are not the two built-in agents.  greedy != 0: first argmax over the legal
 * nodes; else an exact sample of softmax(logits | mask) by Gumbel-max with Philox(seed; env_offset + env, step_counter,
 * 5 + 16 * agent, node / 4).  actions int64 [B, A], -1 (DEFAULT_ACTION) for rows without a legal node. */
int sy_masked_sample(const float* logits, const uint8_t* mask, int32_t num_envs, int32_t num_agents, int32_t num_nodes,
                     int32_t env_offset, uint64_t seed, uint32_t step_counter, int32_t greedy, int64_t* actions,
                     sy_policy_stream_t stream);

/* CentralCritic.forward (mappo_agent.py:32-44): values [M] = W2 relu(W1 x + b1) + b2 for global_obs [M, D];
 * params: W1 [H, D], b1 [H], W2 [H], b2 [1]. */
int sy_mappo_values(const float* global_obs, int32_t num_rows, int32_t obs_size, int32_t hidden, const float* params,
                    float* values, sy_policy_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SY_POLICY_H */
