/*
 * dqn_b200.h — C-ABI of libdqn_b200.so: the B200-native replacement for the hot path of
 * mhauskn/dqn-hfo's learner class dqn::DQN (reference src/dqn.hpp:56-202).
 *
 * The reference has no FFI layer: dqn_main.cpp links class DQN statically and DQN calls Caffe.
 * This header is the boundary a maintainer binds instead of Caffe; the C++ mirror of dqn::DQN
 * (dqn-hfo_b200/host/dqn.hpp) and the ctypes binding (dqn-hfo_b200/binding.py) sit on top of it.
 * Each entry point cites the reference member it replaces (file:line under /root/reference/src).
 *
 * Conventions: plain C types, caller-allocated HOST buffers unless a name says _device, every
 * function returns 0 on success and a negative code on failure (dqnb_last_error() has the text;
 * the C++ mirror turns failures into abort() to keep the reference's CHECK/LOG(FATAL) behaviour,
 * dqn.cpp:698-699 etc.).  A handle is thread-compatible (external synchronisation), like one
 * dqn::DQN per agent thread (dqn_main.cpp:264).  There is no CPU fallback: create() fails when
 * no sm_100 device is present.
 */
#ifndef DQN_B200_H_
#define DQN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DQNB_MAX_HIDDEN 8
#define DQNB_ACTOR_OUT 10 /* dqn.hpp:28 ActorOutput = 4 action logits + 6 action params */

typedef struct dqnb_handle_s *dqnb_handle;

enum { DQNB_ACTOR = 0, DQNB_CRITIC = 1, DQNB_ACTOR_TARGET = 2, DQNB_CRITIC_TARGET = 3 };

/* gemm_mode: how the dense layers are contracted.
 *   DQNB_GEMM_TCGEN05_3XTF32  tcgen05.mma kind::tf32, operands split hi+lo (3 MMAs / k-step),
 *                             TMA-fed, TMEM accumulators: the product path.
 *   DQNB_GEMM_SIMT_FP32       plain fp32 FFMA tiles: on-device verification mode for the parity
 *                             tests (same epilogues, same data layout). */
enum { DQNB_GEMM_TCGEN05_3XTF32 = 0, DQNB_GEMM_SIMT_FP32 = 1 };

typedef struct {
  int32_t struct_size;       /* = sizeof(dqnb_config), for ABI checks */
  int32_t device;            /* CUDA device ordinal */
  int32_t state_size;        /* dqn.hpp:198 state_size_ */
  int32_t batch;             /* dqn.hpp:19 kMinibatchSize (runtime here; per GPU) */
  int32_t n_hidden;          /* dqn.cpp:425,449: {1024,512,256,128} */
  int32_t hidden[DQNB_MAX_HIDDEN];
  int32_t replay_capacity;   /* dqn.cpp:25 FLAGS_memory */
  int32_t max_act_batch;     /* rows SelectActions/EvaluateAction accept (dqn.cpp:699: <= batch) */
  double gamma;              /* dqn.cpp:24 */
  double beta;               /* dqn.cpp:31 */
  float tau;                 /* dqn.cpp:22 */
  int32_t soft_update_freq;  /* dqn.cpp:23 */
  float actor_lr;            /* dqn_main.cpp:33 */
  float critic_lr;           /* dqn_main.cpp:34 */
  float momentum;            /* dqn_main.cpp:31 (Adam beta1) */
  float momentum2;           /* dqn_main.cpp:32 (Adam beta2) */
  float delta;               /* Caffe SolverParameter.delta (Adam eps), default 1e-8 */
  float clip_gradients;      /* dqn_main.cpp:35; < 0 disables */
  uint64_t seed;             /* dqn.cpp:21 FLAGS_seed: keys the device sampler */
  int32_t gemm_mode;         /* DQNB_GEMM_* */
  int32_t use_graph;         /* 1: replay one captured CUDA graph per update */
  int32_t world_size;        /* data-parallel replicas (gradient all-reduce); 1 = single GPU */
  int32_t rank;
} dqnb_config;

/* Fills *cfg with the reference's defaults (dqn.cpp:21-31, dqn_main.cpp:30-37; S=58, B=32). */
void dqnb_default_config(dqnb_config *cfg);

const char *dqnb_last_error(void);
const char *dqnb_version(void);

/* ctor + Initialize + CloneNet x2 (dqn.cpp:457-483, :622-662).  Weights start at zero: the
 * reference's gaussian(0.01) fill uses Caffe's own RNG and is not reproducible, so callers
 * inject weights with dqnb_set_params (or dqnb_init_params). */
int dqnb_create(const dqnb_config *cfg, dqnb_handle *out);
int dqnb_destroy(dqnb_handle h);

/* Learnable-parameter count of a net in Caffe order (Net::learnable_params):
 * ip1.W[H1 x in] ip1.b .. ip4.b, then action_layer.W[4xH4] .b actionpara_layer.W[6xH4] .b
 * (actor, dqn.cpp:418-429) or q_values_layer.W[1xH4] .b (critic, dqn.cpp:431-454). */
int64_t dqnb_param_count(dqnb_handle h, int net);
/* CopyTrainedLayersFrom / ToProto equivalents on flat Caffe-order arrays (dqn.cpp:529,:1024). */
int dqnb_set_params(dqnb_handle h, int net, const float *params);
int dqnb_get_params(dqnb_handle h, int net, float *params);
/* gaussian(std) weights / zero biases drawn on the host from mt19937(seed) (dqn.cpp:350-352),
 * then CloneNet into the targets. */
int dqnb_init_params(dqnb_handle h, uint64_t seed, float std);
/* CloneNet(critic->critic_target), CloneNet(actor->actor_target) (dqn.cpp:660-661). */
int dqnb_clone_targets(dqnb_handle h);
/* ShareParameters (dqn.cpp:1048-1079; ShareLayer :1037-1046 -> Blob::ShareData): the first n layers-with-parameters
 * (Caffe layer order: ip1..ipK, then action_layer, actionpara_layer | q_values_layer) of src's actor / critic AND of
 * their target nets overwrite dst's.  Blob::ShareData aliases memory; two handles cannot alias sub-ranges of a flat
 * buffer, so a sharing group calls this after each member's update (write-through; host/dqn.cpp ShareParameters).
 * Both handles are synchronised; dst's act-path snapshot of the actor is refreshed. */
int dqnb_copy_shared_layers(dqnb_handle dst, dqnb_handle src, int32_t n_actor_layers, int32_t n_critic_layers);
/* Solver::Restore / Snapshot state: Adam history (m, v) and iter (dqn.cpp:545,:554,:589-590). */
int dqnb_set_opt_state(dqnb_handle h, int net, const float *m, const float *v, int32_t iter);
int dqnb_get_opt_state(dqnb_handle h, int net, float *m, float *v, int32_t *iter);
/* actor_iter()/critic_iter() (dqn.hpp:129-130). */
int dqnb_iters(dqnb_handle h, int32_t *actor_iter, int32_t *critic_iter);

/* AddTransitions (dqn.cpp:775-781): evicts while size+n >= capacity, then appends n rows.
 * s, s_next: [n x state_size]; act10: [n x 10]; reward, mc_target: [n]; terminal: [n] (non-zero
 * <=> the transition has no next state, dqn.cpp:878; s_next rows of terminal entries are ignored).
 * mc_target is the LabelTransitions field (dqn.cpp:783-797), computed by the caller. */
int dqnb_add_transitions(dqnb_handle h, int32_t n, const float *s, const float *act10,
                         const float *reward, const float *mc_target, const float *s_next,
                         const uint8_t *terminal);
/* AddTransition (dqn.cpp:768-773): evicts one row only when size == capacity. */
int dqnb_add_transition(dqnb_handle h, const float *s, const float *act10, float reward,
                        float mc_target, const float *s_next, uint8_t terminal);
int32_t dqnb_memory_size(dqnb_handle h);          /* dqn.hpp:112 */
int dqnb_clear_memory(dqnb_handle h);             /* dqn.hpp:106 */
/* Reads rows [first, first+n) in deque order (0 = oldest), for SnapshotReplayMemory
 * (dqn.cpp:1146-1178) and tests.  Any output pointer may be NULL. */
int dqnb_get_transitions(dqnb_handle h, int32_t first, int32_t n, float *s, float *act10,
                         float *reward, float *mc_target, float *s_next, uint8_t *terminal);

/* n x UpdateActorCritic (dqn.cpp:828-972) with minibatch indices drawn on the device
 * (counter-based Philox4x32-10 keyed by seed and update number; replaces
 * SampleTransitionsFromMemory dqn.cpp:501-509).  critic_loss / avg_q: [n_updates] or NULL. */
int dqnb_update(dqnb_handle h, int32_t n_updates, float *critic_loss, float *avg_q);
/* The same updates without waiting for them: the call returns once they are enqueued; *last_step is the
 * 1-based sequence number of the last one.  dqnb_add_transitions stays legal while updates are in flight:
 * appends travel on a copy stream, ordered after the last enqueued minibatch gather and before the next
 * one, so every update samples exactly the memory the reference's sequential Update() (dqn.cpp:799-826)
 * would have seen. */
int dqnb_update_async(dqnb_handle h, int32_t n_updates, int64_t *last_step);
/* (critic_loss, avg_q) of updates [first_step, first_step + n): blocks (spinning on a host-mapped counter
 * the device publishes, no stream sync) until the last of them has finished.  The last 4096 updates are kept. */
int dqnb_results(dqnb_handle h, int64_t first_step, int32_t n, float *critic_loss, float *avg_q);
/* One UpdateActorCritic on caller-chosen deque indices idx[batch] (parity hook: the reference's
 * std::uniform_int_distribution stream is libstdc++-defined, so tests inject indices). */
int dqnb_update_with_indices(dqnb_handle h, const int32_t *idx, float *critic_loss, float *avg_q);
/* Same as dqnb_update but returns the device time of the n updates measured with CUDA events on
 * the handle's stream (the reference's Benchmark(), dqn.cpp:487-498). */
int dqnb_benchmark(dqnb_handle h, int32_t n_updates, float *elapsed_ms);
/* Device time (CUDA events) of just the dense-layer launches of one update, replayed back to back as a captured graph
 * and averaged over reps: the live measurement behind bench.py's roofline line.  Leaves learner state untouched except
 * for scratch activations. */
int dqnb_benchmark_gemms(dqnb_handle h, int32_t reps, float *ms_per_update, int32_t *gemm_launches);
/* Draws the indices the next dqnb_update would use, without updating (tests). */
int dqnb_peek_sample_indices(dqnb_handle h, int32_t *idx);

/* SelectActionGreedily on a batch (dqn.cpp:734-766): states [n x state_size] -> out10 [n x 10].
 * The epsilon coin flip of SelectActions (dqn.cpp:695-711) stays in the host wrapper so that the
 * host RNG stream is the reference's. */
int dqnb_select_actions(dqnb_handle h, int32_t n, const float *states, float *out10);
/* 1..64 rows never wait for the learner: the call runs a captured graph of skinny-M kernels on the handle's ACT stream,
 * reading a snapshot of the actor that the optimiser publishes at the end of every update, with rows and completion
 * travelling through host-mapped memory (csrc/kernels.cuh act_layer_kernel).  More rows (up to max_act_batch) use the
 * tcgen05 layer kernels on the learner's stream.
 * Split form for overlapping the caller's own work: enqueue, then wait. */
int dqnb_select_actions_async(dqnb_handle h, int32_t n, const float *states);
int dqnb_select_actions_wait(dqnb_handle h, int32_t n, float *out10);
/* CriticForward (dqn.cpp:982-1020) / EvaluateAction (dqn.cpp:688-693): q [n]. */
int dqnb_evaluate(dqnb_handle h, int32_t n, const float *states, const float *act10, float *q);

/* Data-parallel replicas: one process per GPU, config.world_size / config.rank (no counterpart in the reference,
 * SURVEY 8e).  Two ways to sum the gradients after each backward pass; the product path is the peer-memory exchange
 * below.  VALIDATION ALTERNATIVE: ncclAllReduce on the handle's stream; id128 is an ncclUniqueId obtained on rank 0 and
 * broadcast by the caller (same results bit for bit, slower). */
int dqnb_comm_unique_id(void *id128);
int dqnb_comm_init(dqnb_handle h, const void *id128);
/* Product path for the gradient exchange: a fused reduce-scatter + all-gather kernel over NVLink peer
 * memory (csrc/kernels.cuh p2p_allreduce_kernel), no NCCL.  Every rank exports the 64-byte CUDA-IPC handle
 * of its exchange buffer, the caller gathers the world_size handles in rank order (any transport) and
 * every rank maps them.  Selecting either exchange rebuilds the captured update graph. */
int dqnb_comm_p2p_handle(dqnb_handle h, void *handle64);
int dqnb_comm_p2p_init(dqnb_handle h, const void *handles /* world_size x 64 bytes */);
/* 0 = healthy; 1 = a peer missed the exchange kernel's timeout (DQNB_P2P_TIMEOUT_MS, default 20 s).  The flag is sticky:
 * from then on the optimiser kernels leave parameters, moments and iteration counters as they are, results read NaN, and
 * dqnb_update / dqnb_update_with_indices / dqnb_results return -1 with a message. */
int dqnb_comm_status(dqnb_handle h);

/* Blocks until all work queued on the handle has finished. */
int dqnb_sync(dqnb_handle h);
/* Number of kernels of this library launched so far on behalf of the handle (graph replays
 * count the kernels inside the graph). */
int64_t dqnb_kernel_launches(dqnb_handle h);

/* Debug taps used by the parity tests: copies an internal fp32 buffer of the last update.
 * Names: "y" "q" "q_next" "a_pi" "q_pi" "d_raw" "d_inv" [batch(,10)], "critic_grad" "actor_grad"
 * (pre-clip, Caffe order), "critic_gnorm" "actor_gnorm" [1].  Returns elements written. */
int64_t dqnb_debug_read(dqnb_handle h, const char *name, float *out, int64_t capacity);

/* Stand-alone dense contraction through the same kernels the layers use (kernel unit tests):
 * C[MxN] = A * B^T-like product with fp32 host operands.
 *   a_mn_major = 0: A is [M x K] row-major;  1: A is stored [K x M] row-major
 *   b_mn_major = 0: B is [N x K] row-major;  1: B is stored [K x N] row-major
 * splits > 1 exercises split-K (partials summed on the device). */
int dqnb_gemm_test(int device, int gemm_mode, int a_mn_major, int b_mn_major, int M, int N, int K,
                   int splits, const float *A, const float *B, float *C, float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* DQN_B200_H_ */
