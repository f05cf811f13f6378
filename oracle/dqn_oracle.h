/*
 * dqn_oracle.h — CPU ORACLE (test infrastructure, NOT the product path).
 *
 * A plain-C restatement of the arithmetic of dqn-hfo's replay-minibatch
 * actor-critic update (reference: src/dqn.cpp:828-972 UpdateActorCritic and
 * the helpers it calls) in the operation order BVLC Caffe executes it.
 *
 * PARITY UNPINNED: the reference ships no tests / golden vectors and its
 * arithmetic lives in external Caffe @2ef584785c8ade90260eb117f189146364494183
 * (reference README.md:7-45), absent from this tree.  This oracle is therefore
 * pinned only against an independent float64 autograd derivation
 * (tests/test_oracle_autograd.py), a second CPU model in Caffe's BLAS call
 * order (tests/test_oracle_blas_order.py), a third-party implementation of the
 * Caffe layers for the forward half (OpenCV's Caffe importer reading the
 * nets as .caffemodel files, tests/test_caffemodel_opencv.py) and its own
 * committed known-answer vectors (tests/golden/), never against outputs of
 * the reference itself.
 *
 * Only tests/, __graft_entry__.smoke(), the parity block of bench.py --gpus N
 * (scripts/dp_parity.py: checker only) and bench.py's cpu_baseline /
 * --impl reference legs may call anything in this file.
 */
#ifndef DQN_ORACLE_H_
#define DQN_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DQO_MAX_HIDDEN 8
#define DQO_ACTION_SIZE 4        /* dqn.hpp:20 kActionSize      */
#define DQO_ACTION_PARAM_SIZE 6  /* dqn.hpp:21 kActionParamSize */
#define DQO_ACTOR_OUT 10         /* dqn.hpp:28 ActorOutput      */

typedef struct {
  int32_t state_size;              /* dqn.hpp:198 state_size_ (50+9*players, hfo_game.hpp:14) */
  int32_t batch;                   /* dqn.hpp:19 kMinibatchSize (32 in the reference) */
  int32_t n_hidden;                /* dqn.cpp:425,449: {1024,512,256,128} -> 4 */
  int32_t hidden[DQO_MAX_HIDDEN];
  double gamma;                    /* dqn.cpp:24 FLAGS_gamma .99 */
  double beta;                     /* dqn.cpp:31 FLAGS_beta  .5  */
  float tau;                       /* dqn.cpp:22 FLAGS_tau .001, narrowed at dqn.cpp:968 */
  int32_t soft_update_freq;        /* dqn.cpp:23 */
  float actor_lr;                  /* dqn_main.cpp:33 1e-5 */
  float critic_lr;                 /* dqn_main.cpp:34 1e-3 */
  float momentum;                  /* dqn_main.cpp:31 .95  (Adam beta1) */
  float momentum2;                 /* dqn_main.cpp:32 .999 (Adam beta2) */
  float delta;                     /* Caffe SolverParameter.delta default 1e-8 (Adam eps) */
  float clip_gradients;            /* dqn_main.cpp:35 10; <0 disables */
  int32_t caffe_wasted_work;       /* 1: also do the work Caffe does but nobody reads
                                      (force_backward layer-1 dX, critic dW in actor phase) */
  int32_t use_blas;                /* 1: route GEMMs through a dlopen()ed cblas_sgemm if loaded */
} dqo_config;

/* Optional taps on intermediate values (any pointer may be NULL). */
typedef struct {
  float *q_next;        /* [B]   target-critic value of s' (0 for terminal rows)      */
  float *y;             /* [B]   mixed TD target, dqn.cpp:892-900                     */
  float *q;             /* [B]   critic(s,a) before the critic step                   */
  float *critic_grad;   /* [Pc]  critic gradient before clipping (Caffe param order)  */
  float *a_pi;          /* [B*10] actor(s) with pre-update actor weights              */
  float *q_pi;          /* [B]   critic(s, a_pi) with post-update critic weights      */
  float *d_raw;         /* [B*10] dL/d(a,p) out of the critic, before inversion       */
  float *d_inv;         /* [B*10] after inverting gradients, dqn.cpp:927-957          */
  float *actor_grad;    /* [Pa]  actor gradient before clipping                       */
  float *critic_gnorm;  /* [1]   L2 norm seen by ClipGradients (critic)               */
  float *actor_gnorm;   /* [1]                                                        */
} dqo_taps;

/* Parameter counts, Caffe learnable_params order:
 * actor : ip1.W[H1xS] ip1.b .. ip4.W ip4.b action_layer.W[4xH4] .b[4] actionpara_layer.W[6xH4] .b[6]
 * critic: ip1.W[H1x(S+10)] ip1.b .. ip4.b q_values_layer.W[1xH4] .b[1]
 * (dqn.cpp:418-454).  W is [out x in] row-major (Caffe InnerProduct). */
int64_t dqo_actor_param_count(const dqo_config *cfg);
int64_t dqo_critic_param_count(const dqo_config *cfg);

/* actor forward (dqn.cpp:734-766 SelectActionGreedily): states [n x S] -> out [n x 10] */
int dqo_actor_forward(const dqo_config *cfg, const float *actor, int32_t n,
                      const float *states, float *out10);
/* critic forward (dqn.cpp:982-1020 CriticForward): -> q [n] */
int dqo_critic_forward(const dqo_config *cfg, const float *critic, int32_t n,
                       const float *states, const float *act10, float *q);

/* One UpdateActorCritic (dqn.cpp:828-972) on an already gathered minibatch.
 * term[n] != 0  <=>  transition n has no next state (dqn.cpp:878).
 * s_next rows of terminal transitions are ignored. */
int dqo_update(const dqo_config *cfg,
               float *actor, float *critic, float *actor_target, float *critic_target,
               float *actor_m, float *actor_v, float *critic_m, float *critic_v,
               int32_t *actor_iter, int32_t *critic_iter,
               const float *s, const float *act10, const float *reward,
               const float *mc_target, const uint8_t *term, const float *s_next,
               float *critic_loss, float *avg_q, const dqo_taps *taps);

/* SoftUpdateNet (dqn.cpp:1085-1096): to = (1-tau)*to ; to += tau*from */
void dqo_soft_update(int64_t n, float tau, const float *from, float *to);

/* LabelTransitions (dqn.cpp:783-797): in-place Monte-Carlo return of one episode */
void dqo_label_transitions(int32_t n, double gamma, const float *reward, float *mc_target);

/* Inverting gradients on one [n x 10] block (dqn.cpp:927-957) */
void dqo_invert_gradients(int32_t n, const float *a_pi, float *d10);

/* Replay-deque bookkeeping (dqn.cpp:768-781).  Models only sizes: returns the number of
 * pop_front()s performed and the new size, so that a ring implementation can be checked. */
int32_t dqo_replay_after_add_one(int32_t size, int32_t capacity, int32_t *new_size);
int32_t dqo_replay_after_add_many(int32_t size, int32_t capacity, int32_t n, int32_t *new_size);

/* GetAction (dqn.cpp:196-208): returns the action index (0 dash,1 turn,2 tackle,3 kick) */
int32_t dqo_get_action(const float *actor_output10, float *arg1, float *arg2);

/* Optional BLAS: dlopen()s a library exporting cblas_sgemm. Returns 0 on success. */
int dqo_load_blas(const char *path);
int dqo_blas_loaded(void);
void dqo_set_threads(int n);
int dqo_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
