/*
 * dqn_oracle.c — CPU ORACLE (test infrastructure, NOT the product path).
 * See dqn_oracle.h for scope and the "parity unpinned" note.
 *
 * Every block cites the reference line (relative to /root/reference/) whose
 * behaviour it restates.  Caffe semantics (InnerProduct, ReLU, Split, Concat,
 * EuclideanLoss, SGDSolver::ClipGradients, AdamSolver::ComputeUpdateValue,
 * Net::Update) are restated from the published BVLC Caffe algorithm at the
 * commit the reference pins (README.md:7-45); that source is not in the tree.
 */
#include "dqn_oracle.h"

#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------- */
/* BLAS-like kernels: deterministic portable versions + optional cblas        */
/* ------------------------------------------------------------------------- */

typedef void (*cblas_sgemm_fn)(int order, int ta, int tb, int M, int N, int K, float alpha,
                               const float *A, int lda, const float *B, int ldb, float beta,
                               float *C, int ldc);
static cblas_sgemm_fn g_sgemm = NULL;
static void *g_blas_handle = NULL;
static int g_threads = 0; /* 0 = OpenMP default */

int dqo_load_blas(const char *path) {
  void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1;
  cblas_sgemm_fn f = (cblas_sgemm_fn)dlsym(h, "cblas_sgemm");
  if (!f) f = (cblas_sgemm_fn)dlsym(h, "scipy_cblas_sgemm"); /* scipy's bundled OpenBLAS (LP64) */
  if (!f) { dlclose(h); return -2; }
  g_blas_handle = h;
  g_sgemm = f;
  return 0;
}
int dqo_blas_loaded(void) { return g_sgemm != NULL; }
void dqo_set_threads(int n) {
  g_threads = n;
  if (g_blas_handle) {
    void (*set)(int) = (void (*)(int))dlsym(g_blas_handle, "openblas_set_num_threads");
    if (!set) set = (void (*)(int))dlsym(g_blas_handle, "scipy_openblas_set_num_threads");
    if (set && n > 0) set(n);
  }
}
int dqo_get_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* Minimal pthread parallel-for over rows (no OpenMP runtime in this image).  Each output row is
 * produced by exactly one thread in a fixed order, so results do not depend on the thread count. */
typedef void (*row_fn)(int row_begin, int row_end, void *ctx);
typedef struct { row_fn fn; void *ctx; int b, e; } row_job;
static void *row_thread(void *p) { row_job *j = (row_job *)p; j->fn(j->b, j->e, j->ctx); return NULL; }
static void parallel_rows(int rows, int64_t work, row_fn fn, void *ctx) {
  int nt = dqo_get_threads();
  if (nt > 64) nt = 64;
  if (work < 400000 || nt <= 1 || rows < 2) { fn(0, rows, ctx); return; }
  if (nt > rows) nt = rows;
  pthread_t th[64]; row_job jobs[64];
  int per = (rows + nt - 1) / nt, started = 0;
  for (int t = 0; t < nt; ++t) {
    jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].b = t * per;
    jobs[t].e = (t + 1) * per < rows ? (t + 1) * per : rows;
    if (jobs[t].b >= jobs[t].e) break;
    if (t == nt - 1 || jobs[t].e == rows) { started = t; fn(jobs[t].b, jobs[t].e, ctx); break; }
    if (pthread_create(&th[t], NULL, row_thread, &jobs[t])) { fn(jobs[t].b, jobs[t].e, ctx); th[t] = 0; }
    started = t + 1;
  }
  for (int t = 0; t < started; ++t) if (th[t]) pthread_join(th[t], NULL);
}

#define LANES 16
/* sdot with LANES partial sums (the shape a SIMD cblas_sdot has). */
static float sdot(int64_t n, const float *x, const float *y) {
  float acc[LANES];
  for (int j = 0; j < LANES; ++j) acc[j] = 0.f;
  int64_t i = 0;
  for (; i + LANES <= n; i += LANES)
    for (int j = 0; j < LANES; ++j) acc[j] += x[i + j] * y[i + j];
  for (int j = 0; i < n; ++i, ++j) acc[j] += x[i] * y[i];
  for (int w = LANES / 2; w > 0; w >>= 1)
    for (int j = 0; j < w; ++j) acc[j] += acc[j + w];
  return acc[0];
}

typedef struct { int M, N, K; const float *A, *B; float *C; int accumulate; } gemm_ctx;
static void gemm_nt_rows(int mb, int me, void *p) {
  const gemm_ctx *g = (const gemm_ctx *)p;
  const int N = g->N, K = g->K;
  for (int m = mb; m < me; ++m)
    for (int n = 0; n < N; ++n) {
      float v = sdot(K, g->A + (int64_t)m * K, g->B + (int64_t)n * K);
      g->C[(int64_t)m * N + n] = g->accumulate ? g->C[(int64_t)m * N + n] + v : v;
    }
}
static void gemm_nn_rows(int mb, int me, void *p) {
  const gemm_ctx *g = (const gemm_ctx *)p;
  const int N = g->N, K = g->K;
  for (int m = mb; m < me; ++m) {
    float *c = g->C + (int64_t)m * N;
    for (int n = 0; n < N; ++n) c[n] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float a = g->A[(int64_t)m * K + k];
      const float *b = g->B + (int64_t)k * N;
      for (int n = 0; n < N; ++n) c[n] += a * b[n];
    }
  }
}
static void gemm_tn_rows(int mb, int me, void *p) {
  const gemm_ctx *g = (const gemm_ctx *)p;
  const int M = g->M, N = g->N, K = g->K;
  for (int m = mb; m < me; ++m) {
    float *c = g->C + (int64_t)m * N;
    for (int k = 0; k < K; ++k) {
      const float a = g->A[(int64_t)k * M + m];
      const float *b = g->B + (int64_t)k * N;
      for (int n = 0; n < N; ++n) c[n] += a * b[n];
    }
  }
}

/* C[MxN] = A[MxK] * B[NxK]^T  (+ C if accumulate) */
static void gemm_nt(int use_blas, int M, int N, int K, const float *A, const float *B, float *C,
                    int accumulate) {
  if (M == 0 || N == 0) return;
  if (use_blas && g_sgemm) {
    g_sgemm(101, 111, 112, M, N, K, 1.f, A, K, B, K, accumulate ? 1.f : 0.f, C, N);
    return;
  }
  gemm_ctx c = {M, N, K, A, B, C, accumulate};
  parallel_rows(M, (int64_t)M * N * K, gemm_nt_rows, &c);
}
/* C[MxN] = A[MxK] * B[KxN] */
static void gemm_nn(int use_blas, int M, int N, int K, const float *A, const float *B, float *C) {
  if (M == 0 || N == 0) return;
  if (use_blas && g_sgemm) {
    g_sgemm(101, 111, 111, M, N, K, 1.f, A, K, B, N, 0.f, C, N);
    return;
  }
  gemm_ctx c = {M, N, K, A, B, C, 0};
  parallel_rows(M, (int64_t)M * N * K, gemm_nn_rows, &c);
}
/* C[MxN] += A[KxM]^T * B[KxN] */
static void gemm_tn_acc(int use_blas, int M, int N, int K, const float *A, const float *B,
                        float *C) {
  if (M == 0 || N == 0 || K == 0) return;
  if (use_blas && g_sgemm) {
    g_sgemm(101, 112, 111, M, N, K, 1.f, A, M, B, N, 1.f, C, N);
    return;
  }
  gemm_ctx c = {M, N, K, A, B, C, 1};
  parallel_rows(M, (int64_t)M * N * K, gemm_tn_rows, &c);
}

/* ------------------------------------------------------------------------- */
/* Net layout (dqn.cpp:400-454)                                               */
/* ------------------------------------------------------------------------- */

typedef struct {
  int in_dim;
  int n_hidden;
  int hidden[DQO_MAX_HIDDEN];
  int n_heads;
  int head_out[2];
  int64_t w_off[DQO_MAX_HIDDEN], b_off[DQO_MAX_HIDDEN];
  int64_t hw_off[2], hb_off[2];
  int64_t count;
  int n_blobs;
  int64_t blob_off[2 * DQO_MAX_HIDDEN + 4], blob_cnt[2 * DQO_MAX_HIDDEN + 4];
} net_layout;

static void make_layout(const dqo_config *cfg, int is_critic, net_layout *L) {
  memset(L, 0, sizeof(*L));
  /* critic input = Concat(states, actions, action_params) axis 2, dqn.cpp:446-448 */
  L->in_dim = cfg->state_size + (is_critic ? DQO_ACTOR_OUT : 0);
  L->n_hidden = cfg->n_hidden;
  int64_t off = 0;
  int in = L->in_dim, nb = 0;
  for (int l = 0; l < cfg->n_hidden; ++l) { /* Tower, dqn.cpp:400-416 */
    L->hidden[l] = cfg->hidden[l];
    L->w_off[l] = off; L->blob_off[nb] = off; L->blob_cnt[nb++] = (int64_t)cfg->hidden[l] * in;
    off += (int64_t)cfg->hidden[l] * in;
    L->b_off[l] = off; L->blob_off[nb] = off; L->blob_cnt[nb++] = cfg->hidden[l];
    off += cfg->hidden[l];
    in = cfg->hidden[l];
  }
  if (is_critic) { /* q_values_layer, dqn.cpp:450 */
    L->n_heads = 1; L->head_out[0] = 1;
  } else {         /* action_layer(4), actionpara_layer(6), dqn.cpp:426-427 */
    L->n_heads = 2; L->head_out[0] = DQO_ACTION_SIZE; L->head_out[1] = DQO_ACTION_PARAM_SIZE;
  }
  for (int h = 0; h < L->n_heads; ++h) {
    L->hw_off[h] = off; L->blob_off[nb] = off; L->blob_cnt[nb++] = (int64_t)L->head_out[h] * in;
    off += (int64_t)L->head_out[h] * in;
    L->hb_off[h] = off; L->blob_off[nb] = off; L->blob_cnt[nb++] = L->head_out[h];
    off += L->head_out[h];
  }
  L->count = off;
  L->n_blobs = nb;
}

int64_t dqo_actor_param_count(const dqo_config *cfg) { net_layout L; make_layout(cfg, 0, &L); return L.count; }
int64_t dqo_critic_param_count(const dqo_config *cfg) { net_layout L; make_layout(cfg, 1, &L); return L.count; }

/* ------------------------------------------------------------------------- */
/* Layers                                                                      */
/* ------------------------------------------------------------------------- */

#define NEG_SLOPE 0.01f /* dqn.cpp:300 relu_param->set_negative_slope(0.01) */

/* InnerProduct forward: top = bottom * W^T ; top += ones * b^T */
static void ip_forward(int ub, int M, int N, int K, const float *X, const float *W, const float *b,
                       float *Y) {
  gemm_nt(ub, M, N, K, X, W, Y, 0);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) Y[(int64_t)m * N + n] += b[n];
}
/* ReLU (leaky) forward in place: max(x,0) + slope*min(x,0) */
static void lrelu_forward(int64_t n, float *y) {
  for (int64_t i = 0; i < n; ++i) {
    const float x = y[i];
    y[i] = (x > 0.f ? x : 0.f) + NEG_SLOPE * (x < 0.f ? x : 0.f);
  }
}
/* ReLU backward in place on the diff; "bottom_data" is the in-place (post-activation) blob */
static void lrelu_backward(int64_t n, const float *y, float *d) {
  for (int64_t i = 0; i < n; ++i)
    d[i] = d[i] * ((y[i] > 0.f ? 1.f : 0.f) + NEG_SLOPE * (y[i] <= 0.f ? 1.f : 0.f));
}
/* InnerProduct backward: dW += dY^T X ; db += dY^T 1 ; dX = dY W */
static void ip_backward(int ub, int M, int N, int K, const float *X, const float *W,
                        const float *dY, float *dW, float *db, float *dX) {
  if (dW) gemm_tn_acc(ub, N, K, M, dY, X, dW);
  if (db)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int m = 0; m < M; ++m) acc += dY[(int64_t)m * N + n];
      db[n] += acc;
    }
  if (dX) gemm_nn(ub, M, K, N, dY, W, dX);
}

typedef struct {
  float *act[DQO_MAX_HIDDEN]; /* post-activation outputs of each tower layer [M x H_l] */
  float *head[2];             /* linear head outputs [M x head_out] */
} net_acts;

static int alloc_acts(const net_layout *L, int M, net_acts *A) {
  memset(A, 0, sizeof(*A));
  int Mx = M > 0 ? M : 1;
  for (int l = 0; l < L->n_hidden; ++l)
    if (!(A->act[l] = (float *)malloc(sizeof(float) * (size_t)Mx * L->hidden[l]))) return -1;
  for (int h = 0; h < L->n_heads; ++h)
    if (!(A->head[h] = (float *)malloc(sizeof(float) * (size_t)Mx * L->head_out[h]))) return -1;
  return 0;
}
static void free_acts(net_acts *A) {
  for (int l = 0; l < DQO_MAX_HIDDEN; ++l) free(A->act[l]);
  free(A->head[0]); free(A->head[1]);
}

/* Net::ForwardPrefilled on the tower + heads (dqn.cpp:751, :1013) */
static void net_forward(int ub, const net_layout *L, const float *P, int M, const float *X,
                        net_acts *A) {
  const float *in = X;
  int K = L->in_dim;
  for (int l = 0; l < L->n_hidden; ++l) {
    ip_forward(ub, M, L->hidden[l], K, in, P + L->w_off[l], P + L->b_off[l], A->act[l]);
    lrelu_forward((int64_t)M * L->hidden[l], A->act[l]);
    in = A->act[l];
    K = L->hidden[l];
  }
  for (int h = 0; h < L->n_heads; ++h) /* heads are linear: no activation (dqn.cpp:426-427, :450) */
    ip_forward(ub, M, L->head_out[h], K, in, P + L->hw_off[h], P + L->hb_off[h], A->head[h]);
}

/* Net::BackwardFrom(head layer) down to the data layers (dqn.cpp:923, :963).
 * head_diff[h] is the top diff of head h.  G accumulates (Caffe accumulates into diffs).
 * want_dw: accumulate parameter gradients.  dX0 (may be NULL): gradient w.r.t. the input. */
static int net_backward(int ub, const net_layout *L, const float *P, float *G, int M,
                        const float *X, const net_acts *A, float *const head_diff[2],
                        int want_dw, float *dX0) {
  const int top = L->n_hidden - 1;
  const int Ht = L->hidden[top];
  float *d = (float *)malloc(sizeof(float) * (size_t)M * Ht);
  float *d2 = NULL;
  if (!d) return -1;
  /* heads, highest layer index first; the auto-inserted Split sums the bottom diffs */
  for (int h = L->n_heads - 1; h >= 0; --h) {
    float *dst = d;
    if (h != L->n_heads - 1) {
      if (!d2 && !(d2 = (float *)malloc(sizeof(float) * (size_t)M * Ht))) { free(d); return -1; }
      dst = d2;
    }
    ip_backward(ub, M, L->head_out[h], Ht, A->act[top], P + L->hw_off[h], head_diff[h],
                want_dw ? G + L->hw_off[h] : NULL, want_dw ? G + L->hb_off[h] : NULL, dst);
  }
  if (L->n_heads == 2) /* SplitLayer::Backward: caffe_add(top0.diff, top1.diff) */
    for (int64_t i = 0; i < (int64_t)M * Ht; ++i) d[i] = d2[i] + d[i];
  free(d2);
  for (int l = top; l >= 0; --l) {
    const int N = L->hidden[l];
    const int K = l > 0 ? L->hidden[l - 1] : L->in_dim;
    const float *in = l > 0 ? A->act[l - 1] : X;
    lrelu_backward((int64_t)M * N, A->act[l], d);
    float *dx = NULL;
    if (l > 0) dx = (float *)malloc(sizeof(float) * (size_t)M * K);
    else dx = dX0;
    if (l > 0 && !dx) { free(d); return -1; }
    ip_backward(ub, M, N, K, in, P + L->w_off[l], d, want_dw ? G + L->w_off[l] : NULL,
                want_dw ? G + L->b_off[l] : NULL, dx);
    free(d);
    d = l > 0 ? dx : NULL;
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Solver (Caffe SGDSolver::ApplyUpdate with AdamSolver::ComputeUpdateValue)   */
/* ------------------------------------------------------------------------- */

static float solver_apply(const dqo_config *cfg, const net_layout *L, float lr, float *P,
                          float *G, float *Mo, float *Vo, int32_t iter) {
  /* ClipGradients: per-blob sumsq_diff (sdot) summed in Dtype, global L2 norm */
  float sumsq = 0.f;
  for (int b = 0; b < L->n_blobs; ++b) sumsq += sdot(L->blob_cnt[b], G + L->blob_off[b], G + L->blob_off[b]);
  const float l2 = sqrtf(sumsq);
  if (cfg->clip_gradients >= 0.f && l2 > cfg->clip_gradients) {
    const float scale = cfg->clip_gradients / l2;
    for (int64_t i = 0; i < L->count; ++i) G[i] *= scale; /* Blob::scale_diff */
  }
  /* Adam: t = iter+1; correction in double (std::pow(float,int) promotes), narrowed to Dtype */
  const float beta1 = cfg->momentum, beta2 = cfg->momentum2;
  const int t = iter + 1;
  const float correction =
      (float)(sqrt((double)1.f - pow((double)beta2, t)) / ((double)1.f - pow((double)beta1, t)));
  const float local_rate = lr * 1.f; /* lr_mult = 1 for every blob */
  const float eps = cfg->delta;
  const float a1 = 1.f - beta1, a2 = 1.f - beta2;
  const float step = local_rate * correction;
  for (int64_t i = 0; i < L->count; ++i) {
    const float g = G[i];
    float m = Mo[i] * beta1;  /* caffe_cpu_axpby = sscal(beta) then saxpy(alpha) */
    m = m + a1 * g;
    const float gg = g * g;   /* caffe_mul */
    float v = Vo[i] * beta2;
    v = v + a2 * gg;
    Mo[i] = m; Vo[i] = v;
    float den = sqrtf(v);     /* caffe_powx(v, 0.5) */
    den = den + eps;          /* caffe_add_scalar */
    const float q = m / den;  /* caffe_div */
    const float d = step * q; /* caffe_cpu_scale -> diff */
    G[i] = d;
    P[i] = P[i] - d;          /* Net::Update -> Blob::Update: data -= diff */
  }
  return l2;
}

/* ------------------------------------------------------------------------- */
/* Public pieces                                                               */
/* ------------------------------------------------------------------------- */

void dqo_soft_update(int64_t n, float tau, const float *from, float *to) {
  /* dqn.cpp:1093 caffe_cpu_axpby(N, tau, from, (1-tau), to): sscal then saxpy */
  const float keep = 1 - tau;
  for (int64_t i = 0; i < n; ++i) {
    float t = to[i] * keep;
    to[i] = t + tau * from[i];
  }
}

void dqo_label_transitions(int32_t n, double gamma, const float *reward, float *mc) {
  /* dqn.cpp:783-797: G_T = r_T ; G_t = r_t + gamma_ * G_{t+1} (double gamma_, float store) */
  if (n <= 0) return;
  mc[n - 1] = reward[n - 1];
  for (int i = n - 2; i >= 0; --i) mc[i] = (float)((double)reward[i] + gamma * (double)mc[i + 1]);
}

void dqo_invert_gradients(int32_t n, const float *a_pi, float *d10) {
  /* dqn.cpp:927-957 */
  for (int i = 0; i < n; ++i) {
    for (int h = 0; h < DQO_ACTION_SIZE; ++h) {
      float diff = d10[i * DQO_ACTOR_OUT + h];
      const float output = a_pi[i * DQO_ACTOR_OUT + h];
      const float min = -1.0f, max = 1.0f;
      if (diff < 0) diff *= (max - output) / (max - min);
      else if (diff > 0) diff *= (output - min) / (max - min);
      d10[i * DQO_ACTOR_OUT + h] = diff;
    }
    for (int h = 0; h < DQO_ACTION_PARAM_SIZE; ++h) {
      float diff = d10[i * DQO_ACTOR_OUT + DQO_ACTION_SIZE + h];
      const float output = a_pi[i * DQO_ACTOR_OUT + DQO_ACTION_SIZE + h];
      float min, max;
      if (h == 0 || h == 4) { min = 0; max = 100; }
      else { min = -180; max = 180; }
      if (diff < 0) diff *= (max - output) / (max - min);
      else if (diff > 0) diff *= (output - min) / (max - min);
      d10[i * DQO_ACTOR_OUT + DQO_ACTION_SIZE + h] = diff;
    }
  }
}

int32_t dqo_replay_after_add_one(int32_t size, int32_t capacity, int32_t *new_size) {
  /* dqn.cpp:768-773 */
  int32_t pops = 0;
  if (size == capacity) { size--; pops++; }
  *new_size = size + 1;
  return pops;
}
int32_t dqo_replay_after_add_many(int32_t size, int32_t capacity, int32_t n, int32_t *new_size) {
  /* dqn.cpp:775-781: while (size + n >= capacity) pop_front(); (pop on empty is UB upstream;
   * the model stops at 0) */
  int32_t pops = 0;
  while (size + n >= capacity && size > 0) { size--; pops++; }
  *new_size = size + n;
  return pops;
}

int32_t dqo_get_action(const float *o, float *arg1, float *arg2) {
  /* dqn.cpp:196-208 with GetParamOffset dqn.cpp:162-178; DASH=0 TURN=1 TACKLE=2 KICK=3 */
  float c[DQO_ACTION_SIZE];
  for (int i = 0; i < DQO_ACTION_SIZE; ++i) c[i] = o[i];
  c[2] = -99999.f;
  int best = 0;
  for (int i = 1; i < DQO_ACTION_SIZE; ++i) if (c[i] > c[best]) best = i; /* max_element: first max */
  static const int off1[4] = {0, 2, 3, 4};
  static const int off2[4] = {1, -1, -1, 5};
  *arg1 = o[DQO_ACTION_SIZE + off1[best]];
  *arg2 = off2[best] < 0 ? 0.f : o[DQO_ACTION_SIZE + off2[best]];
  return best;
}

int dqo_actor_forward(const dqo_config *cfg, const float *actor, int32_t n, const float *states,
                      float *out10) {
  net_layout L; make_layout(cfg, 0, &L);
  net_acts A;
  if (alloc_acts(&L, n, &A)) { free_acts(&A); return -1; }
  net_forward(cfg->use_blas, &L, actor, n, states, &A);
  for (int i = 0; i < n; ++i) { /* dqn.cpp:755-764 */
    for (int c = 0; c < DQO_ACTION_SIZE; ++c) out10[i * DQO_ACTOR_OUT + c] = A.head[0][i * DQO_ACTION_SIZE + c];
    for (int c = 0; c < DQO_ACTION_PARAM_SIZE; ++c)
      out10[i * DQO_ACTOR_OUT + DQO_ACTION_SIZE + c] = A.head[1][i * DQO_ACTION_PARAM_SIZE + c];
  }
  free_acts(&A);
  return 0;
}

static void concat_input(const dqo_config *cfg, int n, const float *s, const float *a10, float *x) {
  const int S = cfg->state_size, K = S + DQO_ACTOR_OUT;
  for (int i = 0; i < n; ++i) {
    memcpy(x + (int64_t)i * K, s + (int64_t)i * S, sizeof(float) * S);
    memcpy(x + (int64_t)i * K + S, a10 + (int64_t)i * DQO_ACTOR_OUT, sizeof(float) * DQO_ACTOR_OUT);
  }
}

int dqo_critic_forward(const dqo_config *cfg, const float *critic, int32_t n, const float *states,
                       const float *act10, float *q) {
  net_layout L; make_layout(cfg, 1, &L);
  net_acts A;
  float *x = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1) * L.in_dim);
  if (!x || alloc_acts(&L, n, &A)) { free(x); return -1; }
  concat_input(cfg, n, states, act10, x);
  net_forward(cfg->use_blas, &L, critic, n, x, &A);
  for (int i = 0; i < n; ++i) q[i] = A.head[0][i];
  free_acts(&A); free(x);
  return 0;
}

int dqo_update(const dqo_config *cfg, float *actor, float *critic, float *actor_target,
               float *critic_target, float *actor_m, float *actor_v, float *critic_m,
               float *critic_v, int32_t *actor_iter, int32_t *critic_iter, const float *s,
               const float *act10, const float *reward, const float *mc_target,
               const uint8_t *term, const float *s_next, float *critic_loss, float *avg_q,
               const dqo_taps *taps) {
  const int B = cfg->batch, S = cfg->state_size, ub = cfg->use_blas;
  net_layout LA, LC;
  make_layout(cfg, 0, &LA);
  make_layout(cfg, 1, &LC);
  int rc = -1;
  float *sn = (float *)malloc(sizeof(float) * (size_t)B * S);
  float *a_next = (float *)malloc(sizeof(float) * (size_t)B * DQO_ACTOR_OUT);
  float *q_next = (float *)malloc(sizeof(float) * (size_t)B);
  float *y = (float *)malloc(sizeof(float) * (size_t)B);
  float *xc = (float *)malloc(sizeof(float) * (size_t)B * LC.in_dim);
  float *dq = (float *)malloc(sizeof(float) * (size_t)B);
  float *gC = (float *)calloc((size_t)LC.count, sizeof(float));
  float *gA = (float *)calloc((size_t)LA.count, sizeof(float));
  float *a_pi = (float *)malloc(sizeof(float) * (size_t)B * DQO_ACTOR_OUT);
  float *dX0 = (float *)malloc(sizeof(float) * (size_t)B * LC.in_dim);
  float *d10 = (float *)malloc(sizeof(float) * (size_t)B * DQO_ACTOR_OUT);
  float *dact = (float *)malloc(sizeof(float) * (size_t)B * DQO_ACTION_SIZE);
  float *dpar = (float *)malloc(sizeof(float) * (size_t)B * DQO_ACTION_PARAM_SIZE);
  float *dXa = cfg->caffe_wasted_work ? (float *)malloc(sizeof(float) * (size_t)B * S) : NULL;
  net_acts AC, AA;
  memset(&AC, 0, sizeof(AC)); memset(&AA, 0, sizeof(AA));
  if (!sn || !a_next || !q_next || !y || !xc || !dq || !gC || !gA || !a_pi || !dX0 || !d10 ||
      !dact || !dpar)
    goto done;
  if (alloc_acts(&LC, B, &AC) || alloc_acts(&LA, B, &AA)) goto done;

  /* dqn.cpp:879-886: compacted list of next states of non-terminal transitions */
  int nn = 0;
  for (int n = 0; n < B; ++n)
    if (!term[n]) memcpy(sn + (int64_t)(nn++) * S, s_next + (int64_t)n * S, sizeof(float) * S);

  /* dqn.cpp:889-891: CriticForwardThroughActor(critic_target, actor_target, next_states) */
  if (dqo_actor_forward(cfg, actor_target, nn, sn, a_next)) goto done;
  if (dqo_critic_forward(cfg, critic_target, nn, sn, a_next, q_next)) goto done;

  /* dqn.cpp:892-900 (double arithmetic, two narrowings to float) */
  {
    int k = 0;
    for (int n = 0; n < B; ++n) {
      float qn = 0.f;
      float off;
      if (term[n]) off = (float)(double)reward[n];
      else { qn = q_next[k++]; off = (float)((double)reward[n] + cfg->gamma * (double)qn); }
      const float on = mc_target[n];
      y[n] = (float)(cfg->beta * (double)on + (1 - cfg->beta) * (double)off);
      if (taps && taps->q_next) taps->q_next[n] = qn;
    }
  }
  if (taps && taps->y) memcpy(taps->y, y, sizeof(float) * B);

  /* dqn.cpp:901-904: critic_solver_->Step(1) = ClearParamDiffs, ForwardBackward, ApplyUpdate, ++iter */
  concat_input(cfg, B, s, act10, xc);
  net_forward(ub, &LC, critic, B, xc, &AC);
  {
    /* EuclideanLoss: diff = q - y ; loss = dot(diff,diff)/num/2 ; bottom_diff = diff/num */
    float *diff = dq;
    for (int n = 0; n < B; ++n) diff[n] = AC.head[0][n] - y[n];
    const float dot = sdot(B, diff, diff);
    *critic_loss = dot / (float)B / 2.f; /* dqn.cpp:905 */
    if (taps && taps->q) memcpy(taps->q, AC.head[0], sizeof(float) * B);
    const float alpha = 1.f / (float)B;
    for (int n = 0; n < B; ++n) dq[n] = alpha * diff[n];
  }
  {
    float *hd[2] = {dq, NULL};
    /* force_backward (dqn.cpp:434) makes Caffe also compute the layer-1 bottom diff */
    if (net_backward(ub, &LC, critic, gC, B, xc, &AC, hd, 1, cfg->caffe_wasted_work ? dX0 : NULL))
      goto done;
  }
  if (taps && taps->critic_grad) memcpy(taps->critic_grad, gC, sizeof(float) * (size_t)LC.count);
  {
    float l2 = solver_apply(cfg, &LC, cfg->critic_lr, critic, gC, critic_m, critic_v, *critic_iter);
    if (taps && taps->critic_gnorm) *taps->critic_gnorm = l2;
    *critic_iter += 1;
  }

  /* dqn.cpp:908-909: ZeroGradParameters on both nets */
  memset(gC, 0, sizeof(float) * (size_t)LC.count);
  memset(gA, 0, sizeof(float) * (size_t)LA.count);

  /* dqn.cpp:910-916: a_pi = actor(s); q = critic(s, a_pi) with the updated critic */
  net_forward(ub, &LA, actor, B, s, &AA);
  for (int n = 0; n < B; ++n) {
    for (int c = 0; c < DQO_ACTION_SIZE; ++c) a_pi[n * DQO_ACTOR_OUT + c] = AA.head[0][n * DQO_ACTION_SIZE + c];
    for (int c = 0; c < DQO_ACTION_PARAM_SIZE; ++c)
      a_pi[n * DQO_ACTOR_OUT + DQO_ACTION_SIZE + c] = AA.head[1][n * DQO_ACTION_PARAM_SIZE + c];
  }
  if (taps && taps->a_pi) memcpy(taps->a_pi, a_pi, sizeof(float) * (size_t)B * DQO_ACTOR_OUT);
  concat_input(cfg, B, s, a_pi, xc);
  net_forward(ub, &LC, critic, B, xc, &AC);
  {
    double acc = 0.0; /* std::accumulate(..., 0.0) / float(size) */
    for (int n = 0; n < B; ++n) acc += (double)AC.head[0][n];
    *avg_q = (float)(acc / (double)(float)B);
    if (taps && taps->q_pi) memcpy(taps->q_pi, AC.head[0], sizeof(float) * B);
  }

  /* dqn.cpp:918-923: q.diff = -1 ; critic.BackwardFrom(q_values_layer) */
  for (int n = 0; n < B; ++n) dq[n] = -1.0f;
  {
    float *hd[2] = {dq, NULL};
    if (net_backward(ub, &LC, critic, gC, B, xc, &AC, hd, cfg->caffe_wasted_work, dX0)) goto done;
  }
  /* Concat backward: slice the input diff into actions / action_params */
  for (int n = 0; n < B; ++n)
    memcpy(d10 + (int64_t)n * DQO_ACTOR_OUT, dX0 + (int64_t)n * LC.in_dim + S,
           sizeof(float) * DQO_ACTOR_OUT);
  if (taps && taps->d_raw) memcpy(taps->d_raw, d10, sizeof(float) * (size_t)B * DQO_ACTOR_OUT);

  /* dqn.cpp:927-957 */
  dqo_invert_gradients(B, a_pi, d10);
  if (taps && taps->d_inv) memcpy(taps->d_inv, d10, sizeof(float) * (size_t)B * DQO_ACTOR_OUT);

  /* dqn.cpp:960-963: ShareDiff + actor.BackwardFrom("actionpara_layer") */
  for (int n = 0; n < B; ++n) {
    for (int c = 0; c < DQO_ACTION_SIZE; ++c) dact[n * DQO_ACTION_SIZE + c] = d10[n * DQO_ACTOR_OUT + c];
    for (int c = 0; c < DQO_ACTION_PARAM_SIZE; ++c)
      dpar[n * DQO_ACTION_PARAM_SIZE + c] = d10[n * DQO_ACTOR_OUT + DQO_ACTION_SIZE + c];
  }
  {
    float *hd[2] = {dact, dpar};
    if (net_backward(ub, &LA, actor, gA, B, s, &AA, hd, 1, dXa)) goto done;
  }
  if (taps && taps->actor_grad) memcpy(taps->actor_grad, gA, sizeof(float) * (size_t)LA.count);

  /* dqn.cpp:964-965: actor_solver_->ApplyUpdate(); set_iter(iter+1) */
  {
    float l2 = solver_apply(cfg, &LA, cfg->actor_lr, actor, gA, actor_m, actor_v, *actor_iter);
    if (taps && taps->actor_gnorm) *taps->actor_gnorm = l2;
    *actor_iter += 1;
  }

  /* dqn.cpp:967-970 */
  {
    const int mx = *actor_iter > *critic_iter ? *actor_iter : *critic_iter;
    if (cfg->soft_update_freq > 0 && mx % cfg->soft_update_freq == 0) {
      dqo_soft_update(LC.count, cfg->tau, critic, critic_target);
      dqo_soft_update(LA.count, cfg->tau, actor, actor_target);
    }
  }
  rc = 0;
done:
  free(sn); free(a_next); free(q_next); free(y); free(xc); free(dq); free(gC); free(gA);
  free(a_pi); free(dX0); free(d10); free(dact); free(dpar); free(dXa);
  free_acts(&AC); free_acts(&AA);
  return rc;
}
