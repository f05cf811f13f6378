// engine.cu — host side of libdqn_b200.so: owns the HBM-resident state of one dqn::DQN
// (reference src/dqn.hpp:183-193: replay memory, 4 nets, 2 solvers) and sequences the kernels of
// UpdateActorCritic (src/dqn.cpp:828-972) / SelectActionGreedily (:734-766) / CriticForward
// (:982-1020) on CUDA streams, replaying one captured CUDA graph per update.
#include "../../include/dqn_b200.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <random>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "gemm.cuh"

namespace dqnb {

std::string &last_error() {
  static thread_local std::string e;
  return e;
}

// ---------------------------------------------------------------------------------------------
// driver entry point for TMA descriptors (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Split-plane matrix [2][rows][ld] -> 3-D tensor map {ld (inner), rows, 2 planes}, fp32,
// 128-byte swizzle, box {32, box_rows, 2}; out-of-bounds elements read as zero.
static int make_tmap(CUtensorMap *tm, const float *base, int rows, int ld, long long plane_elems,
                     int box_rows, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) DQNB_FAIL("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4ull, (cuuint64_t)plane_elems * 4ull};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   // MN-major fp32 operands need the 32-byte-atom flavour of the 128B swizzle (gemm.cuh)
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DQNB_FAIL("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d ld=%d)", (int)r, rows, ld);
  return 0;
}

// Output of a GEMM as the TMA unit stores it: fp32 [d2][rows][ld] clipped to `cols` columns, box {32, 32, box_d2},
// 128-byte swizzle (the epilogue stages its boxes in that pattern, gemm.cuh).
static int make_tmap_out(CUtensorMap *tm, const float *base, int cols, int rows, int d2, int ld, long long d2_stride_elems,
                         int box_d2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) DQNB_FAIL("cuTensorMapEncodeTiled entry point not available");
  if (d2_stride_elems <= 0) d2_stride_elems = (long long)rows * ld;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4ull, (cuuint64_t)d2_stride_elems * 4ull};
  cuuint32_t box[3] = {32, 32, (cuuint32_t)box_d2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DQNB_FAIL("cuTensorMapEncodeTiled (output) failed with CUresult %d (rows=%d ld=%d d2=%d)", (int)r, rows, ld, d2);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen (only when world_size > 1)
// ---------------------------------------------------------------------------------------------
struct Id128 { char b[128]; };
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, /*ncclUniqueId by value: 128 bytes*/ Id128, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi &nccl() {
  static NcclApi api;
  if (!api.lib) {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (int (*)(void *))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void **, int, Id128, int))dlsym(api.lib, "ncclCommInitRank");
      api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
      api.CommDestroy = (int (*)(void *))dlsym(api.lib, "ncclCommDestroy");
      api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
    }
  }
  return api;
}
constexpr int kNcclFloat = 7, kNcclSum = 0;

// ---------------------------------------------------------------------------------------------
// net geometry: Caffe-order blobs <-> padded internal flat layout
// ---------------------------------------------------------------------------------------------
struct LayerGeom {
  int n_real, k_real;     // Caffe blob W [n_real x k_real]
  int Np, Kp;             // padded
  long long w_off, b_off; // internal offsets
  long long cw_off, cb_off;  // Caffe-order offsets
};
struct NetGeom {
  bool critic;
  int n_hidden;
  LayerGeom L[DQNB_MAX_HIDDEN];
  int head_real;          // 10 (actor: action_layer 4 + actionpara_layer 6) or 1 (q_values_layer)
  int Hp;                 // padded width of the top tower layer
  long long hw_off, hb_off;   // internal: head W [16 x Hp], b [16]
  long long chw_off[2], chb_off[2];
  long long flat;         // internal element count (multiple of 1024)
  long long caffe_count;
  int in_real, in_pad;
};

static void make_geom(const dqnb_config &c, bool critic, NetGeom *g) {
  memset(g, 0, sizeof(*g));
  g->critic = critic;
  g->n_hidden = c.n_hidden;
  g->in_real = c.state_size + (critic ? kActorOut : 0);
  g->in_pad = round_up(g->in_real, 64);
  long long off = 0, coff = 0;
  int k_real = g->in_real, Kp = g->in_pad;
  for (int l = 0; l < c.n_hidden; ++l) {
    LayerGeom &L = g->L[l];
    L.n_real = c.hidden[l]; L.k_real = k_real;
    L.Np = round_up(c.hidden[l], 64); L.Kp = Kp;
    L.w_off = off; off += (long long)L.Np * L.Kp;
    L.b_off = off; off += L.Np;
    L.cw_off = coff; coff += (long long)L.n_real * L.k_real;
    L.cb_off = coff; coff += L.n_real;
    k_real = L.n_real; Kp = L.Np;
  }
  g->Hp = Kp;
  g->head_real = critic ? 1 : kActorOut;
  g->hw_off = off; off += 16LL * Kp;
  g->hb_off = off; off += 16;
  if (critic) {
    g->chw_off[0] = coff; coff += k_real;
    g->chb_off[0] = coff; coff += 1;
  } else {
    g->chw_off[0] = coff; coff += 4LL * k_real;
    g->chb_off[0] = coff; coff += 4;
    g->chw_off[1] = coff; coff += 6LL * k_real;
    g->chb_off[1] = coff; coff += 6;
  }
  g->flat = (off + 1023) / 1024 * 1024;
  g->caffe_count = coff;
}

// Caffe order -> internal padded (zero padding); head rows: actor = action_layer rows 0..3 then
// actionpara_layer rows 4..9 of one [16 x Hp] matrix.
static void caffe_to_internal(const NetGeom &g, const float *c, std::vector<float> &out) {
  out.assign((size_t)g.flat, 0.f);
  for (int l = 0; l < g.n_hidden; ++l) {
    const LayerGeom &L = g.L[l];
    for (int n = 0; n < L.n_real; ++n) {
      memcpy(&out[L.w_off + (long long)n * L.Kp], c + L.cw_off + (long long)n * L.k_real, sizeof(float) * L.k_real);
      out[L.b_off + n] = c[L.cb_off + n];
    }
  }
  const int k_real = g.L[g.n_hidden - 1].n_real;
  if (g.critic) {
    memcpy(&out[g.hw_off], c + g.chw_off[0], sizeof(float) * k_real);
    out[g.hb_off] = c[g.chb_off[0]];
  } else {
    for (int j = 0; j < 4; ++j) {
      memcpy(&out[g.hw_off + (long long)j * g.Hp], c + g.chw_off[0] + (long long)j * k_real, sizeof(float) * k_real);
      out[g.hb_off + j] = c[g.chb_off[0] + j];
    }
    for (int j = 0; j < 6; ++j) {
      memcpy(&out[g.hw_off + (long long)(4 + j) * g.Hp], c + g.chw_off[1] + (long long)j * k_real, sizeof(float) * k_real);
      out[g.hb_off + 4 + j] = c[g.chb_off[1] + j];
    }
  }
}
static void internal_to_caffe(const NetGeom &g, const float *in, float *c) {
  for (int l = 0; l < g.n_hidden; ++l) {
    const LayerGeom &L = g.L[l];
    for (int n = 0; n < L.n_real; ++n) {
      memcpy(c + L.cw_off + (long long)n * L.k_real, in + L.w_off + (long long)n * L.Kp, sizeof(float) * L.k_real);
      c[L.cb_off + n] = in[L.b_off + n];
    }
  }
  const int k_real = g.L[g.n_hidden - 1].n_real;
  if (g.critic) {
    memcpy(c + g.chw_off[0], in + g.hw_off, sizeof(float) * k_real);
    c[g.chb_off[0]] = in[g.hb_off];
  } else {
    for (int j = 0; j < 4; ++j) {
      memcpy(c + g.chw_off[0] + (long long)j * k_real, in + g.hw_off + (long long)j * g.Hp, sizeof(float) * k_real);
      c[g.chb_off[0] + j] = in[g.hb_off + j];
    }
    for (int j = 0; j < 6; ++j) {
      memcpy(c + g.chw_off[1] + (long long)j * k_real, in + g.hw_off + (long long)(4 + j) * g.Hp, sizeof(float) * k_real);
      c[g.chb_off[1] + j] = in[g.hb_off + 4 + j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
constexpr int kMaxSide = 2 + DQNB_MAX_HIDDEN;

struct SplitMat {           // [2][rows][ld] fp32 in HBM
  float *p = nullptr;
  uint32_t *bits = nullptr; // optional [rows][ld/32]: ReLU sign bits of a saved activation (gemm.cuh relu_bits_*)
  int rows = 0, ld = 0;
  long long plane() const { return (long long)rows * ld; }
};

struct Op {                 // one kernel launch of the update / act sequence
  enum Kind { GEMM, GATHER, SAMPLE, HEAD_FWD, CRITIC_HEAD, ACTOR_HEAD_BWD, HEAD_BWD_W, COLSUM, REDUCE,
              ALLREDUCE, P2P_ALLREDUCE, ADAM, PREP, FINALIZE, FORK, JOIN, ACT_STAGE, ACT_LAYER } kind;
  int branch = 0;           // 0 = main stream; 1, 2 = side streams between FORK and JOIN
  int wait_ev = -1;         // event the op's stream waits for before the launch (cross-branch edge)
  int rec_ev = -1;          // event recorded on the op's stream after the launch
  int mask = 3;             // FORK / JOIN: which side streams take part
  int variant = 0;          // 0: always; 1: only when indices are drawn on the device; 2: only when injected
  GemmArgs gemm; dim3 grid;
  GatherArgs gather; HeadArgs head; CriticHeadArgs ch; ActorHeadBwdArgs ahb; HeadBwdWArgs hbw; ColsumArgs cs;
  ReduceArgs red; AdamArgs adam; P2PArgs p2p; ActStageArgs ast; ActLayerArgs al;
  float *ar_buf = nullptr; size_t ar_count = 0;
  int blocks = 0;
};

}  // namespace dqnb

using namespace dqnb;

struct dqnb_handle_s {
  dqnb_config cfg;
  int S, Sp, Kc, B, Bp, An /*act rows pad*/;
  NetGeom gA, gC;
  cudaStream_t stream = nullptr;
  // side streams -> parallel branches of the captured graph (Op::branch b runs on side[b - 1]).  Branches 1, 2:
  // independent forward chains / head gradients / column sums; 3 + l: the weight-gradient GEMM of tower layer l
  // (each starts as soon as its dZ exists)
  cudaStream_t side[kMaxSide] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxSide] = {};
  cudaEvent_t evs[8] = {};                            // cross-branch edges inside one update
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<void *> allocs;
  std::vector<void *> pinned;
  // params: [2][flat]
  float *P[4] = {nullptr, nullptr, nullptr, nullptr};
  float *Mo[2] = {nullptr, nullptr}, *Vo[2] = {nullptr, nullptr};
  float *G[2] = {nullptr, nullptr};   // this rank's gradients [flat + 4] (sum of the split-K planes)
  float *Gr[2] = {nullptr, nullptr};  // gradient the optimiser consumes: G, or the all-reduced copy (P2P exchange)
  float *xchg = nullptr; long long xchg_floats = 0;   // IPC-exportable exchange allocation (world_size > 1)
  long long x_in[2] = {0, 0}, x_out[2] = {0, 0}, x_flag = 0;
  unsigned int *flag_ticket = nullptr;
  P2PTable *p2p_tab = nullptr; unsigned int *p2p_epoch = nullptr, *p2p_ticket = nullptr;
  int *p2p_err = nullptr; volatile int *h_p2p_err = nullptr;   // sticky exchange-failure flag: host-mapped pinned word
  int *p2p_err_dev = nullptr;                                  // ... and its device-resident twin, the one kernels READ
                                                               // (a read of mapped host memory costs ~1 us and they serialise)
  float *p2p_block_ss = nullptr;
  std::vector<void *> ipc_opened;
  int comm_mode = 0;                  // 0 none, 1 NCCL all-reduce, 2 P2P exchange kernel
  float *Gpart[2] = {nullptr, nullptr}; long long gpart_stride[2] = {0, 0};   // [actor, critic]
  // bias-gradient partials [planes][bflat]: plane = 128-row block of the minibatch (written by the dX epilogue that
  // produces the dZ tile) or row slice of colsum_kernel; boff[l] = offset of tower layer l inside a plane
  float *Bpart[2] = {nullptr, nullptr}; long long bflat = 0; int bplanes = 0; long long boff[DQNB_MAX_HIDDEN] = {};
  float *norm_part = nullptr; int n_norm[2] = {0, 0};
  double *scal_part = nullptr; int n_scal = 0;
  // replay ring
  float *ring = nullptr; int rw = 0;   // replay ring [cap][rw]: state Sp | next state Sp | act10, r, mc, terminal, pad
  int ring_head = 0, ring_size = 0;
  // minibatch buffers
  int32_t *idx = nullptr;
  SplitMat Xs, Xsn, Xc, Xct, Xcp;
  float *reward = nullptr, *mc = nullptr, *term = nullptr, *y = nullptr;
  float *q_next = nullptr, *q = nullptr, *q_pi = nullptr;
  float *a16_t = nullptr, *a16_pi = nullptr, *d16c = nullptr, *d16a = nullptr;
  float *d_in = nullptr, *tap_raw = nullptr, *tap_inv = nullptr;
  SplitMat actAT[DQNB_MAX_HIDDEN], actCT[DQNB_MAX_HIDDEN], actC[DQNB_MAX_HIDDEN], actA[DQNB_MAX_HIDDEN], dZ[DQNB_MAX_HIDDEN];
  // act path
  SplitMat Xact, Xeval, actE[DQNB_MAX_HIDDEN];
  float *out16_act = nullptr;
  float *h_act_in = nullptr, *h_act_out = nullptr;   // pinned staging
  // skinny act path (kernels.cuh act_*): own stream + graph, fp32 actor snapshots, host-mapped I/O
  cudaStream_t astream = nullptr;
  cudaGraphExec_t act_graph = nullptr;
  float *snapA = nullptr;                       // [2][gA.flat]
  unsigned int *act_cur = nullptr, *act_sel = nullptr;
  ActCtl *h_ctl = nullptr, *d_ctl = nullptr;    // mapped pinned
  float *h_ax = nullptr, *d_ax = nullptr;       // pinned call block {n, seq, -, - | rows [kActMaxRows][Sp]} and its device copy
  float *h_ay = nullptr, *d_ay = nullptr;       // mapped pinned rows out [kActMaxRows][16]
  float *actY[2] = {nullptr, nullptr};
  unsigned int act_seq = 0; int act_rows_pending = 0; bool act_skinny_pending = false;
  int act_kernels = 0;
  // step state / results
  StepState *st = nullptr;
  float *results = nullptr; int max_slots = 4096;
  unsigned int *ticket = nullptr;
  float *h_results = nullptr;       // mapped pinned ring the last optimiser launch writes (critic_loss, avg_q) into
  unsigned long long host_step = 0;  // updates enqueued so far (mirrors StepState::step)
  // staging for replay appends
  // Replay appends travel on their own copy stream, beside the running update: pinned staging slots (ring
  // layout, padding stays zero; the slot's last 16 bytes carry {head, size}) -> H2D straight into the ring.
  // Ordering: a copy waits for the last enqueued gather (the only reader of the ring), the next gather waits
  // for the copy.
  static constexpr int kStageSlots = 2;
  float *h_stage[kStageSlots] = {nullptr, nullptr}; int stage_rows = 0; int stage_next = 0;
  cudaEvent_t ev_stage[kStageSlots] = {nullptr, nullptr};   // H2D copies out of a slot are done
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr, ev_gather = nullptr;
  bool copy_pending = false;          // the compute stream has not yet been ordered after the last append
  volatile unsigned long long *h_done = nullptr; unsigned long long *d_done = nullptr;   // mapped: last finished update
  // The minibatch inputs (gather outputs) exist twice: update t reads set t % 2 while the gather of update t + 1
  // fills the other one on the gather stream, beside the running update (the fields above are the set the op list
  // being built refers to).
  struct InputSet { SplitMat Xs, Xsn, Xc, Xct, Xcp; float *reward = nullptr, *mc = nullptr, *term = nullptr; int32_t *idx = nullptr; };
  InputSet in[2];
  cudaStream_t gstream = nullptr;                      // gathers (the only reader of the replay ring)
  cudaEvent_t ev_set_ready[2] = {nullptr, nullptr};    // gather into set s finished
  cudaEvent_t ev_set_free[2] = {nullptr, nullptr};     // the last update that read set s finished
  int last_set = 0;                                    // set of the most recent update (debug taps, peeks)
  // op lists + graphs (one per input set)
  std::vector<Op> update_ops_set[2], act_ops, eval_ops;
  std::vector<Op> &update_ops = update_ops_set[0];
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  int kernels_per_update = 0;
  int64_t launches = 0;
  void *comm = nullptr;
  long long *trace = nullptr; int trace_ops = 0;   // DQNB_TRACE=1: per-GEMM timeline stamps (gemm.cuh DQNB_STAMP)
  HyperParams hp;
  SegTable segs[2];
};

namespace dqnb {

template <typename T>
static int dalloc(dqnb_handle_s *h, T **p, size_t count, bool zero = true) {
  void *d = nullptr;
  DQNB_CUDA(cudaMalloc(&d, count * sizeof(T)));
  if (zero) DQNB_CUDA(cudaMemset(d, 0, count * sizeof(T)));
  h->allocs.push_back(d);
  *p = (T *)d;
  return 0;
}
template <typename T>
static int halloc(dqnb_handle_s *h, T **p, size_t count) {
  void *d = nullptr;
  DQNB_CUDA(cudaMallocHost(&d, count * sizeof(T)));
  memset(d, 0, count * sizeof(T));
  h->pinned.push_back(d);
  *p = (T *)d;
  return 0;
}
static int alloc_mat(dqnb_handle_s *h, SplitMat *m, int rows, int ld, bool with_bits = false) {
  m->rows = rows; m->ld = ld;
  if (with_bits && dalloc(h, &m->bits, (size_t)rows * (ld / 32))) return -1;
  return dalloc(h, &m->p, (size_t)2 * rows * ld);
}

// ---------------------------------------------------------------------------------------------
// schedule / tile-shape knobs.  Defaults are the measured best (profiles/r01c_sched_sweep.txt); every knob
// can be overridden through the environment for experiments (scripts/sweep_sched.py).
// ---------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
struct Tuning {
  int sched;        // 0: three forward chains at once; 1: (target actor || critic) then (target critic || actor)
  int fuse_colsum;  // 1: bias gradients of the layers below the top one come from the dX epilogues (0: colsum launches)
  int fuse_tl;      // 1: TD target + critic loss head in one launch
  int gather_ahead; // 1: gathers run on a stream of their own into the idle input set, beside the previous update
  int actor_late;   // 1: the actor's forward on s (consumed only by the policy pass) starts with the critic's backward
                    //    pass instead of beside the target critic, which then runs alone with cluster split-K
  int cluster_b;    // 1: cluster split-K also for the target chains, which share the machine with the side chains
  int pdl_early;    // GEMM kernels: 1 = launch_dependents right after the wait, 0 = after the last MMA issue
  int side_delay;   // 1: a side chain of the forward phase starts when the critical chain's first (machine-filling) layer
                    //    of the same pair has finished, and the actor's chain joins only before its consumer
  int bn_fwd, bn_fwd_side, bn_dx, bn_dw;   // N tile (64 / 128) per GEMM class
  int ts_min_kb;        // forward / dX GEMMs with at least this many k-blocks per CTA run in A-in-TMEM mode (0 < x; 9999 = never)
  int bn32_max_tiles;   // forward / dX GEMMs with at most this many 128x64 output tiles use 128x32 tiles instead
  int bn_side_l1;   // N tile of a side chain's first layer: 128 halves its CTA count (128 -> 64), so that it fits beside
                    // the critical chain's second layer (64 CTAs) instead of queueing in front of it
  int st_fwd, st_fwd_side, st_dx, st_dw;   // smem ring depth per GEMM class (0 = deepest that fits)
  Tuning() {
    sched = env_int("DQNB_SCHED", 1);
    fuse_colsum = env_int("DQNB_FUSE_COLSUM", 1);
    fuse_tl = env_int("DQNB_FUSE_TL", 1);
    gather_ahead = env_int("DQNB_GATHER_AHEAD", 1);
    actor_late = env_int("DQNB_ACTOR_LATE", 0);
    cluster_b = env_int("DQNB_CLUSTER_B", 0);
    pdl_early = env_int("DQNB_PDL_EARLY", 0);
    side_delay = env_int("DQNB_SIDE_DELAY", 1);
    bn_fwd = env_int("DQNB_BN_FWD", 64);
    bn_fwd_side = env_int("DQNB_BN_FWD_SIDE", 64);
    ts_min_kb = env_int("DQNB_TS_MIN_KB", 8);
    bn32_max_tiles = env_int("DQNB_BN32_MAX_TILES", 32);
    bn_side_l1 = env_int("DQNB_BN_SIDE_L1", 128);
    bn_dx = env_int("DQNB_BN_DX", 64);
    bn_dw = env_int("DQNB_BN_DW", 128);   // 128-wide tiles, twice the splits: half the mainloop per CTA in the tail of a pass
    st_fwd = env_int("DQNB_ST_FWD", 0);
    st_fwd_side = env_int("DQNB_ST_FWD_SIDE", 0);
    st_dx = env_int("DQNB_ST_DX", 0);
    st_dw = env_int("DQNB_ST_DW", 2);   // two weight-gradient CTAs per SM: -2.4% per update (measured)
  }
};
static const Tuning &tuning() {
  static thread_local Tuning t;
  t = Tuning();           // re-read: the sweep script changes the environment between handles
  return t;
}

// ---------------------------------------------------------------------------------------------
// GEMM op builders
// ---------------------------------------------------------------------------------------------
static int pick_splits(int tiles, int kblocks, int max_splits) {
  int s = 1;
  while (s * 2 <= max_splits && tiles * s * 2 <= 160 && kblocks / (s * 2) >= 2) s *= 2;
  return s;
}

static int finish_gemm(const dqnb_config &cfg, Op *op) {
  GemmParams &p = op->gemm.p;
  op->kind = Op::GEMM;
  if (cfg.gemm_mode == DQNB_GEMM_TCGEN05_3XTF32) {
    // A
    const int a_rows = p.a_mn ? p.K : p.M, b_rows = p.b_mn ? p.K : p.N;
    if (p.bn != 32 && p.bn != 64 && p.bn != 128) p.bn = 64;
    // Layers with few output tiles (the narrow top of a tower and the dX GEMMs that feed it) run on 128x32 tiles:
    // twice the CTAs, each with a shorter mainloop per k-block (the A operand dominates the shared-memory reads) and
    // half the bytes to push through its SM's ~40 B/clk store path in the epilogue.
    bool narrow = false;
    if (p.epi != EPI_PLAIN && p.bn == 64 && p.N % 32 == 0 &&
        ((p.M + BM - 1) / BM) * ((p.N + 63) / 64) <= tuning().bn32_max_tiles) { p.bn = 32; narrow = true; }
    if (p.stages < 2 || p.stages > tc_max_stages(p.bn)) p.stages = tc_max_stages(p.bn);
    p.pdl_early = tuning().pdl_early;
    if (p.epi == EPI_DX && !p.relu_bits_in) DQNB_FAIL("EPI_DX needs the sign bits of the saved activation");
    // Forward / dX GEMMs with a long contraction but few output tiles: split K over a 2-CTA cluster
    // (DSMEM reduction in the epilogue) so that the dependent chain sees half the mainloop latency.
    {
      const int tiles = ((p.M + BM - 1) / BM) * ((p.N + p.bn - 1) / p.bn);
      if (p.epi != EPI_PLAIN && p.splits == 1 && p.K / BK >= 16 && 2 * tiles <= 148 && cfg.use_graph >= 0 &&
          !narrow &&
          !getenv("DQNB_NO_CLUSTER_SPLITK")) {
        p.cluster_k = 1;
        p.splits = 2;
      }
    }
    // long enough K-major-A mainloops read A from tensor memory (four mover warps copy it there, gemm.cuh)
    p.a_ts = (!p.a_mn && p.epi != EPI_PLAIN && (p.K / BK) / p.splits >= tuning().ts_min_kb) ? 1 : 0;
    if (make_tmap(&op->gemm.tmA, p.A, a_rows, p.lda, p.a_plane, p.a_mn ? 32 : BM, p.a_mn != 0)) return -1;
    if (make_tmap(&op->gemm.tmB, p.B, b_rows, p.ldb, p.b_plane, p.b_mn ? 32 : p.bn, p.b_mn != 0)) return -1;
    if (p.epi == EPI_PLAIN) {
      if (make_tmap_out(&op->gemm.tmD, p.out, p.N, p.M, p.splits, p.ldo, p.out_split_stride, 1)) return -1;
    } else {
      if (make_tmap_out(&op->gemm.tmD, p.out_hi, p.N, p.M, 2, p.ldo, (long long)(p.out_lo - p.out_hi), 2)) return -1;
    }
    op->grid = dim3((p.N + p.bn - 1) / p.bn, (p.M + BM - 1) / BM, p.splits);
  } else {
    op->grid = dim3((p.N + ST - 1) / ST, (p.M + ST - 1) / ST, p.splits);
  }
  return 0;
}

// forward of tower layer l: H = lrelu(X W^T + b)
static int op_fwd(const dqnb_config &cfg, const NetGeom &g, int l, const float *P, const SplitMat &X,
                  const SplitMat &H, Op *op, int bn = 64, int stages = 0) {
  const LayerGeom &L = g.L[l];
  GemmParams &p = op->gemm.p;
  memset(&p, 0, sizeof(p));
  p.bn = (bn == 128 && L.Np % 128 == 0) ? 128 : 64;
  p.stages = stages;
  p.M = X.rows; p.N = L.Np; p.K = L.Kp; p.a_mn = 0; p.b_mn = 0; p.splits = 1; p.epi = EPI_FWD;
  p.A = X.p; p.a_plane = X.plane(); p.lda = X.ld;
  p.B = P + L.w_off; p.b_plane = g.flat; p.ldb = L.Kp;
  p.out_hi = H.p; p.out_lo = H.p + H.plane(); p.ldo = H.ld;
  p.bias_hi = P + L.b_off; p.bias_lo = P + g.flat + L.b_off; p.apply_lrelu = 1;
  p.relu_bits_out = H.bits; p.ldbits = H.ld / 32;
  return finish_gemm(cfg, op);
}
// backward w.r.t. bottom of layer l (l >= 1): dZ_{l-1} = (dZ_l W_l) * relu'(H_{l-1})
static int op_dx(const dqnb_config &cfg, const NetGeom &g, int l, const float *P, const SplitMat &dZl,
                 const SplitMat &Hprev, const SplitMat &dZprev, Op *op) {
  const LayerGeom &L = g.L[l];
  GemmParams &p = op->gemm.p;
  memset(&p, 0, sizeof(p));
  p.bn = (tuning().bn_dx == 128 && L.Kp % 128 == 0) ? 128 : 64;
  p.stages = tuning().st_dx;
  p.M = dZl.rows; p.N = L.Kp; p.K = L.Np; p.a_mn = 0; p.b_mn = 1; p.splits = 1; p.epi = EPI_DX;
  p.A = dZl.p; p.a_plane = dZl.plane(); p.lda = dZl.ld;
  p.B = P + L.w_off; p.b_plane = g.flat; p.ldb = L.Kp;
  p.out_hi = dZprev.p; p.out_lo = dZprev.p + dZprev.plane(); p.ldo = dZprev.ld;
  p.mask_hi = Hprev.p; p.mask_lo = Hprev.p + Hprev.plane(); p.ldmask = Hprev.ld;
  p.relu_bits_in = Hprev.bits; p.ldbits = Hprev.ld / 32;
  return finish_gemm(cfg, op);
}
// input diff of layer 0 (critic policy pass): d_in = dZ_0 W_0, raw fp32
static int op_dx_plain(const dqnb_config &cfg, const NetGeom &g, const float *P, const SplitMat &dZ0,
                       float *d_in, long long split_stride, int *splits_out, Op *op) {
  const LayerGeom &L = g.L[0];
  GemmParams &p = op->gemm.p;
  memset(&p, 0, sizeof(p));
  p.M = dZ0.rows; p.N = L.Kp; p.K = L.Np; p.a_mn = 0; p.b_mn = 1; p.epi = EPI_PLAIN;
  // few output tiles (the critic input is narrow) but a long contraction: split-K, partial planes
  // are summed by the consumer (actor_head_bwd_kernel)
  p.splits = pick_splits(((p.M + BM - 1) / BM) * ((p.N + 63) / 64), p.K / BK, kGradSplits);
  *splits_out = p.splits;
  p.A = dZ0.p; p.a_plane = dZ0.plane(); p.lda = dZ0.ld;
  p.B = P + L.w_off; p.b_plane = g.flat; p.ldb = L.Kp;
  p.out = d_in; p.out_split_stride = split_stride; p.ldo = L.Kp;
  return finish_gemm(cfg, op);
}
// weight gradient of layer l: dW_l = dZ_l^T X_{l-1} (contraction over the minibatch, split-K)
static int op_dw(const dqnb_config &cfg, const NetGeom &g, int l, const SplitMat &dZl, const SplitMat &Xin,
                 float *gpart, long long gpart_stride, int *splits_out, Op *op) {
  const LayerGeom &L = g.L[l];
  GemmParams &p = op->gemm.p;
  memset(&p, 0, sizeof(p));
  p.M = L.Np; p.N = L.Kp; p.K = dZl.rows; p.a_mn = 1; p.b_mn = 1; p.epi = EPI_PLAIN;
  p.bn = (tuning().bn_dw == 128 && L.Kp % 128 == 0) ? 128 : 64;
  p.stages = tuning().st_dw;
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + p.bn - 1) / p.bn);
  p.splits = pick_splits(tiles, p.K / BK, kGradSplits);
  *splits_out = p.splits;
  p.A = dZl.p; p.a_plane = dZl.plane(); p.lda = dZl.ld;
  p.B = Xin.p; p.b_plane = Xin.plane(); p.ldb = Xin.ld;
  p.out = gpart + L.w_off; p.out_split_stride = gpart_stride; p.ldo = L.Kp;
  return finish_gemm(cfg, op);
}

static void op_head_fwd(const NetGeom &g, const float *P, const SplitMat &H, int rows, float *out16,
                        const SplitMat *dst, int dst_col, Op *op) {
  op->kind = Op::HEAD_FWD;
  HeadArgs &a = op->head;
  memset(&a, 0, sizeof(a));
  a.H = H.p; a.h_plane = H.plane(); a.ldh = H.ld; a.Kp = g.Hp;
  a.W = P + g.hw_off; a.w_plane = g.flat; a.bias = P + g.hb_off; a.b_plane = g.flat;
  a.J = g.head_real; a.rows = rows; a.out16 = out16;
  if (dst) { a.dst = dst->p; a.dst_plane = dst->plane(); a.ldd = dst->ld; a.dst_col = dst_col; }
  op->blocks = (rows + kHeadRowsPerBlock - 1) / kHeadRowsPerBlock;
}

}  // namespace dqnb

// ---------------------------------------------------------------------------------------------
// launching: every kernel goes out with the programmatic-stream-serialization attribute (PDL);
// the kernels call griddepcontrol.wait before touching global memory (kernels.cuh).
// ---------------------------------------------------------------------------------------------
static int g_cluster_z = 1;   // set around a launch that wants thread-block clusters of (1,1,z)
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (g_cluster_z > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 1; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = (unsigned)g_cluster_z;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

static int launch_op(dqnb_handle_s *h, const Op &op, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
  switch (op.kind) {
    case Op::GEMM:
      if (h->cfg.gemm_mode == DQNB_GEMM_TCGEN05_3XTF32) {
        g_cluster_z = op.gemm.p.cluster_k ? 2 : 1;
        TcKernel k = tc_kernel_for(op.gemm.p.a_mn, op.gemm.p.b_mn, op.gemm.p.bn, op.gemm.p.epi);
        if (!k) DQNB_FAIL("no tcgen05 kernel instance for operand layout (%d,%d) with epilogue %d", op.gemm.p.a_mn, op.gemm.p.b_mn, op.gemm.p.epi);
        e = launch_k(k, op.grid, dim3(TC_THREADS),
                     (size_t)tc_smem_for(op.gemm.p.bn, op.gemm.p.stages), s, op.gemm);
        g_cluster_z = 1;
      }
      else
        e = launch_k(gemm_simt_kernel, op.grid, dim3(256), 0, s, op.gemm.p);
      break;
    case Op::GATHER: e = launch_k(gather_kernel, dim3(h->Bp), dim3(128), 0, s, op.gather); break;
    case Op::SAMPLE: break;
    case Op::HEAD_FWD: e = launch_k(head_fwd_kernel, dim3(op.blocks), dim3(256), 0, s, op.head); break;
    case Op::CRITIC_HEAD: e = launch_k(critic_head_kernel, dim3(op.blocks), dim3(256), 0, s, op.ch); break;
    case Op::ACTOR_HEAD_BWD: e = launch_k(actor_head_bwd_kernel, dim3(op.blocks), dim3(256), 0, s, op.ahb); break;
    case Op::HEAD_BWD_W: e = launch_k(head_bwd_w_kernel, op.grid, dim3(256), 0, s, op.hbw); break;
    case Op::COLSUM: e = launch_k(colsum_kernel, op.grid, dim3(1024), 0, s, op.cs); break;
    case Op::REDUCE: e = launch_k(reduce_kernel, dim3(op.blocks), dim3(256), 0, s, op.red); break;
    case Op::P2P_ALLREDUCE: e = launch_k(p2p_allreduce_kernel, dim3(op.blocks), dim3(512), 0, s, op.p2p); break;
    case Op::ADAM: e = launch_k(adam_kernel, dim3(op.blocks), dim3(256), 0, s, op.adam); break;
    case Op::PREP:
    case Op::FINALIZE:
      return 0;
    case Op::ACT_STAGE: e = launch_k(act_stage_kernel, dim3(1), dim3(32), 0, s, op.ast); break;
    case Op::ACT_LAYER: e = launch_k(act_layer_kernel, dim3(op.blocks), dim3(op.al.last ? 512 : 256), 0, s, op.al); break;
    case Op::FORK:
    case Op::JOIN:
      return 0;   // stream plumbing, handled by run_ops
    case Op::ALLREDUCE: {
      if (!h->comm) DQNB_FAIL("world_size > 1 but dqnb_comm_init was not called");
      int r = nccl().AllReduce(op.ar_buf, op.ar_buf, op.ar_count, kNcclFloat, kNcclSum, h->comm, s);
      if (r != 0) DQNB_FAIL("ncclAllReduce failed: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
      return 0;
    }
  }
  DQNB_CUDA(e);
  return 0;
}

static int run_ops(dqnb_handle_s *h, const std::vector<Op> &ops, cudaStream_t s, bool skip_sample, int *count) {
  int n = 0;
  for (const Op &op : ops) {
    if (op.kind == Op::GATHER) continue;       // launched by enqueue_update ahead of the graph
    if (op.kind == Op::FORK) {          // side streams pick up after everything queued on the main one
      DQNB_CUDA(cudaEventRecord(h->ev_fork, s));
      for (int b = 0; b < kMaxSide; ++b)
        if (op.mask & (1 << b)) DQNB_CUDA(cudaStreamWaitEvent(h->side[b], h->ev_fork, 0));
      continue;
    }
    if (op.kind == Op::JOIN) {
      for (int b = 0; b < kMaxSide; ++b)
        if (op.mask & (1 << b)) {
          DQNB_CUDA(cudaEventRecord(h->ev_join[b], h->side[b]));
          DQNB_CUDA(cudaStreamWaitEvent(s, h->ev_join[b], 0));
        }
      continue;
    }
    cudaStream_t os = op.branch ? h->side[op.branch - 1] : s;
    if (op.wait_ev >= 0) DQNB_CUDA(cudaStreamWaitEvent(os, h->evs[op.wait_ev], 0));
    if (launch_op(h, op, os)) return -1;
    if (op.rec_ev >= 0) DQNB_CUDA(cudaEventRecord(h->evs[op.rec_ev], os));
    if (op.kind != Op::ALLREDUCE) ++n;
  }
  if (count) *count = n;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// building the op lists
// ---------------------------------------------------------------------------------------------
static int build_forward(dqnb_handle_s *h, const NetGeom &g, const float *P, const SplitMat &X,
                         SplitMat *acts, std::vector<Op> &ops, bool critical_chain = true, bool allow_cluster = true) {
  const SplitMat *in = &X;
  for (int l = 0; l < g.n_hidden; ++l) {
    Op op;
    if (op_fwd(h->cfg, g, l, P, *in, acts[l], &op, critical_chain ? tuning().bn_fwd : (l == 0 ? tuning().bn_side_l1 : tuning().bn_fwd_side),
               critical_chain ? tuning().st_fwd : tuning().st_fwd_side)) return -1;
    if ((!critical_chain || !allow_cluster) && op.gemm.p.cluster_k) {
      // side-branch passes run beside the critical chain: they should not grab twice the SMs for a
      // latency that nobody waits on, so they keep one CTA per tile
      op.gemm.p.cluster_k = 0; op.gemm.p.splits = 1; op.grid.z = 1;
    }
    ops.push_back(op);
    in = &acts[l];
  }
  return 0;
}

// bias gradients (column sums of dZ) of tower layers [l0, l1)
static Op make_colsum(dqnb_handle_s *h, const NetGeom &g, int l0, int l1) {
  Op op;
  op.kind = Op::COLSUM;
  ColsumArgs &a = op.cs;
  memset(&a, 0, sizeof(a));
  a.n_layers = l1 - l0; a.rows_pad = h->Bp;
  int blk = 0;
  for (int l = l0; l < l1; ++l) {
    const int i = l - l0;
    a.dZ[i] = h->dZ[l].p; a.plane[i] = h->dZ[l].plane(); a.ld[i] = h->dZ[l].ld; a.Np[i] = g.L[l].Np;
    a.b_off[i] = h->boff[l]; a.blk_begin[i] = blk; blk += (g.L[l].Np + kCsCols - 1) / kCsCols;
  }
  a.blk_begin[l1 - l0] = blk;
  a.gpart = h->Bpart[g.critic]; a.gpart_stride = h->bflat;
  op.grid = dim3(blk, kGradSplits);
  return op;
}

// tower backward from dZ[top] (already masked by the head backward).  The dX chain is the critical
// path and stays on the main stream.  With want_dw every weight-gradient GEMM runs on a side stream of its
// own (branch 3 + l) and the head gradient on branch 2, each gated by an event on the dZ it consumes, so a
// gradient starts the moment its dZ exists; all JOIN back before the reduction.  Bias gradients (column sums of
// dZ): the dX epilogue that produces dZ[l-1] also writes its per-row-block column sums (gemm.cuh); only the top
// layer, whose dZ comes from a head kernel, keeps a colsum launch (behind its weight gradient).
static int build_backward(dqnb_handle_s *h, const NetGeom &g, const float *P, const SplitMat &X,
                          SplitMat *acts, bool want_dw, SegTable *segs, const Op *head_bwd_w,
                          std::vector<Op> &ops, int extra_join_mask = 0) {
  const int top = g.n_hidden - 1;
  if (g.n_hidden + 1 > 8) DQNB_FAIL("too many layers for the event table");
  const bool fuse_cs = want_dw && h->cfg.gemm_mode == DQNB_GEMM_TCGEN05_3XTF32 && tuning().fuse_colsum;
  int fork_mask = 3;
  if (want_dw) for (int l = 0; l <= top; ++l) fork_mask |= 1 << (2 + l);
  if (want_dw) {
    // the side streams pick up once dZ[top] exists
    Op f; f.kind = Op::FORK; f.mask = fork_mask; ops.push_back(f);
    if (head_bwd_w) { Op w = *head_bwd_w; w.branch = 2; ops.push_back(w); }
  }
  for (int l = top; l >= 0; --l) {
    if (want_dw) {
      Op op;
      int splits = 1;
      if (op_dw(h->cfg, g, l, h->dZ[l], l > 0 ? acts[l - 1] : X, h->Gpart[g.critic], h->gpart_stride[g.critic], &splits, &op)) return -1;
      op.branch = 3 + l;
      if (l < top) op.wait_ev = l;                              // dZ[l] is produced by the dX op below
      ops.push_back(op);
      const bool fused_here = fuse_cs && l < top;
      if (!fused_here) {
        // bias column sums of dZ[l] by a launch of their own: behind the weight gradient that waits for the same dZ;
        // the last two (whose GEMMs form the tail of the pass) beside their GEMMs on branch 2, after the head gradient
        Op c = make_colsum(h, g, l, l + 1);
        if (l > 1 || l == top) c.branch = 3 + l; else { c.branch = 2; c.wait_ev = l; }
        ops.push_back(c);
      }
      // segment table entries (internal flat order: W_l then b_l)
      SegTable &T = *segs;
      T.begin[2 * l] = g.L[l].w_off; T.end[2 * l] = g.L[l].b_off; T.nsplit[2 * l] = splits;
      T.src[2 * l] = h->Gpart[g.critic] + g.L[l].w_off; T.stride[2 * l] = h->gpart_stride[g.critic];
      T.begin[2 * l + 1] = g.L[l].b_off; T.end[2 * l + 1] = g.L[l].b_off + g.L[l].Np;
      T.nsplit[2 * l + 1] = fused_here ? h->Bp / BM : kGradSplits;
      T.src[2 * l + 1] = h->Bpart[g.critic] + h->boff[l]; T.stride[2 * l + 1] = h->bflat;
    }
    if (l > 0) {
      Op op;
      if (op_dx(h->cfg, g, l, P, h->dZ[l], acts[l - 1], h->dZ[l - 1], &op)) return -1;
      if (want_dw) op.rec_ev = l - 1;                           // dZ[l-1] ready
      if (fuse_cs) { op.gemm.p.colsum_out = h->Bpart[g.critic] + h->boff[l - 1]; op.gemm.p.colsum_stride = h->bflat; }
      ops.push_back(op);
    }
  }
  if (want_dw) {
    SegTable &T = *segs;
    const int hs = 2 * g.n_hidden;
    T.begin[hs] = g.hw_off; T.end[hs] = g.flat; T.nsplit[hs] = kGradSplits;   // head W, b (+ zero tail)
    T.src[hs] = h->Gpart[g.critic] + g.hw_off; T.stride[hs] = h->gpart_stride[g.critic];
    T.n = hs + 1;
    Op j; j.kind = Op::JOIN; j.mask = fork_mask | extra_join_mask; ops.push_back(j);
  }
  return 0;
}

static void build_solver(dqnb_handle_s *h, int is_critic, const SegTable &segs, float scal_scale, std::vector<Op> &ops) {
  const NetGeom &g = is_critic ? h->gC : h->gA;
  const int blocks = (int)(g.flat / 1024);
  const bool multi = h->cfg.world_size > 1;
  Op r;
  r.kind = Op::REDUCE;
  ReduceArgs &a = r.red;
  memset(&a, 0, sizeof(a));
  a.segs = segs; a.flat = g.flat;
  a.G = h->G[is_critic]; a.norm_part = h->norm_part; a.scal_part = h->scal_part; a.n_scal = h->n_scal;
  a.scal_scale = scal_scale; a.do_reduce = 1; a.do_sumsq = multi ? 0 : 1;
  r.blocks = blocks;
  ops.push_back(r);
  if (multi) {
    Op ar;
    if (h->comm_mode == 2) {        // fused reduce-scatter + all-gather over NVLink peer memory
      ar.kind = Op::P2P_ALLREDUCE;
      P2PArgs &x = ar.p2p;
      memset(&x, 0, sizeof(x));
      x.tab = h->p2p_tab; x.world = h->cfg.world_size; x.rank = h->cfg.rank; x.net = is_critic;
      x.in_off = h->x_in[is_critic]; x.out_off = h->x_out[is_critic]; x.flag_off = h->x_flag;
      ReduceArgs &ra = ops.back().red;      // the reduction in front of the exchange releases flag A
      ra.tab = h->p2p_tab; ra.world = h->cfg.world_size; ra.rank = h->cfg.rank; ra.net = is_critic; ra.flag_off = h->x_flag;
      ra.epoch = h->p2p_epoch; ra.flag_ticket = h->flag_ticket;
      x.count = g.flat + 4; x.epoch = h->p2p_epoch; x.ticket = h->p2p_ticket; x.err = h->p2p_err; x.err_dev = h->p2p_err_dev;
      x.timeout_ns = (unsigned long long)std::max(1, env_int("DQNB_P2P_TIMEOUT_MS", 20000)) * 1000000ull;
      x.block_ss = h->p2p_block_ss;
      ar.blocks = 128;
      ops.push_back(ar);            // also delivers every rank's share of ||g||^2: no separate norm pass
    } else {                        // NCCL (validation alternative); launch fails loudly if no communicator
      ar.kind = Op::ALLREDUCE; ar.ar_buf = h->G[is_critic]; ar.ar_count = (size_t)g.flat + 4;
      ops.push_back(ar);
      Op r2 = r;
      r2.red.do_reduce = 0; r2.red.do_sumsq = 1; r2.red.G = h->Gr[is_critic];
      ops.push_back(r2);
    }
  }
  Op ad;
  ad.kind = Op::ADAM;
  AdamArgs &d = ad.adam;
  memset(&d, 0, sizeof(d));
  d.flat = g.flat; d.G = h->Gr[is_critic]; d.norm_part = h->norm_part; d.n_norm = blocks;
  if (multi && h->comm_mode == 2) {   // per-rank slice norms delivered by the exchange kernel
    d.norm_part = h->xchg + h->x_flag + 4 * kMaxPeers + is_critic * kMaxPeers; d.n_norm = h->cfg.world_size;
  }
  d.M = h->Mo[is_critic]; d.V = h->Vo[is_critic];
  d.P = h->P[is_critic ? DQNB_CRITIC : DQNB_ACTOR]; d.p_plane = g.flat;
  d.T = h->P[is_critic ? DQNB_CRITIC_TARGET : DQNB_ACTOR_TARGET]; d.t_plane = g.flat;
  d.st = h->st; d.st_out = h->st; d.is_critic = is_critic; d.hp = h->hp;
  d.comm_err = (multi && h->comm_mode == 2) ? h->p2p_err_dev : nullptr;
  ad.blocks = blocks;
  ops.push_back(ad);
}

// critic head (+ TD target | loss + head backward | policy seed + head backward), one fused kernel
static void push_critic_head(dqnb_handle_s *h, int mode, const float *P, const SplitMat &Htop, float *q_tap,
                             std::vector<Op> &ops) {
  const NetGeom &g = h->gC;
  Op op;
  op.kind = Op::CRITIC_HEAD;
  CriticHeadArgs &a = op.ch;
  memset(&a, 0, sizeof(a));
  a.mode = mode; a.B = h->B; a.rows_pad = h->Bp;
  a.H = Htop.p; a.h_plane = Htop.plane(); a.ldh = Htop.ld; a.Kp = g.Hp;
  a.W = P + g.hw_off; a.w_plane = g.flat; a.bias = P + g.hb_off; a.b_plane = g.flat;
  a.reward = h->reward; a.mc = h->mc; a.term = h->term; a.y = h->y; a.q_tap = q_tap; a.d16 = h->d16c;
  a.dZ = h->dZ[g.n_hidden - 1].p; a.dz_plane = h->dZ[g.n_hidden - 1].plane();
  a.part = h->scal_part; a.hp = h->hp;
  op.blocks = (h->Bp + kHeadRowsPerBlock - 1) / kHeadRowsPerBlock;
  ops.push_back(op);
}
static Op make_head_bwd_w(dqnb_handle_s *h, const NetGeom &g, const float *d16, const SplitMat &Htop) {
  Op w;
  w.kind = Op::HEAD_BWD_W;
  HeadBwdWArgs &b = w.hbw;
  memset(&b, 0, sizeof(b));
  b.d16 = d16; b.J = g.head_real; b.H = Htop.p; b.h_plane = Htop.plane(); b.ldh = Htop.ld; b.Kp = g.Hp;
  b.rows_pad = h->Bp; b.gpart = h->Gpart[g.critic]; b.gpart_stride = h->gpart_stride[g.critic]; b.hw_off = g.hw_off; b.hb_off = g.hb_off;
  w.grid = dim3((g.Hp + kHbwCols - 1) / kHbwCols, kGradSplits);
  return w;
}

static void select_input_set(dqnb_handle_s *h, int set) {
  const dqnb_handle_s::InputSet &I = h->in[set];
  h->Xs = I.Xs; h->Xsn = I.Xsn; h->Xc = I.Xc; h->Xct = I.Xct; h->Xcp = I.Xcp;
  h->reward = I.reward; h->mc = I.mc; h->term = I.term; h->idx = I.idx;
}

static int build_update_ops_for(dqnb_handle_s *h, int set) {
  select_input_set(h, set);
  std::vector<Op> &ops = h->update_ops_set[set];
  ops.clear();
  const NetGeom &gA = h->gA, &gC = h->gC;
  float *PA = h->P[DQNB_ACTOR], *PC = h->P[DQNB_CRITIC], *PAT = h->P[DQNB_ACTOR_TARGET], *PCT = h->P[DQNB_CRITIC_TARGET];
  const int topA = gA.n_hidden - 1, topC = gC.n_hidden - 1;
  Op op;
  // dqn.cpp:846-887: draw the minibatch (device sampler, or caller-injected indices) and gather it;
  // the same launch refreshes the per-update Adam scalars
  for (int variant = 1; variant <= 2; ++variant) {
    op.kind = Op::GATHER;
    op.variant = variant;
    GatherArgs &a = op.gather;
    memset(&a, 0, sizeof(a));
    a.st = h->st; a.idx = h->idx; a.sample = variant == 1; a.seed = h->cfg.seed; a.hp = h->hp;
    a.ring_s = h->ring; a.ring_sn = h->ring + h->Sp; a.ring_misc = h->ring + 2 * h->Sp; a.rw = h->rw;
    a.cap = h->cfg.replay_capacity; a.B = h->B; a.Bp = h->Bp; a.S = h->S; a.Sp = h->Sp; a.Kc = h->Kc;
    a.Xs = h->Xs.p; a.Xsn = h->Xsn.p; a.Xc = h->Xc.p; a.Xct = h->Xct.p; a.Xcp = h->Xcp.p;
    a.reward = h->reward; a.mc = h->mc; a.term = h->term;
    ops.push_back(op);
  }
  op.variant = 0;
  // Independent forward chains on separate streams (-> parallel graph branches):
  //   main  : dqn.cpp:889-891 CriticForwardThroughActor(critic_target, actor_target, s') + TD target
  //   side 1: forward half of critic_solver_->Step(1) on (s, a, p)            (dqn.cpp:904)
  //   side 2: actor forward on s with the pre-update actor                    (dqn.cpp:910-911)
  // sched 1 (default): only two chains compete for the SMs at any time: (target actor || critic), then
  // (target critic || actor).  The actor chain waits for the target actor's head (event 7) and is joined with the
  // critic's gradient branches, long before its consumer (the critic forward on (s, a_pi)) runs.
  // sched 0: all three from the start (the previous default; DQNB_SIDE_SERIAL=1 puts both side chains on one stream).
  const int sched = tuning().sched;
  const int side2 = getenv("DQNB_SIDE_SERIAL") ? 1 : 2;
  const int fork_mask = sched == 1 ? 1 : (side2 == 2 ? 3 : 1);
  constexpr int kEvActorStart = 7;
  const bool side_delay = sched == 1 && tuning().side_delay != 0 && !tuning().actor_late;
  constexpr int kEvCriticStart = 6;
  op.kind = Op::FORK; op.mask = fork_mask; ops.push_back(op);
  std::vector<Op> ta_ops;                           // target actor tower (main stream)
  if (build_forward(h, gA, PAT, h->Xsn, h->actAT, ta_ops, true, tuning().cluster_b != 0)) return -1;
  if (side_delay) {
    // The first layer of a tower fills the machine (128 CTAs) for a few microseconds.  Launched together, the side
    // chain's first layer takes the SMs the critical chain's next layer needs (measured: +5 us on each pair), so the
    // side chain starts when the critical chain's first layer is done.
    ta_ops[0].rec_ev = kEvCriticStart;
    ops.push_back(ta_ops[0]);
    ta_ops.erase(ta_ops.begin());
  }
  {
    const size_t mark = ops.size();
    if (build_forward(h, gC, PC, h->Xc, h->actC, ops, false)) return -1;
    for (size_t i = mark; i < ops.size(); ++i) ops[i].branch = 1;
    if (side_delay) ops[mark].wait_ev = kEvCriticStart;
  }
  auto push_actor_chain = [&](int branch, int wait_ev) -> int {
    const size_t mark = ops.size();
    if (build_forward(h, gA, PA, h->Xs, h->actA, ops, false)) return -1;
    Op hd;
    op_head_fwd(gA, PA, h->actA[topA], h->B, h->a16_pi, &h->Xcp, h->S, &hd); ops.push_back(hd);
    for (size_t i = mark; i < ops.size(); ++i) ops[i].branch = branch;
    ops[mark].wait_ev = wait_ev;
    return 0;
  };
  if (sched != 1 && push_actor_chain(side2, -1)) return -1;   // same side stream: at most two chains compete
  op.branch = 0;
  ops.insert(ops.end(), ta_ops.begin(), ta_ops.end());
  op_head_fwd(gA, PAT, h->actAT[topA], h->B, h->a16_t, &h->Xct, h->S, &op);
  const bool actor_late = sched == 1 && tuning().actor_late;
  constexpr int kLateBranch = 8;                  // a side stream nothing else uses
  if (sched == 1 && !actor_late && !side_delay) op.rec_ev = kEvActorStart;
  ops.push_back(op);
  op.rec_ev = -1;
  if (sched == 1 && !actor_late && !side_delay && push_actor_chain(2, kEvActorStart)) return -1;
  if (side_delay) {
    // target critic: first layer, then the actor's chain starts on a stream of its own; it is joined with the critic's
    // gradient branches (its consumer is the critic forward on (s, a_pi), after the critic's optimiser step)
    std::vector<Op> tc_ops;
    if (build_forward(h, gC, PCT, h->Xct, h->actCT, tc_ops, true, tuning().cluster_b != 0)) return -1;
    tc_ops[0].rec_ev = kEvActorStart;
    ops.push_back(tc_ops[0]);
    if (push_actor_chain(kLateBranch, kEvActorStart)) return -1;
    ops.insert(ops.end(), tc_ops.begin() + 1, tc_ops.end());
  } else
  if (build_forward(h, gC, PCT, h->Xct, h->actCT, ops, true, tuning().cluster_b != 0 || actor_late)) return -1;
  if (tuning().fuse_tl) {
    // dqn.cpp:892-900 TD target and the head of critic_solver_->Step(1) (loss + head backward) in one launch
    op.kind = Op::JOIN; op.mask = fork_mask; ops.push_back(op);
    push_critic_head(h, QMODE_TARGET_LOSS, PC, h->actC[topC], h->q, ops);
    CriticHeadArgs &a = ops.back().ch;
    a.H2 = h->actCT[topC].p; a.h2_plane = h->actCT[topC].plane();
    a.W2 = PCT + gC.hw_off; a.w2_plane = gC.flat; a.bias2 = PCT + gC.hb_off; a.b2_plane = gC.flat; a.q_tap2 = h->q_next;
  } else {
    push_critic_head(h, QMODE_TARGET, PCT, h->actCT[topC], h->q_next, ops);   // dqn.cpp:892-900
    op.kind = Op::JOIN; op.mask = fork_mask; ops.push_back(op);
    // rest of critic_solver_->Step(1): loss, backward, clip, Adam (+ soft update of the target critic)
    push_critic_head(h, QMODE_LOSS, PC, h->actC[topC], h->q, ops);
  }
  if (actor_late) {
    ops.back().rec_ev = kEvActorStart;            // the critic's loss head is done: the backward pass starts
    if (push_actor_chain(kLateBranch, kEvActorStart)) return -1;
  }
  Op hbw = make_head_bwd_w(h, gC, h->d16c, h->actC[topC]);
  if (build_backward(h, gC, PC, h->Xc, h->actC, true, &h->segs[1], &hbw, ops, (actor_late || side_delay) ? 1 << (kLateBranch - 1) : 0)) return -1;
  build_solver(h, 1, h->segs[1], 0.5f * h->hp.inv_batch_global, ops);
  // dqn.cpp:913-916 critic forward on (s, a_pi) with the updated critic
  if (build_forward(h, gC, PC, h->Xcp, h->actC, ops)) return -1;
  push_critic_head(h, QMODE_POLICY, PC, h->actC[topC], h->q_pi, ops);       // dqn.cpp:918-921
  // dqn.cpp:923 critic.BackwardFrom(q_values_layer): only the input diff is consumed
  if (build_backward(h, gC, PC, h->Xcp, h->actC, false, nullptr, nullptr, ops)) return -1;
  int din_splits = 1;
  if (op_dx_plain(h->cfg, gC, PC, h->dZ[0], h->d_in, (long long)h->Bp * h->Kc, &din_splits, &op)) return -1;
  ops.push_back(op);
  op.kind = Op::ACTOR_HEAD_BWD;                                  // dqn.cpp:927-961 inverting gradients + ShareDiff
  {
    ActorHeadBwdArgs &a = op.ahb;
    memset(&a, 0, sizeof(a));
    a.B = h->B; a.rows_pad = h->Bp; a.S = h->S; a.ldin = h->Kc; a.d_in = h->d_in; a.din_splits = din_splits;
    a.din_stride = (long long)h->Bp * h->Kc; a.a16 = h->a16_pi; a.d16 = h->d16a;
    a.tap_raw = h->tap_raw; a.tap_inv = h->tap_inv;
    a.W = PA + gA.hw_off; a.w_plane = gA.flat; a.Kp = gA.Hp;
    a.H = h->actA[topA].p; a.h_plane = h->actA[topA].plane(); a.ldh = h->actA[topA].ld;
    a.dZ = h->dZ[topA].p; a.dz_plane = h->dZ[topA].plane();
    op.blocks = (h->Bp + kHeadRowsPerBlock - 1) / kHeadRowsPerBlock;
  }
  ops.push_back(op);
  // dqn.cpp:963-965 actor backward + ApplyUpdate
  hbw = make_head_bwd_w(h, gA, h->d16a, h->actA[topA]);
  if (build_backward(h, gA, PA, h->Xs, h->actA, true, &h->segs[0], &hbw, ops)) return -1;
  build_solver(h, 0, h->segs[0], h->hp.inv_batch_global, ops);
  {   // the last optimiser launch also publishes (critic_loss, avg_q) and advances the counters
    AdamArgs &d = ops.back().adam;
    d.finalize = 1; d.ticket = h->ticket; d.g_critic_tail = h->Gr[1] + gC.flat; d.g_actor_tail = h->Gr[0] + gA.flat;
    d.results = h->results; d.max_slots = h->max_slots; d.done = h->d_done;
    d.snap = h->snapA; d.snap_stride = gA.flat; d.act_cur = h->act_cur;
  }
  for (Op &o : ops)     // the critic's reduction (first optimiser kernel) refreshes the per-update Adam scalars
    if (o.kind == Op::REDUCE && o.red.do_reduce) { o.red.do_prep = 1; o.red.st = h->st; o.red.hp = h->hp; break; }
  return 0;
}
static int build_update_ops(dqnb_handle_s *h) {
  for (int set = 1; set >= 0; --set)
    if (build_update_ops_for(h, set)) return -1;
  return 0;
}

// DQNB_TRACE=1: every op of the update sequence gets a kTraceSlots-slot timeline record (kernels.cuh trace_begin/end)
static void attach_trace(dqnb_handle_s *h) {
  if (!h->trace) return;
  for (int set = 0; set < 2; ++set)
  for (int i = 0; i < (int)h->update_ops_set[set].size() && i < h->trace_ops; ++i) {
    Op &op = h->update_ops_set[set][i];
    long long *t = h->trace + kTraceSlots * i;
    switch (op.kind) {
      case Op::GEMM: op.gemm.p.dbg_clk = t; break;
      case Op::GATHER: op.gather.trace = t; break;
      case Op::HEAD_FWD: op.head.trace = t; break;
      case Op::CRITIC_HEAD: op.ch.trace = t; break;
      case Op::ACTOR_HEAD_BWD: op.ahb.trace = t; break;
      case Op::HEAD_BWD_W: op.hbw.trace = t; break;
      case Op::COLSUM: op.cs.trace = t; break;
      case Op::REDUCE: op.red.trace = t; break;
      case Op::ADAM: op.adam.trace = t; break;
      case Op::P2P_ALLREDUCE: op.p2p.trace = t; break;
      default: break;
    }
  }
}

// SelectActionGreedily for up to kActMaxRows states: stage + one launch per layer + head, on the act stream
static void build_skinny_act_ops(dqnb_handle_s *h, std::vector<Op> &ops) {
  const NetGeom &g = h->gA;
  ops.clear();
  Op op;
  op.kind = Op::ACT_STAGE;
  memset(&op.ast, 0, sizeof(op.ast));
  op.ast.cur = h->act_cur; op.ast.sel = h->act_sel;
  ops.push_back(op);
  const float *x = h->d_ax + kActHdr;
  int ldx = h->Sp;
  for (int l = 0; l <= g.n_hidden; ++l) {
    const bool head = l == g.n_hidden;
    Op o;
    o.kind = Op::ACT_LAYER;
    ActLayerArgs &a = o.al;
    memset(&a, 0, sizeof(a));
    a.ctl = h->d_ctl; a.hdr = reinterpret_cast<const int *>(h->d_ax); a.snap = h->snapA; a.snap_stride = g.flat; a.sel = h->act_sel;
    a.w_off = head ? g.hw_off : g.L[l].w_off; a.b_off = head ? g.hb_off : g.L[l].b_off;
    a.Kp = head ? g.Hp : g.L[l].Kp; a.N = head ? g.head_real : g.L[l].Np;
    a.X = x; a.ldx = ldx; a.Y = h->actY[l & 1]; a.ldy = head ? 16 : g.L[l].Np;
    a.lrelu = head ? 0 : 1; a.last = head ? 1 : 0; a.h_out = h->d_ay;
    o.blocks = head ? 1 : (a.N + 7) / 8;
    ops.push_back(o);
    x = a.Y; ldx = a.ldy;
  }
}

static int build_act_ops(dqnb_handle_s *h) {
  Op op;
  h->act_ops.clear();
  if (build_forward(h, h->gA, h->P[DQNB_ACTOR], h->Xact, h->actE, h->act_ops)) return -1;
  op_head_fwd(h->gA, h->P[DQNB_ACTOR], h->actE[h->gA.n_hidden - 1], h->An, h->out16_act, nullptr, 0, &op);
  h->act_ops.push_back(op);
  h->eval_ops.clear();
  if (build_forward(h, h->gC, h->P[DQNB_CRITIC], h->Xeval, h->actE, h->eval_ops)) return -1;
  op_head_fwd(h->gC, h->P[DQNB_CRITIC], h->actE[h->gC.n_hidden - 1], h->An, h->out16_act, nullptr, 0, &op);
  h->eval_ops.push_back(op);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *dqnb_last_error(void) { return last_error().c_str(); }
const char *dqnb_version(void) { return "dqn_b200 0.1 (sm_100a; tcgen05 3xTF32 + TMA + TMEM)"; }

void dqnb_default_config(dqnb_config *c) {
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(*c);
  c->device = 0;
  c->state_size = 58;
  c->batch = 32;
  c->n_hidden = 4;
  const int hid[4] = {1024, 512, 256, 128};
  for (int i = 0; i < 4; ++i) c->hidden[i] = hid[i];
  c->replay_capacity = 500000;
  c->max_act_batch = 32;
  c->gamma = 0.99; c->beta = 0.5; c->tau = 0.001f; c->soft_update_freq = 1;
  c->actor_lr = 1e-5f; c->critic_lr = 1e-3f; c->momentum = 0.95f; c->momentum2 = 0.999f;
  c->delta = 1e-8f; c->clip_gradients = 10.f;
  c->seed = 1; c->gemm_mode = DQNB_GEMM_TCGEN05_3XTF32; c->use_graph = 1; c->world_size = 1; c->rank = 0;
}

int dqnb_destroy(dqnb_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->gstream) cudaStreamSynchronize(h->gstream);
  if (h->astream) cudaStreamSynchronize(h->astream);
  if (h->act_graph) cudaGraphExecDestroy(h->act_graph);
  for (int i = 0; i < 2; ++i) if (h->graph[i]) cudaGraphExecDestroy(h->graph[i]);
  if (h->comm && nccl().CommDestroy) nccl().CommDestroy(h->comm);
  for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  for (void *p : h->allocs) cudaFree(p);
  for (void *p : h->pinned) cudaFreeHost(p);
  for (int b = 0; b < kMaxSide; ++b) {
    if (h->ev_join[b]) cudaEventDestroy(h->ev_join[b]);
    if (h->side[b]) cudaStreamDestroy(h->side[b]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int i = 0; i < 8; ++i) if (h->evs[i]) cudaEventDestroy(h->evs[i]);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (int i = 0; i < dqnb_handle_s::kStageSlots; ++i) if (h->ev_stage[i]) cudaEventDestroy(h->ev_stage[i]);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->ev_gather) cudaEventDestroy(h->ev_gather);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_set_ready[i]) cudaEventDestroy(h->ev_set_ready[i]);
    if (h->ev_set_free[i]) cudaEventDestroy(h->ev_set_free[i]);
  }
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->astream) cudaStreamDestroy(h->astream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

static int create_impl(const dqnb_config *cfg, dqnb_handle_s *h) {
  h->cfg = *cfg;
  const dqnb_config &c = h->cfg;
  if (c.struct_size != (int32_t)sizeof(dqnb_config)) DQNB_FAIL("dqnb_config.struct_size mismatch (%d vs %d)", c.struct_size, (int)sizeof(dqnb_config));
  if (c.state_size <= 0 || c.batch <= 0 || c.n_hidden < 1 || c.n_hidden > DQNB_MAX_HIDDEN) DQNB_FAIL("bad dimensions");
  for (int l = 0; l < c.n_hidden; ++l) if (c.hidden[l] <= 0) DQNB_FAIL("bad hidden size");
  if (round_up(c.hidden[c.n_hidden - 1], 64) > kHeadMaxK) DQNB_FAIL("top tower layer wider than %d is not supported by the head kernels", kHeadMaxK);
  if (c.replay_capacity < 2) DQNB_FAIL("replay_capacity must be >= 2");
  if (c.world_size < 1 || c.rank < 0 || c.rank >= c.world_size) DQNB_FAIL("bad world_size/rank");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    DQNB_FAIL("no CUDA device: libdqn_b200 has no CPU fallback");
  if (c.device < 0 || c.device >= ndev) DQNB_FAIL("device %d out of range (%d devices)", c.device, ndev);
  cudaDeviceProp prop;
  DQNB_CUDA(cudaGetDeviceProperties(&prop, c.device));
  if (prop.major != 10) DQNB_FAIL("device %d is sm_%d%d; this library is built for sm_100a only", c.device, prop.major, prop.minor);
  DQNB_CUDA(cudaSetDevice(c.device));
  // the main stream carries the dependent chain of the update: give it priority over the side branches so
  // that its CTAs are placed first when independent forward chains compete for SMs
  int prio_lo = 0, prio_hi = 0;
  DQNB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  DQNB_CUDA(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi));
  DQNB_CUDA(cudaEventCreate(&h->ev0));
  DQNB_CUDA(cudaEventCreate(&h->ev1));
  DQNB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  if (tuning().gather_ahead) DQNB_CUDA(cudaStreamCreateWithPriority(&h->gstream, cudaStreamNonBlocking, prio_lo));
  for (int i = 0; i < 2; ++i) {
    DQNB_CUDA(cudaEventCreateWithFlags(&h->ev_set_ready[i], cudaEventDisableTiming));
    DQNB_CUDA(cudaEventCreateWithFlags(&h->ev_set_free[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < dqnb_handle_s::kStageSlots; ++i) DQNB_CUDA(cudaEventCreateWithFlags(&h->ev_stage[i], cudaEventDisableTiming));
  DQNB_CUDA(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
  DQNB_CUDA(cudaEventCreateWithFlags(&h->ev_gather, cudaEventDisableTiming));
  DQNB_CUDA(cudaEventRecord(h->ev_gather, h->stream));
  for (int b = 0; b < kMaxSide; ++b) {
    DQNB_CUDA(cudaStreamCreateWithPriority(&h->side[b], cudaStreamNonBlocking, prio_lo));
    DQNB_CUDA(cudaEventCreateWithFlags(&h->ev_join[b], cudaEventDisableTiming));
  }
  DQNB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 8; ++i) DQNB_CUDA(cudaEventCreateWithFlags(&h->evs[i], cudaEventDisableTiming));
  DQNB_CUDA(tc_prepare_all());

  h->S = c.state_size; h->Sp = round_up(c.state_size, 64); h->Kc = round_up(c.state_size + kActorOut, 64);
  h->B = c.batch; h->Bp = round_up(c.batch, 128);
  h->An = round_up(std::max(1, c.max_act_batch), 128);
  make_geom(c, false, &h->gA);
  make_geom(c, true, &h->gC);
  h->hp.gamma = c.gamma; h->hp.beta = c.beta; h->hp.tau = c.tau; h->hp.soft_update_freq = c.soft_update_freq;
  h->hp.actor_lr = c.actor_lr; h->hp.critic_lr = c.critic_lr; h->hp.beta1 = c.momentum; h->hp.beta2 = c.momentum2;
  h->hp.eps = c.delta; h->hp.clip = c.clip_gradients;
  h->hp.inv_batch_global = 1.f / (float)(c.batch * c.world_size);

  const long long fA = h->gA.flat, fC = h->gC.flat, fmax = std::max(fA, fC);
  for (int n = 0; n < 4; ++n) if (dalloc(h, &h->P[n], (size_t)2 * ((n & 1) ? fC : fA))) return -1;
  if (dalloc(h, &h->Mo[0], fA) || dalloc(h, &h->Vo[0], fA) || dalloc(h, &h->Mo[1], fC) || dalloc(h, &h->Vo[1], fC)) return -1;
  if (c.world_size > 1) {
    // gradients live in one plain cudaMalloc allocation so that it can be exported with CUDA IPC:
    // [in actor][in critic][out actor][out critic][flags]
    h->x_in[0] = 0; h->x_in[1] = fA + 4; h->x_out[0] = h->x_in[1] + fC + 4; h->x_out[1] = h->x_out[0] + fA + 4;
    h->x_flag = h->x_out[1] + fC + 4;
    h->xchg_floats = h->x_flag + 256;
    if (dalloc(h, &h->xchg, (size_t)h->xchg_floats)) return -1;
    for (int n = 0; n < 2; ++n) { h->G[n] = h->xchg + h->x_in[n]; h->Gr[n] = h->G[n]; }
    if (dalloc(h, &h->p2p_tab, 1) || dalloc(h, &h->p2p_epoch, 2) || dalloc(h, &h->p2p_ticket, 2) ||
        dalloc(h, &h->p2p_block_ss, 2 * 256) || dalloc(h, &h->p2p_err_dev, 1) || dalloc(h, &h->flag_ticket, 2)) return -1;
    {
      void *hp = nullptr, *dp = nullptr;
      DQNB_CUDA(cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
      memset(hp, 0, 64);
      h->pinned.push_back(hp);
      DQNB_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
      h->h_p2p_err = (volatile int *)hp; h->p2p_err = (int *)dp;
    }
  } else {
    if (dalloc(h, &h->G[0], fA + 4) || dalloc(h, &h->G[1], fC + 4)) return -1;
    h->Gr[0] = h->G[0]; h->Gr[1] = h->G[1];
  }
  h->gpart_stride[0] = fA; h->gpart_stride[1] = fC;
  if (dalloc(h, &h->Gpart[0], (size_t)kGradSplits * fA) || dalloc(h, &h->Gpart[1], (size_t)kGradSplits * fC)) return -1;
  h->bflat = 0;
  for (int l = 0; l < c.n_hidden; ++l) { h->boff[l] = h->bflat; h->bflat += h->gA.L[l].Np; }
  h->bplanes = std::max((int)kGradSplits, h->Bp / BM);
  if (dalloc(h, &h->Bpart[0], (size_t)h->bplanes * h->bflat) || dalloc(h, &h->Bpart[1], (size_t)h->bplanes * h->bflat)) return -1;
  if (dalloc(h, &h->norm_part, (size_t)(fmax / 1024))) return -1;
  h->n_scal = (h->Bp + kHeadRowsPerBlock - 1) / kHeadRowsPerBlock;     // one loss / avg-q partial per critic_head block
  if (dalloc(h, &h->scal_part, (size_t)h->n_scal)) return -1;
  // replay ring (rows padded to Sp floats = 256 B multiples: aligned, vectorisable gathers)
  const size_t cap = (size_t)c.replay_capacity;
  h->rw = 2 * h->Sp + kMiscStride;
  if (dalloc(h, &h->ring, cap * h->rw)) return -1;
  for (int set = 0; set < 2; ++set) {
    dqnb_handle_s::InputSet &I = h->in[set];
    if (dalloc(h, &I.idx, (size_t)h->Bp)) return -1;
    if (alloc_mat(h, &I.Xs, h->Bp, h->Sp) || alloc_mat(h, &I.Xsn, h->Bp, h->Sp) || alloc_mat(h, &I.Xc, h->Bp, h->Kc) ||
        alloc_mat(h, &I.Xct, h->Bp, h->Kc) || alloc_mat(h, &I.Xcp, h->Bp, h->Kc)) return -1;
    if (dalloc(h, &I.reward, (size_t)h->Bp) || dalloc(h, &I.mc, (size_t)h->Bp) || dalloc(h, &I.term, (size_t)h->Bp)) return -1;
  }
  select_input_set(h, 0);
  float **vecs[] = {&h->y, &h->q_next, &h->q, &h->q_pi};
  for (float **v : vecs) if (dalloc(h, v, (size_t)h->Bp)) return -1;
  const int rows16 = std::max(h->Bp, h->An);
  float **m16[] = {&h->a16_t, &h->a16_pi, &h->d16c, &h->d16a, &h->out16_act};
  for (float **v : m16) if (dalloc(h, v, (size_t)rows16 * 16)) return -1;
  if (dalloc(h, &h->d_in, (size_t)kGradSplits * h->Bp * h->Kc) || dalloc(h, &h->tap_raw, (size_t)h->Bp * kActorOut) || dalloc(h, &h->tap_inv, (size_t)h->Bp * kActorOut)) return -1;
  for (int l = 0; l < c.n_hidden; ++l) {
    const int Np = h->gA.L[l].Np;
    if (alloc_mat(h, &h->actAT[l], h->Bp, Np) || alloc_mat(h, &h->actCT[l], h->Bp, Np) || alloc_mat(h, &h->actC[l], h->Bp, Np, true) ||
        alloc_mat(h, &h->actA[l], h->Bp, Np, true) || alloc_mat(h, &h->dZ[l], h->Bp, Np) || alloc_mat(h, &h->actE[l], h->An, Np)) return -1;
  }
  if (alloc_mat(h, &h->Xact, h->An, h->Sp) || alloc_mat(h, &h->Xeval, h->An, h->Kc)) return -1;
  if (halloc(h, &h->h_act_in, (size_t)2 * h->An * h->Kc) || halloc(h, &h->h_act_out, (size_t)h->An * 16)) return -1;
  {   // skinny act path: snapshots, staging, host-mapped control / rows
    int maxNp = h->Sp;
    for (int l = 0; l < c.n_hidden; ++l) maxNp = std::max(maxNp, h->gA.L[l].Np);
    if (dalloc(h, &h->snapA, (size_t)2 * fA) || dalloc(h, &h->act_cur, 1) || dalloc(h, &h->act_sel, 1) ||
        dalloc(h, &h->d_ax, (size_t)kActHdr + (size_t)kActMaxRows * h->Sp) || dalloc(h, &h->actY[0], (size_t)kActMaxRows * maxNp) ||
        dalloc(h, &h->actY[1], (size_t)kActMaxRows * maxNp)) return -1;
    auto mapped = [&](size_t bytes, void **hp, void **dp) -> int {
      DQNB_CUDA(cudaHostAlloc(hp, bytes, cudaHostAllocMapped));
      memset(*hp, 0, bytes);
      h->pinned.push_back(*hp);
      DQNB_CUDA(cudaHostGetDevicePointer(dp, *hp, 0));
      return 0;
    };
    if (mapped(sizeof(ActCtl), (void **)&h->h_ctl, (void **)&h->d_ctl) ||
        mapped(sizeof(float) * kActMaxRows * 16, (void **)&h->h_ay, (void **)&h->d_ay)) return -1;
    if (halloc(h, &h->h_ax, (size_t)kActHdr + (size_t)kActMaxRows * h->Sp)) return -1;
    if (h->gA.head_real > 16) DQNB_FAIL("act path: head wider than 16");
    DQNB_CUDA(cudaStreamCreateWithPriority(&h->astream, cudaStreamNonBlocking, prio_hi));
  }
  if (dalloc(h, &h->st, 1) || dalloc(h, &h->ticket, 1)) return -1;
  {   // results live in mapped pinned host memory: reading them back needs a stream sync, no copy
    void *hp = nullptr, *dp = nullptr;
    DQNB_CUDA(cudaHostAlloc(&hp, sizeof(float) * 2 * h->max_slots, cudaHostAllocMapped));
    memset(hp, 0, sizeof(float) * 2 * h->max_slots);
    h->pinned.push_back(hp);
    DQNB_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
    h->h_results = (float *)hp; h->results = (float *)dp;
  }
  h->stage_rows = 4096;
  for (int i = 0; i < dqnb_handle_s::kStageSlots; ++i)
    if (halloc(h, &h->h_stage[i], (size_t)h->stage_rows * h->rw + 4)) return -1;
  {   // update counter the last optimiser launch publishes for the host (dqnb_results polls it)
    void *hp = nullptr, *dp = nullptr;
    DQNB_CUDA(cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
    memset(hp, 0, 64);
    h->pinned.push_back(hp);
    DQNB_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
    h->h_done = (volatile unsigned long long *)hp; h->d_done = (unsigned long long *)dp;
  }
  if (build_update_ops(h) || build_act_ops(h)) return -1;
  if (getenv("DQNB_TRACE")) {
    h->trace_ops = (int)h->update_ops.size() + 16;   // head room: the op list grows when a communicator is attached
    if (dalloc(h, &h->trace, (size_t)kTraceSlots * h->trace_ops)) return -1;
    attach_trace(h);
  }
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  DQNB_CUDA(cudaDeviceSynchronize());
  return 0;
}

int dqnb_create(const dqnb_config *cfg, dqnb_handle *out) {
  if (!cfg || !out) DQNB_FAIL("null argument");
  dqnb_handle_s *h = new dqnb_handle_s();
  if (create_impl(cfg, h)) {
    std::string keep = last_error();
    dqnb_destroy(h);
    last_error() = keep;
    *out = nullptr;
    return -1;
  }
  *out = h;
  return 0;
}

int64_t dqnb_param_count(dqnb_handle h, int net) {
  if (!h || net < 0 || net > 3) return -1;
  return (net & 1) ? h->gC.caffe_count : h->gA.caffe_count;
}

int dqnb_set_params(dqnb_handle h, int net, const float *params) {
  if (!h || net < 0 || net > 3 || !params) DQNB_FAIL("bad argument");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  const NetGeom &g = (net & 1) ? h->gC : h->gA;
  std::vector<float> in;
  caffe_to_internal(g, params, in);
  float *tmp = nullptr;
  DQNB_CUDA(cudaMalloc(&tmp, sizeof(float) * g.flat));
  DQNB_CUDA(cudaMemcpyAsync(tmp, in.data(), sizeof(float) * g.flat, cudaMemcpyHostToDevice, h->stream));
  split_kernel<<<(unsigned)((g.flat + 255) / 256), 256, 0, h->stream>>>(tmp, h->P[net], h->P[net] + g.flat, g.flat);
  DQNB_CUDA(cudaGetLastError());
  if (net == DQNB_ACTOR) {
    // the act path reads fp32 snapshots of the actor: both buffers take the new weights (no act call is in flight
    // on the caller's side while it replaces the weights: same single-threaded contract as the reference)
    DQNB_CUDA(cudaStreamSynchronize(h->astream));
    DQNB_CUDA(cudaMemcpyAsync(h->snapA, tmp, sizeof(float) * g.flat, cudaMemcpyDeviceToDevice, h->stream));
    DQNB_CUDA(cudaMemcpyAsync(h->snapA + g.flat, tmp, sizeof(float) * g.flat, cudaMemcpyDeviceToDevice, h->stream));
  }
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  DQNB_CUDA(cudaFree(tmp));
  return 0;
}

// Multi-agent parameter sharing (dqn.cpp:1048-1079 ShareParameters -> ShareLayer -> Blob::ShareData): the first
// n layers-with-parameters (Caffe layer order: tower layers, then the head layers) of src's actor / critic and of their
// target nets become dst's.  Two handles cannot alias sub-ranges of each other's flat buffers, so the host mirror calls
// this after every update of a sharing group (write-through): what a member reads at its next update or act call is
// what Blob::ShareData would have shown it.
static void shared_ranges(const NetGeom &g, int n_layers, std::vector<std::pair<long long, long long>> &out) {
  int i = 0;
  for (; i < g.n_hidden && i < n_layers; ++i) {
    out.push_back({g.L[i].w_off, (long long)g.L[i].Np * g.L[i].Kp});
    out.push_back({g.L[i].b_off, (long long)g.L[i].Np});
  }
  if (i >= n_layers) return;
  if (g.critic) {                                   // q_values_layer
    out.push_back({g.hw_off, (long long)g.Hp}); out.push_back({g.hb_off, 1});
    return;
  }
  out.push_back({g.hw_off, 4LL * g.Hp}); out.push_back({g.hb_off, 4});                       // action_layer
  if (i + 1 < n_layers) { out.push_back({g.hw_off + 4LL * g.Hp, 6LL * g.Hp}); out.push_back({g.hb_off + 4, 6}); }   // actionpara_layer
}
static int sync_all(dqnb_handle h);
int dqnb_copy_shared_layers(dqnb_handle dst, dqnb_handle src, int32_t n_actor_layers, int32_t n_critic_layers) {
  if (!dst || !src || dst == src || n_actor_layers < 0 || n_critic_layers < 0) DQNB_FAIL("bad argument");
  if (dst->gA.flat != src->gA.flat || dst->gC.flat != src->gC.flat || dst->cfg.n_hidden != src->cfg.n_hidden)
    DQNB_FAIL("sharing needs identical net shapes");
  if (n_actor_layers > src->gA.n_hidden + 2 || n_critic_layers > src->gC.n_hidden + 1) DQNB_FAIL("more layers to share than the net has");
  if (sync_all(src) || sync_all(dst)) return -1;
  DQNB_CUDA(cudaSetDevice(dst->cfg.device));
  DQNB_CUDA(cudaStreamSynchronize(dst->astream));
  for (int critic = 0; critic < 2; ++critic) {
    const NetGeom &g = critic ? dst->gC : dst->gA;
    std::vector<std::pair<long long, long long>> rs;
    shared_ranges(g, critic ? n_critic_layers : n_actor_layers, rs);
    for (int target = 0; target < 2; ++target) {
      const int net = (critic ? DQNB_CRITIC : DQNB_ACTOR) + (target ? 2 : 0);
      for (auto &r : rs)
        for (int plane = 0; plane < 2; ++plane)
          DQNB_CUDA(cudaMemcpyAsync(dst->P[net] + plane * g.flat + r.first, src->P[net] + plane * g.flat + r.first,
                                    sizeof(float) * r.second, cudaMemcpyDefault, dst->stream));
    }
    if (!critic && dst->snapA)                      // dst's act path reads fp32 snapshots of its actor: refresh both
      for (auto &r : rs)
        for (int buf = 0; buf < 2; ++buf) {
          join_kernel<<<(unsigned)((r.second + 255) / 256), 256, 0, dst->stream>>>(
              dst->P[DQNB_ACTOR] + r.first, dst->P[DQNB_ACTOR] + g.flat + r.first, dst->snapA + buf * g.flat + r.first, r.second);
          DQNB_CUDA(cudaGetLastError());
        }
  }
  DQNB_CUDA(cudaStreamSynchronize(dst->stream));
  return 0;
}

static int read_flat(dqnb_handle h, const NetGeom &g, const float *dev_hi, const float *dev_lo, float *caffe_out) {
  std::vector<float> a((size_t)g.flat), b;
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  DQNB_CUDA(cudaMemcpy(a.data(), dev_hi, sizeof(float) * g.flat, cudaMemcpyDeviceToHost));
  if (dev_lo) {
    b.resize((size_t)g.flat);
    DQNB_CUDA(cudaMemcpy(b.data(), dev_lo, sizeof(float) * g.flat, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < g.flat; ++i) a[i] += b[i];
  }
  internal_to_caffe(g, a.data(), caffe_out);
  return 0;
}

int dqnb_get_params(dqnb_handle h, int net, float *params) {
  if (!h || net < 0 || net > 3 || !params) DQNB_FAIL("bad argument");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  const NetGeom &g = (net & 1) ? h->gC : h->gA;
  return read_flat(h, g, h->P[net], h->P[net] + g.flat, params);
}

int dqnb_clone_targets(dqnb_handle h) {
  if (!h) DQNB_FAIL("null handle");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  DQNB_CUDA(cudaMemcpyAsync(h->P[DQNB_ACTOR_TARGET], h->P[DQNB_ACTOR], sizeof(float) * 2 * h->gA.flat, cudaMemcpyDeviceToDevice, h->stream));
  DQNB_CUDA(cudaMemcpyAsync(h->P[DQNB_CRITIC_TARGET], h->P[DQNB_CRITIC], sizeof(float) * 2 * h->gC.flat, cudaMemcpyDeviceToDevice, h->stream));
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int dqnb_init_params(dqnb_handle h, uint64_t seed, float stddev) {
  if (!h) DQNB_FAIL("null handle");
  std::mt19937 rng((uint32_t)seed);
  std::normal_distribution<float> nd(0.f, stddev);
  for (int net = 0; net < 2; ++net) {
    const NetGeom &g = net ? h->gC : h->gA;
    std::vector<float> p((size_t)g.caffe_count, 0.f);
    for (int l = 0; l < g.n_hidden; ++l) {
      const LayerGeom &L = g.L[l];
      for (long long i = 0; i < (long long)L.n_real * L.k_real; ++i) p[L.cw_off + i] = nd(rng);
    }
    const int k_real = g.L[g.n_hidden - 1].n_real;
    const int nh = g.critic ? 1 : 2;
    const int rows[2] = {g.critic ? 1 : 4, 6};
    for (int hd = 0; hd < nh; ++hd)
      for (long long i = 0; i < (long long)rows[hd] * k_real; ++i) p[g.chw_off[hd] + i] = nd(rng);
    if (dqnb_set_params(h, net, p.data())) return -1;
  }
  return dqnb_clone_targets(h);
}

static int upload_flat(dqnb_handle h, const NetGeom &g, const float *caffe, float *dev) {
  std::vector<float> in;
  caffe_to_internal(g, caffe, in);
  DQNB_CUDA(cudaMemcpy(dev, in.data(), sizeof(float) * g.flat, cudaMemcpyHostToDevice));
  return 0;
}

static int sync_all(dqnb_handle h);
static int pull_state(dqnb_handle h, StepState *s) {
  if (sync_all(h)) return -1;
  DQNB_CUDA(cudaMemcpy(s, h->st, sizeof(StepState), cudaMemcpyDeviceToHost));
  return 0;
}
static int push_state(dqnb_handle h, const StepState *s) {
  DQNB_CUDA(cudaMemcpy(h->st, s, sizeof(StepState), cudaMemcpyHostToDevice));
  return 0;
}

int dqnb_set_opt_state(dqnb_handle h, int net, const float *m, const float *v, int32_t iter) {
  if (!h || net < 0 || net > 1) DQNB_FAIL("bad argument");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  const NetGeom &g = net ? h->gC : h->gA;
  if (m && upload_flat(h, g, m, h->Mo[net])) return -1;
  if (v && upload_flat(h, g, v, h->Vo[net])) return -1;
  StepState s;
  if (pull_state(h, &s)) return -1;
  if (net) s.critic_iter = iter; else s.actor_iter = iter;
  return push_state(h, &s);
}

int dqnb_get_opt_state(dqnb_handle h, int net, float *m, float *v, int32_t *iter) {
  if (!h || net < 0 || net > 1) DQNB_FAIL("bad argument");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  const NetGeom &g = net ? h->gC : h->gA;
  if (m && read_flat(h, g, h->Mo[net], nullptr, m)) return -1;
  if (v && read_flat(h, g, h->Vo[net], nullptr, v)) return -1;
  if (iter) {
    StepState s;
    if (pull_state(h, &s)) return -1;
    *iter = net ? s.critic_iter : s.actor_iter;
  }
  return 0;
}

int dqnb_iters(dqnb_handle h, int32_t *actor_iter, int32_t *critic_iter) {
  if (!h) DQNB_FAIL("null handle");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  StepState s;
  if (pull_state(h, &s)) return -1;
  if (actor_iter) *actor_iter = s.actor_iter;
  if (critic_iter) *critic_iter = s.critic_iter;
  return 0;
}

// ----------------------------------- replay ring -----------------------------------------------
// Everything queued on the copy stream happens before whatever the compute stream gets next.
static int order_compute_after_copies(dqnb_handle h) {
  if (h->copy_pending) {
    DQNB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
    h->copy_pending = false;
  }
  return 0;
}
static int sync_all(dqnb_handle h) {
  DQNB_CUDA(cudaStreamSynchronize(h->copy_stream));
  if (h->gstream) DQNB_CUDA(cudaStreamSynchronize(h->gstream));
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

static int push_ring_state(dqnb_handle h) {
  // head/size live in StepState so the captured graph always sees the current ring (rare path: blocking)
  if (sync_all(h)) return -1;
  int hs[2] = {h->ring_head, h->ring_size};
  DQNB_CUDA(cudaMemcpy(&h->st->ring_head, hs, sizeof(hs), cudaMemcpyHostToDevice));
  return 0;
}

static int append_rows(dqnb_handle h, int32_t n, const float *s, const float *act10, const float *reward,
                       const float *mc, const float *s_next, const uint8_t *terminal) {
  const int cap = h->cfg.replay_capacity, S = h->S, Sp = h->Sp, rw = h->rw;
  int done = 0;
  while (done < n) {
    const int chunk = std::min(n - done, h->stage_rows);
    const int slot = h->stage_next;
    h->stage_next = (slot + 1) % dqnb_handle_s::kStageSlots;
    DQNB_CUDA(cudaEventSynchronize(h->ev_stage[slot]));   // the copies out of this slot are done
    float *stage = h->h_stage[slot];
    for (int i = 0; i < chunk; ++i) {              // padding columns of the staging rows are zero for good
      const int r = done + i;
      float *row = stage + (size_t)i * rw;
      memcpy(row, s + (size_t)r * S, sizeof(float) * S);
      const bool t = terminal[r] != 0;
      if (!t && s_next) memcpy(row + Sp, s_next + (size_t)r * S, sizeof(float) * S); else memset(row + Sp, 0, sizeof(float) * S);
      float *dm = row + 2 * Sp;
      memcpy(dm, act10 + (size_t)r * kActorOut, sizeof(float) * kActorOut);
      dm[10] = reward[r]; dm[11] = mc[r]; dm[12] = t ? 1.f : 0.f;
    }
    // the only reader of the ring (and of head/size) is the gather kernel of an update: wait for the last one queued
    DQNB_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_gather, 0));
    int tail = (h->ring_head + h->ring_size) % cap;
    int left = chunk, src = 0;
    while (left > 0) {                              // one copy per contiguous run (two when the ring wraps)
      const int run = std::min(left, cap - tail);
      DQNB_CUDA(cudaMemcpyAsync(h->ring + (size_t)tail * rw, stage + (size_t)src * rw, sizeof(float) * (size_t)run * rw, cudaMemcpyHostToDevice, h->copy_stream));
      tail = (tail + run) % cap; src += run; left -= run;
    }
    h->ring_size += chunk;
    int *hs = reinterpret_cast<int *>(stage + (size_t)h->stage_rows * rw);
    hs[0] = h->ring_head; hs[1] = h->ring_size;
    DQNB_CUDA(cudaMemcpyAsync(&h->st->ring_head, hs, 2 * sizeof(int), cudaMemcpyHostToDevice, h->copy_stream));
    DQNB_CUDA(cudaEventRecord(h->ev_stage[slot], h->copy_stream));
    done += chunk;
  }
  DQNB_CUDA(cudaEventRecord(h->ev_copy, h->copy_stream));
  h->copy_pending = true;
  return 0;
}

int dqnb_add_transitions(dqnb_handle h, int32_t n, const float *s, const float *act10, const float *reward,
                         const float *mc_target, const float *s_next, const uint8_t *terminal) {
  if (!h || n < 0 || (n > 0 && (!s || !act10 || !reward || !mc_target || !terminal))) DQNB_FAIL("bad argument");
  if (n == 0) return 0;
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  const int cap = h->cfg.replay_capacity;
  if (n >= cap) DQNB_FAIL("AddTransitions: %d rows do not fit capacity %d (dqn.cpp:776 keeps size < capacity)", n, cap);
  // dqn.cpp:776-778: while (size + n >= capacity) pop_front()
  while (h->ring_size + n >= cap && h->ring_size > 0) { h->ring_head = (h->ring_head + 1) % cap; h->ring_size--; }
  return append_rows(h, n, s, act10, reward, mc_target, s_next, terminal);
}

int dqnb_add_transition(dqnb_handle h, const float *s, const float *act10, float reward, float mc_target,
                        const float *s_next, uint8_t terminal) {
  if (!h || !s || !act10) DQNB_FAIL("bad argument");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  const int cap = h->cfg.replay_capacity;
  // dqn.cpp:769-771: if (size == capacity) pop_front()
  if (h->ring_size == cap) { h->ring_head = (h->ring_head + 1) % cap; h->ring_size--; }
  return append_rows(h, 1, s, act10, &reward, &mc_target, s_next, &terminal);
}

int32_t dqnb_memory_size(dqnb_handle h) { return h ? h->ring_size : -1; }

int dqnb_clear_memory(dqnb_handle h) {
  if (!h) DQNB_FAIL("null handle");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  h->ring_head = 0; h->ring_size = 0;
  return push_ring_state(h);
}

int dqnb_get_transitions(dqnb_handle h, int32_t first, int32_t n, float *s, float *act10, float *reward,
                         float *mc_target, float *s_next, uint8_t *terminal) {
  if (!h || first < 0 || n < 0 || first + n > h->ring_size) DQNB_FAIL("bad range");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  if (sync_all(h)) return -1;
  const int cap = h->cfg.replay_capacity, S = h->S, Sp = h->Sp;
  const int rw = h->rw;
  std::vector<float> rows((size_t)n * rw);
  int phys = (h->ring_head + first) % cap, left = n, dst = 0;
  while (left > 0) {
    const int run = std::min(left, cap - phys);
    DQNB_CUDA(cudaMemcpy(rows.data() + (size_t)dst * rw, h->ring + (size_t)phys * rw, sizeof(float) * (size_t)run * rw, cudaMemcpyDeviceToHost));
    phys = (phys + run) % cap; dst += run; left -= run;
  }
  for (int i = 0; i < n; ++i) {
    const float *row = rows.data() + (size_t)i * rw, *m = row + 2 * Sp;
    if (s) memcpy(s + (size_t)i * S, row, sizeof(float) * S);
    if (s_next) memcpy(s_next + (size_t)i * S, row + Sp, sizeof(float) * S);
    if (act10) memcpy(act10 + (size_t)i * kActorOut, m, sizeof(float) * kActorOut);
    if (reward) reward[i] = m[10];
    if (mc_target) mc_target[i] = m[11];
    if (terminal) terminal[i] = m[12] != 0.f;
  }
  return 0;
}

// ----------------------------------- update ----------------------------------------------------
static int ensure_graph(dqnb_handle h, int set) {
  cudaGraphExec_t *slot = &h->graph[set];
  if (*slot) return 0;
  cudaGraph_t graph = nullptr;
  int count = 0;
  DQNB_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  int rc = run_ops(h, h->update_ops_set[set], h->stream, false, &count);
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return -1; }
  DQNB_CUDA(e);
  DQNB_CUDA(cudaGraphInstantiate(slot, graph, 0));
  DQNB_CUDA(cudaGraphDestroy(graph));
  h->kernels_per_update = count;
  return 0;
}

// One update = the gather (dqn.cpp:846-887), the one reader of the replay ring, then the captured graph.  The gather of
// update t fills input set t % 2 on the gather stream: it waits for the appends queued so far (ev_copy) and for the
// last update that read this set (t - 2), not for update t - 1, so it runs beside it; an event tells the copy stream
// "ring consumed" (append_rows).  Injected indices travel on the same stream ahead of the gather.
static int enqueue_update(dqnb_handle h, const int32_t *injected_idx) {
  const int set = (int)(h->host_step & 1ull);
  cudaStream_t gs = h->gstream ? h->gstream : h->stream;
  if (h->trace) DQNB_CUDA(cudaMemsetAsync(h->trace, 0xFF, sizeof(long long) * kTraceSlots * h->trace_ops, h->stream));
  if (h->gstream) {
    DQNB_CUDA(cudaStreamWaitEvent(gs, h->ev_copy, 0));
    DQNB_CUDA(cudaStreamWaitEvent(gs, h->ev_set_free[set], 0));
    h->copy_pending = false;
  } else if (order_compute_after_copies(h)) {
    return -1;
  }
  if (injected_idx) DQNB_CUDA(cudaMemcpyAsync(h->in[set].idx, injected_idx, sizeof(int32_t) * h->B, cudaMemcpyHostToDevice, gs));
  for (const Op &op : h->update_ops_set[set])
    if (op.kind == Op::GATHER && op.variant == (injected_idx ? 2 : 1)) {
      Op g = op;
      g.gather.step = h->host_step;          // == StepState::step when this update runs
      if (launch_op(h, g, gs)) return -1;
      h->launches += 1;
    }
  DQNB_CUDA(cudaEventRecord(h->ev_gather, gs));
  if (h->gstream) {
    DQNB_CUDA(cudaEventRecord(h->ev_set_ready[set], gs));
    DQNB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_set_ready[set], 0));
  }
  if (h->cfg.use_graph) {
    if (ensure_graph(h, set)) return -1;
    DQNB_CUDA(cudaGraphLaunch(h->graph[set], h->stream));
    h->launches += h->kernels_per_update;
  } else {
    int count = 0;
    if (run_ops(h, h->update_ops_set[set], h->stream, false, &count)) return -1;
    h->launches += count;
  }
  if (h->gstream) DQNB_CUDA(cudaEventRecord(h->ev_set_free[set], h->stream));
  h->last_set = set;
  h->host_step += 1;
  return 0;
}

// (critic_loss, avg_q) of the last n updates: the optimiser launch of update k wrote slot k % max_slots of
// the mapped pinned ring; a stream sync makes them visible
static int comm_failed(dqnb_handle h) {
  if (h->h_p2p_err && *h->h_p2p_err != 0)
    DQNB_FAIL("gradient exchange failed: a peer did not arrive within DQNB_P2P_TIMEOUT_MS; the replicas have stopped updating "
              "(parameters, Adam moments and iteration counters are those of the last good update)");
  return 0;
}
static int fetch_results(dqnb_handle h, int n, float *critic_loss, float *avg_q) {
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  if (comm_failed(h)) return -1;
  for (int i = 0; i < n; ++i) {
    const int slot = (int)((h->host_step - (unsigned long long)n + i) % (unsigned long long)h->max_slots);
    if (critic_loss) critic_loss[i] = h->h_results[2 * slot];
    if (avg_q) avg_q[i] = h->h_results[2 * slot + 1];
  }
  return 0;
}

// Asynchronous pair: enqueue updates without waiting ...
int dqnb_update_async(dqnb_handle h, int32_t n_updates, int64_t *last_step) {
  if (!h || n_updates < 0) DQNB_FAIL("bad argument");
  if (n_updates > 0 && h->ring_size <= 0) DQNB_FAIL("Update on an empty replay memory");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  for (int i = 0; i < n_updates; ++i) if (enqueue_update(h, nullptr)) return -1;
  if (last_step) *last_step = (int64_t)h->host_step;
  return 0;
}

// ... and collect (critic_loss, avg_q) of updates [first_step, first_step + n) (1-based sequence numbers as returned
// by dqnb_update_async) once the last of them has finished.  The optimiser's final block publishes the results and
// the number of finished updates in mapped pinned memory, so this is a spin on a host word, not a stream sync.
int dqnb_results(dqnb_handle h, int64_t first_step, int32_t n, float *critic_loss, float *avg_q) {
  if (!h || first_step < 1 || n < 0) DQNB_FAIL("bad argument");
  if (n == 0) return 0;
  const unsigned long long last = (unsigned long long)first_step + n - 1;
  if (last > h->host_step) DQNB_FAIL("results of update %llu requested but only %llu were enqueued", last, h->host_step);
  if (h->host_step - (unsigned long long)first_step >= (unsigned long long)h->max_slots) DQNB_FAIL("results of update %lld are no longer kept (%d slots)", (long long)first_step, h->max_slots);
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  unsigned spins = 0;
  while (*h->h_done < last) {
    if ((++spins & 0x3ff) == 0) {
      const cudaError_t q = cudaStreamQuery(h->stream);
      if (q == cudaSuccess) break;                         // stream drained: the counter is final
      if (q != cudaErrorNotReady) DQNB_CUDA(q);
    }
  }
  if (*h->h_done < last) DQNB_FAIL("update %llu did not complete (device reports %llu)", last, (unsigned long long)*h->h_done);
  if (comm_failed(h)) return -1;
  for (int i = 0; i < n; ++i) {
    const int slot = (int)(((unsigned long long)first_step - 1 + i) % (unsigned long long)h->max_slots);
    if (critic_loss) critic_loss[i] = h->h_results[2 * slot];
    if (avg_q) avg_q[i] = h->h_results[2 * slot + 1];
  }
  return 0;
}

int dqnb_update(dqnb_handle h, int32_t n_updates, float *critic_loss, float *avg_q) {
  if (!h || n_updates < 0) DQNB_FAIL("bad argument");
  if (n_updates == 0) return 0;
  if (h->ring_size <= 0) DQNB_FAIL("Update on an empty replay memory");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  int done = 0;
  while (done < n_updates) {
    const int chunk = std::min(n_updates - done, h->max_slots);
    for (int i = 0; i < chunk; ++i) if (enqueue_update(h, nullptr)) return -1;
    if (fetch_results(h, chunk, critic_loss ? critic_loss + done : nullptr, avg_q ? avg_q + done : nullptr)) return -1;
    done += chunk;
  }
  return 0;
}

int dqnb_update_with_indices(dqnb_handle h, const int32_t *idx, float *critic_loss, float *avg_q) {
  if (!h || !idx) DQNB_FAIL("bad argument");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  for (int i = 0; i < h->B; ++i)
    if (idx[i] < 0 || idx[i] >= h->ring_size) DQNB_FAIL("index %d out of range [0,%d)", idx[i], h->ring_size);
  if (enqueue_update(h, idx)) return -1;     // idx is consumed by an asynchronous copy: wait before returning
  if (h->gstream) DQNB_CUDA(cudaStreamSynchronize(h->gstream));
  return fetch_results(h, 1, critic_loss, avg_q);
}

int dqnb_benchmark(dqnb_handle h, int32_t n_updates, float *elapsed_ms) {
  if (!h || n_updates <= 0 || !elapsed_ms) DQNB_FAIL("bad argument");
  if (h->ring_size <= 0) DQNB_FAIL("Benchmark on an empty replay memory");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  if (h->cfg.use_graph && (ensure_graph(h, 0) || ensure_graph(h, 1))) return -1;
  if (sync_all(h)) return -1;
  DQNB_CUDA(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < n_updates; ++i) if (enqueue_update(h, nullptr)) return -1;
  DQNB_CUDA(cudaEventRecord(h->ev1, h->stream));
  DQNB_CUDA(cudaEventSynchronize(h->ev1));
  DQNB_CUDA(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  return 0;
}

// Device time of the dense-layer launches of one update, back to back (bench.py roofline leg).
int dqnb_benchmark_gemms(dqnb_handle h, int32_t reps, float *ms_per_update, int32_t *gemm_launches) {
  if (!h || reps <= 0 || !ms_per_update) DQNB_FAIL("bad argument");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  int n = 0;
  for (const Op &op : h->update_ops) if (op.kind == Op::GEMM) ++n;
  if (sync_all(h)) return -1;
  // one update's GEMM launches captured as a graph (eager launches of 38 kernels with 400-byte argument blocks are
  // bound by the host, not by the device: 10-15 us per launch measured), replayed reps times between two events
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  DQNB_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  int rc = 0;
  for (const Op &op : h->update_ops)
    if (op.kind == Op::GEMM && launch_op(h, op, h->stream)) { rc = -1; break; }
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return -1; }
  DQNB_CUDA(e);
  DQNB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  DQNB_CUDA(cudaGraphDestroy(graph));
  for (int r = 0; r < reps + 2; ++r) {
    if (r == 2) DQNB_CUDA(cudaEventRecord(h->ev0, h->stream));
    DQNB_CUDA(cudaGraphLaunch(exec, h->stream));
  }
  DQNB_CUDA(cudaEventRecord(h->ev1, h->stream));
  DQNB_CUDA(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  DQNB_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  DQNB_CUDA(cudaGraphExecDestroy(exec));
  *ms_per_update = ms / reps;
  if (gemm_launches) *gemm_launches = n;
  h->launches += (int64_t)n * (reps + 2);
  return 0;
}

int dqnb_peek_sample_indices(dqnb_handle h, int32_t *idx) {
  if (!h || !idx) DQNB_FAIL("bad argument");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  if (order_compute_after_copies(h)) return -1;
  if (sync_all(h)) return -1;               // neither input set is in use: borrow the idle one's index buffer
  int32_t *scratch = h->in[(h->host_step & 1ull) ? 1 : 0].idx;
  sample_kernel<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->st, h->cfg.seed, h->host_step, h->B, scratch);
  DQNB_CUDA(cudaGetLastError());
  DQNB_CUDA(cudaMemcpyAsync(idx, scratch, sizeof(int32_t) * h->B, cudaMemcpyDeviceToHost, h->stream));
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// ----------------------------------- act path --------------------------------------------------
static int stage_rows_split(dqnb_handle h, int n, const float *src, int src_ld, int cols, const float *src2, int cols2,
                            const SplitMat &X) {
  // host-side split into (hi, lo) planes of the padded input block, one H2D per plane
  const int ld = X.ld;
  float *hi = h->h_act_in, *lo = h->h_act_in + (size_t)h->An * h->Kc;
  memset(hi, 0, sizeof(float) * (size_t)n * ld);
  memset(lo, 0, sizeof(float) * (size_t)n * ld);
  for (int i = 0; i < n; ++i) {
    for (int c = 0; c < cols; ++c) {
      const float x = src[(size_t)i * src_ld + c], hh = tf32_hi(x);
      hi[(size_t)i * ld + c] = hh; lo[(size_t)i * ld + c] = x - hh;
    }
    for (int c = 0; c < cols2; ++c) {
      const float x = src2[(size_t)i * cols2 + c], hh = tf32_hi(x);
      hi[(size_t)i * ld + cols + c] = hh; lo[(size_t)i * ld + cols + c] = x - hh;
    }
  }
  DQNB_CUDA(cudaMemcpyAsync(X.p, hi, sizeof(float) * (size_t)n * ld, cudaMemcpyHostToDevice, h->stream));
  DQNB_CUDA(cudaMemcpyAsync(X.p + X.plane(), lo, sizeof(float) * (size_t)n * ld, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

static int ensure_act_graph(dqnb_handle h) {
  if (h->act_graph) return 0;
  std::vector<Op> ops;
  build_skinny_act_ops(h, ops);
  cudaGraph_t graph = nullptr;
  int count = 0;
  DQNB_CUDA(cudaStreamBeginCapture(h->astream, cudaStreamCaptureModeThreadLocal));
  // the call block {n, seq | rows} travels by DMA: one memcpy node at the head of the graph
  cudaError_t ec = cudaMemcpyAsync(h->d_ax, h->h_ax, sizeof(float) * ((size_t)kActHdr + (size_t)kActMaxRows * h->Sp), cudaMemcpyHostToDevice, h->astream);
  int rc = ec == cudaSuccess ? run_ops(h, ops, h->astream, false, &count) : -1;
  cudaError_t e = cudaStreamEndCapture(h->astream, &graph);
  if (ec != cudaSuccess) { if (graph) cudaGraphDestroy(graph); DQNB_CUDA(ec); }
  if (rc) { if (graph) cudaGraphDestroy(graph); return -1; }
  DQNB_CUDA(e);
  DQNB_CUDA(cudaGraphInstantiate(&h->act_graph, graph, 0));
  DQNB_CUDA(cudaGraphDestroy(graph));
  h->act_kernels = count;
  return 0;
}

// waits (spinning on the host-mapped sequence word) for the skinny act call in flight, if any
static int act_wait_done(dqnb_handle h) {
  if (!h->act_skinny_pending) return 0;
  unsigned spins = 0;
  while (h->h_ctl->done != h->act_seq) {
    if ((++spins & 0xfff) == 0) {
      const cudaError_t q = cudaStreamQuery(h->astream);
      if (q == cudaSuccess) break;
      if (q != cudaErrorNotReady) DQNB_CUDA(q);
    }
  }
  if (h->h_ctl->done != h->act_seq) DQNB_FAIL("act call %u did not complete (device reports %u)", h->act_seq, h->h_ctl->done);
  h->act_skinny_pending = false;
  return 0;
}

int dqnb_select_actions_async(dqnb_handle h, int32_t n, const float *states) {
  if (!h || !states || n <= 0) DQNB_FAIL("bad argument");
  if (n > h->cfg.max_act_batch) DQNB_FAIL("SelectActions: batch %d > max_act_batch %d (dqn.cpp:699)", n, h->cfg.max_act_batch);
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  if (act_wait_done(h)) return -1;
  h->act_rows_pending = n;
  if (n <= kActMaxRows && !getenv("DQNB_ACT_GEMM")) {
    // skinny path: own stream and graph, actor snapshot, host-mapped rows: the learner's stream is never touched
    const int S = h->S, Sp = h->Sp;
    float *rows = h->h_ax + kActHdr;
    for (int i = 0; i < n; ++i) memcpy(rows + (size_t)i * Sp, states + (size_t)i * S, sizeof(float) * S);   // padding stays 0
    reinterpret_cast<int *>(h->h_ax)[0] = n;
    reinterpret_cast<unsigned int *>(h->h_ax)[1] = ++h->act_seq;
    if (ensure_act_graph(h)) return -1;
    DQNB_CUDA(cudaGraphLaunch(h->act_graph, h->astream));
    h->launches += h->act_kernels;
    h->act_skinny_pending = true;
    return 0;
  }
  // large batches: the tcgen05 layer kernels on the learner's stream, after whatever is queued there
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  if (stage_rows_split(h, n, states, h->S, h->S, nullptr, 0, h->Xact)) return -1;
  int count = 0;
  if (run_ops(h, h->act_ops, h->stream, false, &count)) return -1;
  h->launches += count;
  DQNB_CUDA(cudaMemcpyAsync(h->h_act_out, h->out16_act, sizeof(float) * (size_t)n * 16, cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

int dqnb_select_actions_wait(dqnb_handle h, int32_t n, float *out10) {
  if (!h || !out10 || n <= 0) DQNB_FAIL("bad argument");
  if (n != h->act_rows_pending) DQNB_FAIL("select_actions_wait: %d rows requested, %d enqueued", n, h->act_rows_pending);
  if (h->act_skinny_pending) {
    if (act_wait_done(h)) return -1;
    for (int i = 0; i < n; ++i) memcpy(out10 + (size_t)i * kActorOut, h->h_ay + (size_t)i * 16, sizeof(float) * kActorOut);
    return 0;
  }
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; ++i) memcpy(out10 + (size_t)i * kActorOut, h->h_act_out + (size_t)i * 16, sizeof(float) * kActorOut);
  return 0;
}

int dqnb_select_actions(dqnb_handle h, int32_t n, const float *states, float *out10) {
  if (dqnb_select_actions_async(h, n, states)) return -1;
  return dqnb_select_actions_wait(h, n, out10);
}

int dqnb_evaluate(dqnb_handle h, int32_t n, const float *states, const float *act10, float *q) {
  if (!h || !states || !act10 || !q || n <= 0) DQNB_FAIL("bad argument");
  if (n > h->cfg.max_act_batch) DQNB_FAIL("Evaluate: batch %d > max_act_batch %d", n, h->cfg.max_act_batch);
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  if (stage_rows_split(h, n, states, h->S, h->S, act10, kActorOut, h->Xeval)) return -1;
  int count = 0;
  if (run_ops(h, h->eval_ops, h->stream, false, &count)) return -1;
  h->launches += count;
  DQNB_CUDA(cudaMemcpyAsync(h->h_act_out, h->out16_act, sizeof(float) * (size_t)n * 16, cudaMemcpyDeviceToHost, h->stream));
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; ++i) q[i] = h->h_act_out[(size_t)i * 16];
  return 0;
}

// ----------------------------------- multi-GPU -------------------------------------------------
}  // extern "C"
// choosing how gradients are exchanged changes the kernel sequence: rebuild it and drop captured graphs
static int set_comm_mode(dqnb_handle h, int mode) {
  if (sync_all(h)) return -1;
  h->comm_mode = mode;
  if (mode != 2) for (int n = 0; n < 2; ++n) h->Gr[n] = h->G[n];
  for (int i = 0; i < 2; ++i) if (h->graph[i]) { cudaGraphExecDestroy(h->graph[i]); h->graph[i] = nullptr; }
  if (build_update_ops(h)) return -1;
  attach_trace(h);
  return 0;
}
extern "C" {
int dqnb_comm_unique_id(void *id128) {
  if (!id128) DQNB_FAIL("null argument");
  if (!nccl().lib || !nccl().GetUniqueId) DQNB_FAIL("libnccl.so.2 not loadable: %s", dlerror());
  int r = nccl().GetUniqueId(id128);
  if (r != 0) DQNB_FAIL("ncclGetUniqueId failed (%d)", r);
  return 0;
}

int dqnb_comm_init(dqnb_handle h, const void *id128) {
  if (!h || !id128) DQNB_FAIL("bad argument");
  if (!nccl().lib || !nccl().CommInitRank) DQNB_FAIL("libnccl.so.2 not loadable");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  Id128 id;
  memcpy(&id, id128, sizeof(id));
  int r = nccl().CommInitRank(&h->comm, h->cfg.world_size, id, h->cfg.rank);
  if (r != 0) DQNB_FAIL("ncclCommInitRank failed: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
  return set_comm_mode(h, 1);
}

// P2P exchange: every rank exports its exchange allocation (64-byte cudaIpcMemHandle) ...
int dqnb_comm_p2p_handle(dqnb_handle h, void *handle64) {
  if (!h || !handle64) DQNB_FAIL("bad argument");
  if (h->cfg.world_size < 2 || !h->xchg) DQNB_FAIL("P2P exchange needs world_size > 1");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  cudaIpcMemHandle_t mh;
  DQNB_CUDA(cudaIpcGetMemHandle(&mh, h->xchg));
  static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &mh, 64);
  return 0;
}

// ... and maps everybody else's (handles: world_size x 64 bytes, in rank order)
int dqnb_comm_p2p_init(dqnb_handle h, const void *handles) {
  if (!h || !handles) DQNB_FAIL("bad argument");
  const int W = h->cfg.world_size;
  if (W < 2 || W > kMaxPeers || !h->xchg) DQNB_FAIL("P2P exchange supports 2..%d ranks", kMaxPeers);
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  P2PTable tab;
  memset(&tab, 0, sizeof(tab));
  for (int p = 0; p < W; ++p) {
    if (p == h->cfg.rank) { tab.base[p] = h->xchg; continue; }
    cudaIpcMemHandle_t mh;
    memcpy(&mh, (const char *)handles + 64 * p, 64);
    void *ptr = nullptr;
    DQNB_CUDA(cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(ptr);
    tab.base[p] = (float *)ptr;
  }
  DQNB_CUDA(cudaMemcpy(h->p2p_tab, &tab, sizeof(tab), cudaMemcpyHostToDevice));
  for (int n = 0; n < 2; ++n) h->Gr[n] = h->xchg + h->x_out[n];
  return set_comm_mode(h, 2);
}

// 0 = ok; 1 = a peer did not show up within the exchange kernel's timeout (results are invalid)
int dqnb_comm_status(dqnb_handle h) {
  if (!h) return -1;
  if (!h->p2p_err) return 0;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  return h->h_p2p_err ? *h->h_p2p_err : 0;
}

int dqnb_sync(dqnb_handle h) {
  if (!h) DQNB_FAIL("null handle");
  DQNB_CUDA(cudaSetDevice(h->cfg.device));
  DQNB_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int64_t dqnb_kernel_launches(dqnb_handle h) { return h ? h->launches : -1; }

// DQNB_TRACE=1 only: per-op timeline of the last update, kTraceSlots int64 per op of the update sequence
// ({kind, branch} in slots 7/6 are filled here; slots 0-5 are globaltimer ns stamps of GEMM ops).
int64_t dqnb_debug_trace(dqnb_handle h, long long *out, int64_t capacity) {
  if (!h || !h->trace || !out) return -1;
  const int64_t n = (int64_t)kTraceSlots * h->trace_ops;
  if (capacity < n) return -1;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  if (cudaMemcpy(out, h->trace, sizeof(long long) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  const int n_ops = std::min(h->trace_ops, (int)h->update_ops.size());
  for (int i = 0; i < n_ops; ++i) {
    const Op &op = h->update_ops[i];
    long long meta = (long long)op.kind * 1000 + op.branch;          // 16 bits
    if (op.kind == Op::GEMM)                                         // grid x:12 y:12 z:8, k-blocks:12
      meta |= ((long long)(op.grid.x & 0xfff) << 16) | ((long long)(op.grid.y & 0xfff) << 28) | ((long long)(op.grid.z & 0xff) << 40) |
              ((long long)((op.gemm.p.K / 32) & 0xfff) << 48);
    out[kTraceSlots * i + 7] = meta;
  }
  return (int64_t)kTraceSlots * n_ops;
}

int64_t dqnb_debug_read(dqnb_handle h, const char *name, float *out, int64_t capacity) {
  if (!h || !name || !out) return -1;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return -1;
  cudaStreamSynchronize(h->stream);
  const std::string n(name);
  const float *src = nullptr;
  int64_t cnt = 0;
  if (n == "y") { src = h->y; cnt = h->B; }
  else if (n == "q") { src = h->q; cnt = h->B; }
  else if (n == "q_next") { src = h->q_next; cnt = h->B; }
  else if (n == "q_pi") { src = h->q_pi; cnt = h->B; }
  else if (n == "d_raw") { src = h->tap_raw; cnt = (int64_t)h->B * kActorOut; }
  else if (n == "d_inv") { src = h->tap_inv; cnt = (int64_t)h->B * kActorOut; }
  else if (n == "a_pi") {
    cnt = (int64_t)h->B * kActorOut;
    if (capacity < cnt) return -1;
    std::vector<float> t((size_t)h->B * 16);
    if (cudaMemcpy(t.data(), h->a16_pi, sizeof(float) * t.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (int i = 0; i < h->B; ++i) memcpy(out + (size_t)i * kActorOut, t.data() + (size_t)i * 16, sizeof(float) * kActorOut);
    return cnt;
  } else if (n == "critic_grad" || n == "actor_grad") {
    const int c = n == "critic_grad";
    const NetGeom &g = c ? h->gC : h->gA;
    cnt = g.caffe_count;
    if (capacity < cnt) return -1;
    if (read_flat(h, g, h->Gr[c], nullptr, out)) return -1;
    return cnt;
  } else if (n == "critic_gnorm" || n == "actor_gnorm") {
    StepState s;
    if (pull_state(h, &s)) return -1;
    if (capacity < 1) return -1;
    out[0] = s.gnorm[n == "critic_gnorm"];
    return 1;
  } else return -1;
  if (capacity < cnt) return -1;
  if (cudaMemcpy(out, src, sizeof(float) * cnt, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return cnt;
}

// ----------------------------------- kernel unit test ------------------------------------------
static long long g_dbg_clk[kTraceSlots * 32];
void dqnb_gemm_test_clocks(long long *out, int n) { for (int i = 0; i < n && i < kTraceSlots * 32; ++i) out[i] = g_dbg_clk[i]; }

int dqnb_gemm_test(int device, int gemm_mode, int a_mn, int b_mn, int M, int N, int K, int splits,
                   const float *A, const float *B, float *C, float *elapsed_ms) {
  if (M % 64 || N % 64 || K % 32 || splits < 1 || splits > kGradSplits) DQNB_FAIL("gemm_test: M,N multiples of 64, K of 32, splits<=8");
  DQNB_CUDA(cudaSetDevice(device));
  DQNB_CUDA(tc_prepare_all());
  const size_t na = (size_t)M * K, nb = (size_t)N * K, nc = (size_t)M * N;
  long long *dclk = nullptr;
  DQNB_CUDA(cudaMalloc(&dclk, sizeof(g_dbg_clk)));
  DQNB_CUDA(cudaMemset(dclk, 0, sizeof(g_dbg_clk)));
  float *dA, *dB, *dAs, *dBs, *dP, *dC;
  DQNB_CUDA(cudaMalloc(&dA, na * 4)); DQNB_CUDA(cudaMalloc(&dB, nb * 4));
  DQNB_CUDA(cudaMalloc(&dAs, 2 * na * 4)); DQNB_CUDA(cudaMalloc(&dBs, 2 * nb * 4));
  DQNB_CUDA(cudaMalloc(&dP, (size_t)splits * nc * 4)); DQNB_CUDA(cudaMalloc(&dC, nc * 4));
  DQNB_CUDA(cudaMemcpy(dA, A, na * 4, cudaMemcpyHostToDevice));
  DQNB_CUDA(cudaMemcpy(dB, B, nb * 4, cudaMemcpyHostToDevice));
  split_kernel<<<(unsigned)((na + 255) / 256), 256>>>(dA, dAs, dAs + na, (long long)na);
  split_kernel<<<(unsigned)((nb + 255) / 256), 256>>>(dB, dBs, dBs + nb, (long long)nb);
  dqnb_config cfg;
  dqnb_default_config(&cfg);
  const int dbg = gemm_mode >> 8;
  gemm_mode &= 0xff;
  cfg.gemm_mode = gemm_mode;
  Op op;
  GemmParams &p = op.gemm.p;
  memset(&p, 0, sizeof(p));
  p.dbg = dbg & 7;
  p.stages = (dbg >> 4) & 0xf;           // 0: deepest ring that fits
  p.bn = (dbg >> 8) ? (dbg >> 8) : 64;
  p.dbg_clk = dclk;
  p.M = M; p.N = N; p.K = K; p.a_mn = a_mn; p.b_mn = b_mn; p.splits = splits; p.epi = EPI_PLAIN;
  p.A = dAs; p.a_plane = (long long)na; p.lda = a_mn ? M : K;
  p.B = dBs; p.b_plane = (long long)nb; p.ldb = b_mn ? N : K;
  p.out = dP; p.out_split_stride = (long long)nc; p.ldo = N;
  if (finish_gemm(cfg, &op)) return -1;
  p.a_ts = ((dbg & 4) && !a_mn) ? 1 : 0;   // unit test of the A-in-TMEM mainloop with the raw epilogue
  cudaEvent_t e0, e1;
  DQNB_CUDA(cudaEventCreate(&e0)); DQNB_CUDA(cudaEventCreate(&e1));
  const int reps = 20;
  for (int r = 0; r < 1 + reps; ++r) {
    if (r == 1) DQNB_CUDA(cudaEventRecord(e0, 0));
    p.dbg_clk = dclk + kTraceSlots * r;            // per-launch timeline slot
    if (gemm_mode == DQNB_GEMM_TCGEN05_3XTF32)
      DQNB_CUDA(launch_k(tc_kernel_for(a_mn, b_mn, p.bn, EPI_PLAIN), op.grid, dim3(TC_THREADS), (size_t)tc_smem_for(p.bn, p.stages), (cudaStream_t)0, op.gemm));
    else gemm_simt_kernel<<<op.grid, 256>>>(op.gemm.p);
  }
  DQNB_CUDA(cudaEventRecord(e1, 0));
  DQNB_CUDA(cudaGetLastError());
  DQNB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  DQNB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  if (elapsed_ms) *elapsed_ms = ms / reps;
  sum_planes_kernel<<<(unsigned)((nc + 255) / 256), 256>>>(dP, (long long)nc, splits, dC, (long long)nc);
  DQNB_CUDA(cudaGetLastError());
  DQNB_CUDA(cudaMemcpy(C, dC, nc * 4, cudaMemcpyDeviceToHost));
  DQNB_CUDA(cudaMemcpy(g_dbg_clk, dclk, sizeof(g_dbg_clk), cudaMemcpyDeviceToHost));
  cudaFree(dclk);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dB); cudaFree(dAs); cudaFree(dBs); cudaFree(dP); cudaFree(dC);
  return 0;
}

}  // extern "C"
