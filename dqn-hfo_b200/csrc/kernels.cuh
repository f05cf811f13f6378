// kernels.cuh — the non-GEMM kernels of one UpdateActorCritic (dqn.cpp:828-972): index sampling,
// replay gather, linear heads, TD target / Euclidean loss, inverting gradients, bias gradients,
// gradient reduction + global-norm clip, Adam + soft target update (+ the actor snapshot of the act path), the skinny-M
// act-path kernels (SelectActionGreedily, dqn.cpp:734-766) and the multi-GPU gradient exchange over NVLink peer memory.
// All HBM/L2/latency-bound; loads are coalesced along the contiguous (feature) axis and vectorised where the layout allows.
#pragma once
#include "common.cuh"

namespace dqnb {

// Programmatic dependent launch: every kernel of the update sequence is launched with the
// programmatic-stream-serialization attribute, so its CTAs may start (and run their prologue)
// before the previous kernel has drained.  pdl_wait() blocks until the prerequisite grid has
// completed and its writes are visible; nothing before it may touch global memory.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define DQNB_PDL_PROLOGUE() do { pdl_wait(); pdl_launch_dependents(); } while (0)

// DQNB_TRACE=1 timeline (scripts/trace_update.py): every kernel of the update records, over all its CTAs,
// [0] the earliest and [1] the latest time a CTA passed griddepcontrol.wait and [2] the latest CTA exit
// (globaltimer ns; slots preset to ~0 by the host, so min is an unsigned and max a signed atomic).
constexpr int kTraceSlots = 16;      // int64 slots per op in the trace buffer
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_begin(long long *t) {
  if (t && threadIdx.x == 0) {
    const unsigned long long now = gtime_ns();
    atomicMin(reinterpret_cast<unsigned long long *>(t), now);
    atomicMax(t + 1, (long long)now);
  }
}
__device__ __forceinline__ void trace_end(long long *t) {
  if (t && threadIdx.x == 0) atomicMax(t + 2, (long long)gtime_ns());
}

constexpr int kGradSplits = 8;      // planes of the gradient-partial buffer
constexpr int kMaxSegs = 2 * 8 + 4; // parameter blobs per net

// Device-resident scalars so that one captured graph serves every update.
struct StepState {
  int actor_iter, critic_iter;      // Solver::iter() of each net (dqn.hpp:129-130)
  unsigned long long step;          // update counter: keys the index sampler
  int ring_head, ring_size;         // replay ring: physical index of the oldest row, fill
  float step_critic, step_actor;    // lr * Adam bias correction for this update
  int do_soft;                      // dqn.cpp:967
  int result_slot;                  // where finalize writes (critic_loss, avg_q)
  float gnorm[2];                   // [actor, critic] L2 norm seen by ClipGradients
};

struct HyperParams {
  double gamma, beta;
  float tau;
  int soft_update_freq;
  float actor_lr, critic_lr, beta1, beta2, eps, clip;
  float inv_batch_global;           // 1 / (batch * world_size): EuclideanLoss 1/N
};

// SampleTransitionsFromMemory (dqn.cpp:501-509): B uniform draws with replacement in [0,size)
__global__ void sample_kernel(const StepState *st, unsigned long long seed, unsigned long long step, int B, int32_t *idx) {
  DQNB_PDL_PROLOGUE();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B) return;
  const int size = st->ring_size;
  idx[n] = size > 0 ? sample_index(seed, step, (uint32_t)n, (uint32_t)size) : 0;
}

// -------------------------------------------------------------------------------------------
// Replay gather (dqn.cpp:859-887): one block per minibatch row, threads along the feature axis.
struct GatherArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  StepState *st;
  int32_t *idx;                // deque indices of the minibatch rows
  int sample;                  // 1: draw idx on the device (SampleTransitionsFromMemory, dqn.cpp:501-509)
  unsigned long long seed;
  unsigned long long step;     // update counter keying the sampler (host mirror of StepState::step: the gather of update
                               // t may run beside update t-1, whose last kernel advances the device counter)
  HyperParams hp;
  const float *ring_s, *ring_sn, *ring_misc;   // one row-interleaved ring: [state Sp | next Sp | misc 16]
  int rw;                      // floats per ring row
  int cap, B, Bp, S, Sp, Kc;
  float *Xs, *Xsn;           // [2][Bp][Sp]   actor inputs: s, s'
  float *Xc, *Xct, *Xcp;     // [2][Bp][Kc]   critic inputs: (s,a,p), (s',.), (s,.)
  float *reward, *mc, *term; // [Bp]
};

__global__ void __launch_bounds__(128) gather_kernel(const GatherArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  const int n = blockIdx.x;
  const bool valid = n < a.B;
  long long phys = 0;
  float misc_term = 1.f;
  if (valid) {
    int id;
    if (a.sample) {
      const int size = a.st->ring_size;
      id = size > 0 ? sample_index(a.seed, a.step, (uint32_t)n, (uint32_t)size) : 0;
      if (threadIdx.x == 0) a.idx[n] = id;
    } else {
      id = a.idx[n];
    }
    phys = ((long long)a.st->ring_head + id) % a.cap;
    misc_term = a.ring_misc[phys * a.rw + 12];
  }
  const bool has_next = valid && misc_term == 0.f;
  const long long pS = (long long)a.Bp * a.Sp, pK = (long long)a.Bp * a.Kc;
  for (int c = threadIdx.x; c < a.Kc; c += blockDim.x) {
    float sv = 0.f, snv = 0.f;
    if (c < a.Sp) {
      if (valid) sv = a.ring_s[phys * a.rw + c];
      if (has_next) snv = a.ring_sn[phys * a.rw + c];
      const long long o = (long long)n * a.Sp + c;
      float h = tf32_hi(sv);
      a.Xs[o] = h; a.Xs[o + pS] = sv - h;
      h = tf32_hi(snv);
      a.Xsn[o] = h; a.Xsn[o + pS] = snv - h;
    }
    float xc = 0.f;
    if (c < a.S) xc = sv;
    else if (c < a.S + kActorOut && valid) xc = a.ring_misc[phys * a.rw + (c - a.S)];
    const long long o = (long long)n * a.Kc + c;
    float h = tf32_hi(xc);
    a.Xc[o] = h; a.Xc[o + pK] = xc - h;
    const float xt = c < a.S ? snv : 0.f;   // (s', a') : a' filled by the target-actor head
    h = tf32_hi(xt);
    a.Xct[o] = h; a.Xct[o + pK] = xt - h;
    const float xp = c < a.S ? sv : 0.f;    // (s, a_pi): a_pi filled by the actor head
    h = tf32_hi(xp);
    a.Xcp[o] = h; a.Xcp[o + pK] = xp - h;
  }
  if (threadIdx.x == 0) {
    a.reward[n] = valid ? a.ring_misc[phys * a.rw + 10] : 0.f;
    a.mc[n] = valid ? a.ring_misc[phys * a.rw + 11] : 0.f;
    a.term[n] = misc_term;
  }
  trace_end(a.trace);
}

// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Linear heads (dqn.cpp:426-427, :450): out[n][j] = h[n] . W[j] + b[j], one warp per row.
struct HeadArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  const float *H; long long h_plane; int ldh; int Kp;   // top tower activation [2][rows][ldh]
  const float *W; long long w_plane;                    // head weights [2][..] rows of Kp
  const float *bias; long long b_plane;
  int J;                       // rows of W in use (10 actor, 1 critic)
  int rows;                    // valid rows (B, or n for the act path)
  float *out16;                // [rows_pad][16] fp32 (nullable)
  float *dst; long long dst_plane; int ldd; int dst_col;  // also scatter split(out) into a critic input
};

constexpr int kHeadRowsPerBlock = 8;    // one warp per minibatch row
constexpr int kHeadMaxK = 4096;

// The per-row head kernels are latency-bound (a few KB per row): every lane issues all of its
// float4 loads for a 128-wide K chunk before using any of them, so a row costs about one trip to L2.
__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 add4(const float4 a, const float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ void store_split4(float *hi, float *lo, const float4 v) {
  const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  *reinterpret_cast<float4 *>(hi) = h;
  *reinterpret_cast<float4 *>(lo) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}
__device__ __forceinline__ float4 relu_bwd4(const float4 v, const float4 y) {
  return make_float4(v.x * (y.x > 0.f ? 1.f : kNegSlope), v.y * (y.y > 0.f ? 1.f : kNegSlope),
                     v.z * (y.z > 0.f ? 1.f : kNegSlope), v.w * (y.w > 0.f ? 1.f : kNegSlope));
}

__global__ void __launch_bounds__(256) head_fwd_kernel(const HeadArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * kHeadRowsPerBlock + warp;
  if (n >= a.rows) return;
  const float *hh = a.H + (long long)n * a.ldh;
  float acc[kActorOut];
#pragma unroll
  for (int j = 0; j < kActorOut; ++j) acc[j] = 0.f;
  for (int k = lane * 4; k < a.Kp; k += 128) {      // Kp is a multiple of 64: the tail chunk is half-used
    const float4 x = add4(ld4(hh + k), ld4(hh + k + a.h_plane));
    float4 w[kActorOut];
#pragma unroll
    for (int j = 0; j < kActorOut; ++j)
      if (j < a.J) w[j] = add4(ld4(a.W + (long long)j * a.Kp + k), ld4(a.W + (long long)j * a.Kp + k + a.w_plane));
#pragma unroll
    for (int j = 0; j < kActorOut; ++j)
      if (j < a.J) acc[j] = dot4(x, w[j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < kActorOut; ++j)
    if (j < a.J) acc[j] = warp_sum(acc[j]);
  if (lane < a.J) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < kActorOut; ++j)
      if (j == lane) v = acc[j];
    v += a.bias[lane] + a.bias[lane + a.b_plane];
    if (a.out16) a.out16[(long long)n * 16 + lane] = v;
    if (a.dst) {
      const long long o = (long long)n * a.ldd + a.dst_col + lane;
      const float h = tf32_hi(v);
      a.dst[o] = h; a.dst[o + a.dst_plane] = v - h;
    }
  }
  trace_end(a.trace);
}

// Critic head, fused per minibatch row (one warp per row):
//   q = h . w_q + b_q                                         (q_values_layer, dqn.cpp:450)
//   TARGET: y = beta*mc + (1-beta)*(terminal ? r : r + gamma*q)   in double, dqn.cpp:892-900
//   LOSS  : EuclideanLoss forward/backward: dq = (q - y)/N, loss partial sum (q-y)^2
//   POLICY: dq = -1 (dqn.cpp:918-921), avg-q partial sum q
//   LOSS/POLICY also write the head backward into the tower top, masked by that layer's ReLU':
//   dZ[n][k] = dq * w_q[k] * relu'(h[n][k])
//   TARGET_LOSS: both of the above in one launch (the target tower's top activation H2 / head W2 give y, then the
//   online tower's H / W the loss): one kernel boundary less on the critical chain of the update
enum { QMODE_TARGET = 0, QMODE_LOSS = 1, QMODE_POLICY = 2, QMODE_TARGET_LOSS = 3 };
struct CriticHeadArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  int mode, B, rows_pad;
  const float *H; long long h_plane; int ldh; int Kp;
  const float *W; long long w_plane; const float *bias; long long b_plane;
  // TARGET_LOSS only: the target critic's top activation and head (H / W above are the online critic's)
  const float *H2; long long h2_plane; const float *W2; long long w2_plane; const float *bias2; long long b2_plane;
  float *q_tap2;               // TARGET_LOSS: q_next
  const float *reward, *mc, *term;
  float *y;                    // TARGET: out ; LOSS: in ; TARGET_LOSS: out (tap)
  float *q_tap;                // TARGET: q_next ; LOSS / TARGET_LOSS: q ; POLICY: q_pi
  float *d16;                  // LOSS: dq for the head weight gradient (column 0 of [rows][16])
  float *dZ; long long dz_plane;
  double *part;                // per-block partial: LOSS sum (q-y)^2 ; POLICY sum q
  HyperParams hp;
};
__global__ void __launch_bounds__(256) critic_head_kernel(const CriticHeadArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  __shared__ double red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * kHeadRowsPerBlock + warp;
  double contrib = 0.0;
  if (n < a.rows_pad) {
    const float *hh = a.H + (long long)n * a.ldh;
    float dq = 0.f;
    if (n < a.B) {
      float acc = 0.f;
      for (int k = lane * 4; k < a.Kp; k += 128)
        acc = dot4(add4(ld4(hh + k), ld4(hh + k + a.h_plane)), add4(ld4(a.W + k), ld4(a.W + k + a.w_plane)), acc);
      const float q = warp_sum(acc) + (a.bias[0] + a.bias[a.b_plane]);
      float y_here = 0.f;
      if (a.mode == QMODE_TARGET_LOSS) {
        const float *h2 = a.H2 + (long long)n * a.ldh;
        float acc2 = 0.f;
        for (int k = lane * 4; k < a.Kp; k += 128)
          acc2 = dot4(add4(ld4(h2 + k), ld4(h2 + k + a.h2_plane)), add4(ld4(a.W2 + k), ld4(a.W2 + k + a.w2_plane)), acc2);
        const float qn = warp_sum(acc2) + (a.bias2[0] + a.bias2[a.b2_plane]);
        const bool terminal = a.term[n] != 0.f;
        const float off = terminal ? a.reward[n] : (float)((double)a.reward[n] + a.hp.gamma * (double)qn);
        y_here = (float)(a.hp.beta * (double)a.mc[n] + (1 - a.hp.beta) * (double)off);
        if (lane == 0) { a.y[n] = y_here; a.q_tap2[n] = terminal ? 0.f : qn; }
      }
      if (a.mode == QMODE_TARGET) {
        if (lane == 0) {
          const bool terminal = a.term[n] != 0.f;
          // float off = terminal ? r : r + gamma_*q' (double expr narrowed); float target = beta*on + (1-beta)*off
          const float off = terminal ? a.reward[n] : (float)((double)a.reward[n] + a.hp.gamma * (double)q);
          a.y[n] = (float)(a.hp.beta * (double)a.mc[n] + (1 - a.hp.beta) * (double)off);
          a.q_tap[n] = terminal ? 0.f : q;
        }
      } else if (a.mode == QMODE_LOSS || a.mode == QMODE_TARGET_LOSS) {
        const float diff = q - (a.mode == QMODE_LOSS ? a.y[n] : y_here);
        dq = __fmul_rn(a.hp.inv_batch_global, diff);
        if (lane == 0) { a.d16[(long long)n * 16] = dq; a.q_tap[n] = q; contrib = (double)diff * (double)diff; }
      } else {
        dq = -1.0f;
        if (lane == 0) { a.q_tap[n] = q; contrib = (double)q; }
      }
    }
    if (a.mode != QMODE_TARGET) {
      float *zh = a.dZ + (long long)n * a.ldh, *zl = zh + a.dz_plane;
      for (int k = lane * 4; k < a.Kp; k += 128) {        // operands are L1 hits from the dot product above
        const float4 y = add4(ld4(hh + k), ld4(hh + k + a.h_plane));
        const float4 w = add4(ld4(a.W + k), ld4(a.W + k + a.w_plane));
        store_split4(zh + k, zl + k, relu_bwd4(make_float4(dq * w.x, dq * w.y, dq * w.z, dq * w.w), y));
      }
    }
  }
  if (lane == 0) red[warp] = contrib;
  __syncthreads();
  if (threadIdx.x == 0 && a.mode != QMODE_TARGET) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    a.part[blockIdx.x] = s;
  }
  trace_end(a.trace);
}

// Actor heads backward, fused per row: inverting gradients (dqn.cpp:927-957) on the critic's input
// diff columns [S, S+10), then Split-sum of the two heads' bottom diffs and the ReLU' of the tower top:
//   dZ[n][k] = (sum_j d10[n][j] * W[j][k]) * relu'(h[n][k])
struct ActorHeadBwdArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  int B, rows_pad, S, ldin;
  const float *d_in;           // [din_splits][Bp][ldin] fp32 split-K partials of dL/d(critic input)
  int din_splits; long long din_stride;
  const float *a16;            // actor outputs [Bp][16]
  float *d16;                  // out: top diffs of the actor heads [Bp][16] (for the head weight gradient)
  float *tap_raw, *tap_inv;    // [Bp][10] debug taps
  const float *W; long long w_plane; int Kp;
  const float *H; long long h_plane; int ldh;
  float *dZ; long long dz_plane;
};
__global__ void __launch_bounds__(256) actor_head_bwd_kernel(const ActorHeadBwdArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * kHeadRowsPerBlock + warp;
  if (n >= a.rows_pad) return;
  float diff = 0.f;
  if (n < a.B && lane < kActorOut) {
    float part[kGradSplits];
#pragma unroll
    for (int sp = 0; sp < kGradSplits; ++sp)       // all partial planes in flight, then summed in plane order
      if (sp < a.din_splits) part[sp] = a.d_in[sp * a.din_stride + (long long)n * a.ldin + a.S + lane];
#pragma unroll
    for (int sp = 0; sp < kGradSplits; ++sp)
      if (sp < a.din_splits) diff += part[sp];
    a.tap_raw[n * kActorOut + lane] = diff;
    const float output = a.a16[(long long)n * 16 + lane];
    float mn, mx;
    if (lane < kActionSize) { mn = -1.0f; mx = 1.0f; }
    else if (lane == kActionSize + 0 || lane == kActionSize + 4) { mn = 0.f; mx = 100.f; }
    else { mn = -180.f; mx = 180.f; }
    if (diff < 0) diff *= (mx - output) / (mx - mn);
    else if (diff > 0) diff *= (output - mn) / (mx - mn);
    a.tap_inv[n * kActorOut + lane] = diff;
  }
  if (lane < 16) a.d16[(long long)n * 16 + lane] = diff;
  float d[kActorOut];
#pragma unroll
  for (int j = 0; j < kActorOut; ++j) d[j] = __shfl_sync(0xffffffffu, diff, j);
  const float *hh = a.H + (long long)n * a.ldh;
  float *zh = a.dZ + (long long)n * a.ldh, *zl = zh + a.dz_plane;
  for (int k = lane * 4; k < a.Kp; k += 128) {
    const float4 y = add4(ld4(hh + k), ld4(hh + k + a.h_plane));
    float4 w[kActorOut];
#pragma unroll
    for (int j = 0; j < kActorOut; ++j)
      w[j] = add4(ld4(a.W + (long long)j * a.Kp + k), ld4(a.W + (long long)j * a.Kp + k + a.w_plane));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kActorOut; ++j) {
      acc.x = fmaf(d[j], w[j].x, acc.x); acc.y = fmaf(d[j], w[j].y, acc.y);
      acc.z = fmaf(d[j], w[j].z, acc.z); acc.w = fmaf(d[j], w[j].w, acc.w);
    }
    store_split4(zh + k, zl + k, relu_bwd4(acc, y));
  }
  if (warp == 0) trace_end(a.trace);
}

// Head weight/bias gradients: dW[j][k] = sum_n d16[n][j] h[n][k] ; db[j] = sum_n d16[n][j].
// grid (Kp/32, kGradSplits); 256 threads = 8 row sub-groups x 32 columns (many small blocks: the kernel
// sits at the head of the gradient side chain, so its latency matters more than its efficiency).
constexpr int kRedSub = 8;
constexpr int kHbwCols = 32;
struct HeadBwdWArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  const float *d16; int J;
  const float *H; long long h_plane; int ldh; int Kp;
  int rows_pad;
  float *gpart; long long gpart_stride; long long hw_off, hb_off;
};
__global__ void __launch_bounds__(256) head_bwd_w_kernel(const HeadBwdWArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  __shared__ float red[kRedSub][kActorOut][kHbwCols];
  __shared__ float sd[128 * 16];             // up to 128 rows of d16 per chunk
  const int col = threadIdx.x & (kHbwCols - 1), sub = threadIdx.x / kHbwCols, split = blockIdx.y;
  const int k = blockIdx.x * kHbwCols + col;
  const int per = a.rows_pad / kGradSplits;  // rows of this split (multiple of 16)
  const int r0 = split * per;
  float acc[kActorOut];
  float bias_acc = 0.f;                      // thread (sub 0, col j < J): sum_n d16[n][j]
#pragma unroll
  for (int j = 0; j < kActorOut; ++j) acc[j] = 0.f;
  for (int c0 = 0; c0 < per; c0 += 128) {    // chunks of <= 128 rows
    const int rows = min(128, per - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * 16; i += blockDim.x) sd[i] = a.d16[(long long)(r0 + c0) * 16 + i];
    __syncthreads();
    if (sub == 0 && col < a.J)
      for (int n = 0; n < rows; ++n) bias_acc += sd[n * 16 + col];
    if (k < a.Kp) {
#pragma unroll 8
      for (int n = sub; n < rows; n += kRedSub) {
        const long long ho = (long long)(r0 + c0 + n) * a.ldh + k;
        const float x = a.H[ho] + a.H[ho + a.h_plane];
#pragma unroll
        for (int j = 0; j < kActorOut; ++j)
          if (j < a.J) acc[j] = fmaf(sd[n * 16 + j], x, acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kActorOut; ++j) red[sub][j][col] = acc[j];
  __syncthreads();
  float *g = a.gpart + (long long)split * a.gpart_stride;
  for (int j = sub; j < a.J; j += kRedSub) {   // thread -> (j = sub, sub+8, .. ; col)
    if (k < a.Kp) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < kRedSub; ++t) s += red[t][j][col];
      g[a.hw_off + (long long)j * a.Kp + k] = s;
    }
  }
  if (blockIdx.x == 0 && sub == 0 && col < a.J) g[a.hb_off + col] = bias_acc;
  trace_end(a.trace);
}

// Bias gradients of the tower layers: db_l[c] = sum_n dZ_l[n][c] (Caffe: gemv(dY^T, ones)).
// grid (column blocks of 32 over all layers, kGradSplits row slices); 1024 threads = 32 row sub-groups x 32
// columns: every warp reads 128 contiguous bytes of one row, and a thread adds rows_pad / (8 * 32) rows, so
// the kernel is many short independent chains (it sits beside the tail of the backward pass).
constexpr int kCsCols = 32, kCsSub = 32;
struct ColsumArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  int n_layers, rows_pad;
  const float *dZ[8]; long long plane[8]; int ld[8]; int Np[8]; long long b_off[8];
  int blk_begin[9];            // prefix sums of Np/32 column blocks
  float *gpart; long long gpart_stride;
};
__global__ void __launch_bounds__(1024) colsum_kernel(const ColsumArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  __shared__ float red[kCsSub][kCsCols + 1];
  int l = 0;
  while (l + 1 < a.n_layers && (int)blockIdx.x >= a.blk_begin[l + 1]) ++l;
  const int col = threadIdx.x & (kCsCols - 1), sub = threadIdx.x / kCsCols, split = blockIdx.y;
  const int c = ((int)blockIdx.x - a.blk_begin[l]) * kCsCols + col;
  const int per = a.rows_pad / kGradSplits;
  const int r0 = split * per, r1 = r0 + per;
  const float *zh = a.dZ[l], *zl = a.dZ[l] + a.plane[l];
  const int ld = a.ld[l];
  float acc = 0.f;
  if (c < a.Np[l]) {
#pragma unroll 4
    for (int n = r0 + sub; n < r1; n += kCsSub) {
      const long long o = (long long)n * ld + c;
      acc += zh[o] + zl[o];
    }
  }
  red[sub][col] = acc;
  __syncthreads();
  if (sub == 0 && c < a.Np[l]) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < kCsSub; ++t) s += red[t][col];
    a.gpart[(long long)split * a.gpart_stride + a.b_off[l] + c] = s;
  }
  trace_end(a.trace);
}

// peer table + system-scope flag helpers of the multi-GPU gradient exchange (p2p_allreduce_kernel below); the gradient
// reduction in front of it already tells the peers that this rank's gradient is complete
constexpr int kMaxPeers = 8;
struct P2PTable {                 // device-resident, filled by dqnb_comm_p2p_init
  float *base[kMaxPeers];         // peer exchange allocations (IPC-mapped); base[rank] is our own
};
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_volatile_f4(const float *p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// -------------------------------------------------------------------------------------------
// Gradient reduction over split planes + sum of squares (SGDSolver::ClipGradients numerator).
struct SegTable {
  int n;
  long long begin[kMaxSegs], end[kMaxSegs];
  int nsplit[kMaxSegs];
  // where the partial planes of a segment live: plane p of parameter i at src[s] + p * stride[s] + (i - begin[s]).
  // Weight gradients: the split-K planes of the dW GEMMs; bias gradients: one plane per 128-row block of the
  // minibatch, written by the epilogue that produced dZ (gemm.cuh EPI_DX) or by colsum_kernel.
  const float *src[kMaxSegs]; long long stride[kMaxSegs];
};
struct ReduceArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  SegTable segs;
  long long flat;
  float *G;                    // [flat + 4]; tail[0] carries the scalar of this pass
  float *norm_part;            // per-block sum of squares
  const double *scal_part; int n_scal; float scal_scale;   // tail[0] = scale * sum(scal_part)
  int do_reduce, do_sumsq;
  // the critic's reduction (first optimiser kernel of an update) also refreshes the per-update Adam scalars
  int do_prep; StepState *st; HyperParams hp;
  // data-parallel runs with the peer-memory exchange: the last block releases flag A ("my gradient is complete") to
  // every peer, one kernel boundary earlier than the exchange kernel could
  const P2PTable *tab; int world, rank, net; long long flag_off;
  const unsigned int *epoch; unsigned int *flag_ticket;
};
__global__ void __launch_bounds__(256) reduce_kernel(const ReduceArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  float ss = 0.f;
  if (i < a.flat) {
    float4 g;
    if (a.do_reduce) {
      int s = 0;
      while (s + 1 < a.segs.n && i >= a.segs.end[s]) ++s;
      const int ns = a.segs.nsplit[s];
      const float *src = a.segs.src[s] + (i - a.segs.begin[s]);
      const long long stride = a.segs.stride[s];
      // every plane's load (of a batch of kGradSplits) is issued before the first add (a load-add-load-add loop is
      // one L2 round trip per plane); the summation order stays plane 0, 1, 2, ..
      g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p0 = 0; p0 < ns; p0 += kGradSplits) {
        float4 t[kGradSplits];
#pragma unroll
        for (int p = 0; p < kGradSplits; ++p)
          if (p0 + p < ns) t[p] = *reinterpret_cast<const float4 *>(src + (long long)(p0 + p) * stride);
#pragma unroll
        for (int p = 0; p < kGradSplits; ++p)
          if (p0 + p < ns) { g.x += t[p].x; g.y += t[p].y; g.z += t[p].z; g.w += t[p].w; }
      }
      *reinterpret_cast<float4 *>(a.G + i) = g;
    } else {
      g = *reinterpret_cast<const float4 *>(a.G + i);
    }
    ss = g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
  }
  if (a.do_sumsq) {
    __shared__ float red[8];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w];
      a.norm_part[blockIdx.x] = s;
    }
  }
  if (a.do_reduce && blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int j = 0; j < a.n_scal; ++j) s += a.scal_part[j];
    a.G[a.flat] = (float)(s * (double)a.scal_scale);
  }
  if (a.do_prep && blockIdx.x == 0 && threadIdx.x == 32) {
    // AdamSolver::ComputeUpdateValue: t = iter+1, correction evaluated in double (std::pow(float,int) promotes);
    // consumed by both adam_kernel launches of this update
    StepState *st = a.st;
    const int tc = st->critic_iter + 1, ta = st->actor_iter + 1;
    const double b1 = (double)a.hp.beta1, b2 = (double)a.hp.beta2;
    const float cc = (float)(sqrt(1.0 - pow(b2, (double)tc)) / (1.0 - pow(b1, (double)tc)));
    const float ca = (float)(sqrt(1.0 - pow(b2, (double)ta)) / (1.0 - pow(b1, (double)ta)));
    st->step_critic = __fmul_rn(a.hp.critic_lr, cc);
    st->step_actor = __fmul_rn(a.hp.actor_lr, ca);
    const int mx = max(ta, tc);       // max_iter() after both solvers stepped (dqn.cpp:967)
    st->do_soft = (a.hp.soft_update_freq > 0 && mx % a.hp.soft_update_freq == 0) ? 1 : 0;
  }
  if (a.tab) {
    // barrier, then ONE fence by the thread that takes the ticket (the grid-synchronisation pattern of cooperative
    // groups: the barrier orders the block's stores before thread 0's fence, which is cumulative); a fence per thread cost
    // the 750-block reduction 5 us
    __shared__ int s_last_blk;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last_blk = (atomicAdd(a.flag_ticket + a.net, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last_blk) {
      __threadfence();
      if ((int)threadIdx.x < a.world) {
        const unsigned int e = a.epoch[a.net] + 1u;
        unsigned int *pf = reinterpret_cast<unsigned int *>(a.tab->base[threadIdx.x] + a.flag_off) + a.net * kMaxPeers;
        st_release_sys_u32(pf + a.rank, e);
      }
      if (threadIdx.x == 0) a.flag_ticket[a.net] = 0;
    }
  }
  trace_end(a.trace);
}

// ClipGradients scale + AdamSolver::ComputeUpdateValue + Net::Update (+ SoftUpdateNet dqn.cpp:1085-1096)
struct AdamArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  long long flat;
  const float *G;
  const float *norm_part; int n_norm;
  float *M, *V;
  float *P; long long p_plane;      // online params [2][flat]
  float *T; long long t_plane;      // target params [2][flat]
  const StepState *st; StepState *st_out;
  int is_critic;
  HyperParams hp;
  // set on the last optimiser launch of an update: the block that finishes last publishes
  // (critic_loss, avg_q) and advances the iteration / sampler counters (formerly finalize_kernel)
  int finalize; unsigned int *ticket; const float *g_critic_tail, *g_actor_tail; float *results; int max_slots;
  unsigned long long *done;         // mapped pinned word: number of finished updates (dqnb_results polls it)
  // actor only: fp32 snapshot of the updated weights for the act path, written into buffer 1 - *act_cur;
  // the finalize step flips *act_cur once every block of the update has retired
  float *snap; long long snap_stride; unsigned int *act_cur;
  // data-parallel runs: sticky flag of the gradient exchange (a peer missed its timeout).  Once set the reduced
  // gradient is not trustworthy: no parameter, moment or iteration counter moves any more, the update only reports
  // NaN results and advances the sequence number the host waits on.
  const int *comm_err;
};
__device__ __forceinline__ void adam_finalize(const AdamArgs &a, bool failed) {
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence();                  // after the barrier: cumulative over the block's stores (grid-sync pattern)
  const unsigned int done = atomicAdd(a.ticket, 1u);
  if (done != gridDim.x - 1) return;
  *a.ticket = 0;
  StepState *st = a.st_out;
  const int slot = (int)(st->step % (unsigned long long)a.max_slots);   // host-visible ring (mapped pinned memory)
  a.results[2 * slot + 0] = failed ? __int_as_float(0x7fc00000) : a.g_critic_tail[0];     // critic_loss (dqn.cpp:905)
  a.results[2 * slot + 1] = failed ? __int_as_float(0x7fc00000) : a.g_actor_tail[0];      // avg_q       (dqn.cpp:915)
  if (!failed) {
    st->critic_iter += 1;                           // Solver::Step ++iter_ (dqn.cpp:904)
    st->actor_iter += 1;                            // set_iter(iter+1)     (dqn.cpp:965)
  }
  st->step += 1;
  if (a.act_cur && !failed) { *a.act_cur ^= 1u; __threadfence(); }   // the snapshot the actor's blocks just wrote becomes current
  if (a.done) {
    __threadfence_system();                         // results first, then the counter the host spins on
    *reinterpret_cast<volatile unsigned long long *>(a.done) = st->step;
  }
}
__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  __shared__ float s_scale;
  __shared__ double red[8];
  if (a.comm_err && *reinterpret_cast<const volatile int *>(a.comm_err) != 0) {   // uniform over the grid: set before this launch
    if (a.finalize) adam_finalize(a, true);
    trace_end(a.trace);
    return;
  }
  {
    double s = 0.0;
    for (int j = threadIdx.x; j < a.n_norm; j += blockDim.x) s += (double)a.norm_part[j];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      const float l2 = sqrtf((float)t);
      s_scale = (a.hp.clip >= 0.f && l2 > a.hp.clip) ? a.hp.clip / l2 : 1.0f;
      if (blockIdx.x == 0) a.st_out->gnorm[a.is_critic] = l2;
    }
    __syncthreads();
  }
  // (Requesting the operands before the norm reduction was tried: 48 instead of 40 registers, so the 804 blocks
  // of the critic no longer fit in one wave - 9.8 us instead of 7.7.  The kernel moves 52 B per parameter, mostly
  // from HBM: it is bandwidth-bound as it stands.)
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= a.flat) {
    if (a.finalize) adam_finalize(a, false);
    trace_end(a.trace);
    return;
  }
  const float scale = s_scale;
  const float step = a.is_critic ? a.st->step_critic : a.st->step_actor;
  const int do_soft = a.st->do_soft;
  const float b1 = a.hp.beta1, b2 = a.hp.beta2, a1 = 1.f - a.hp.beta1, a2 = 1.f - a.hp.beta2;
  const float keep = 1 - a.hp.tau;
  const float4 g4 = *reinterpret_cast<const float4 *>(a.G + i);
  float4 m4 = *reinterpret_cast<const float4 *>(a.M + i);
  float4 v4 = *reinterpret_cast<const float4 *>(a.V + i);
  const float4 ph = *reinterpret_cast<const float4 *>(a.P + i);
  const float4 pl = *reinterpret_cast<const float4 *>(a.P + a.p_plane + i);
  float g[4] = {g4.x, g4.y, g4.z, g4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
  float w[4] = {ph.x + pl.x, ph.y + pl.y, ph.z + pl.z, ph.w + pl.w};
  float wh[4], wl[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float gi = scale != 1.0f ? __fmul_rn(g[t], scale) : g[t];   // Blob::scale_diff
    m[t] = __fadd_rn(__fmul_rn(m[t], b1), __fmul_rn(a1, gi));         // axpby = sscal + saxpy
    v[t] = __fadd_rn(__fmul_rn(v[t], b2), __fmul_rn(a2, __fmul_rn(gi, gi)));
    const float den = __fadd_rn(sqrtf(v[t]), a.hp.eps);
    const float d = __fmul_rn(step, m[t] / den);
    w[t] = __fsub_rn(w[t], d);                                         // Blob::Update
    wh[t] = tf32_hi(w[t]);
    wl[t] = w[t] - wh[t];
  }
  *reinterpret_cast<float4 *>(a.M + i) = make_float4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<float4 *>(a.V + i) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4 *>(a.P + i) = make_float4(wh[0], wh[1], wh[2], wh[3]);
  *reinterpret_cast<float4 *>(a.P + a.p_plane + i) = make_float4(wl[0], wl[1], wl[2], wl[3]);
  if (a.snap)
    *reinterpret_cast<float4 *>(a.snap + (long long)((*a.act_cur & 1u) ^ 1u) * a.snap_stride + i) = make_float4(w[0], w[1], w[2], w[3]);
  if (do_soft) {
    const float4 th = *reinterpret_cast<const float4 *>(a.T + i);
    const float4 tl = *reinterpret_cast<const float4 *>(a.T + a.t_plane + i);
    float tt[4] = {th.x + tl.x, th.y + tl.y, th.z + tl.z, th.w + tl.w};
    float oh[4], ol[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      // caffe_cpu_axpby(N, tau, from, 1-tau, to): to *= (1-tau); to += tau*from
      const float x = __fadd_rn(__fmul_rn(tt[t], keep), __fmul_rn(a.hp.tau, w[t]));
      oh[t] = tf32_hi(x);
      ol[t] = x - oh[t];
    }
    *reinterpret_cast<float4 *>(a.T + i) = make_float4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<float4 *>(a.T + a.t_plane + i) = make_float4(ol[0], ol[1], ol[2], ol[3]);
  }
  if (a.finalize) adam_finalize(a, false);
  trace_end(a.trace);
}

// -------------------------------------------------------------------------------------------
// Data-parallel gradient exchange over NVLink peer memory (one process per GPU, buffers mapped with
// CUDA IPC): reduce-scatter + all-gather in ONE kernel, no NCCL on the path.
//   A. every rank announces "my local gradient is complete" by storing the epoch into each peer's flag
//   B. rank r owns slice r: it sums that slice over all peers' input buffers in fixed rank order
//      (remote loads, volatile: peer data must not be served from a stale L1 line) and stores the result
//      into every peer's output buffer (remote stores) -> all replicas receive bit-identical values
//   C. the last block of every rank signals "my slice is delivered" to all peers and waits for theirs;
//      when the kernel retires, the whole reduced gradient is in this rank's output buffer.
// Epochs are device-resident and advance once per call, so the kernel is CUDA-graph replayable.  Every
// spin has a timeout (err flag) so that a missing peer cannot wedge the GPU.
struct P2PArgs {
  long long *trace;              // DQNB_TRACE timeline slot (nullable)
  const P2PTable *tab;
  int world, rank, net;
  long long in_off, out_off;      // float offsets of this net's input / output buffer inside an exchange allocation
  long long flag_off;             // float offset of the flag area: uint32 flagA[2][8], flagB[2][8], float norm[2][8]
  float *block_ss;                // [2][gridDim] local scratch: per-block sum of squares of this rank's slice
  long long count;                // floats to reduce (multiple of 4)
  unsigned int *epoch;            // [2] per net, local memory
  unsigned int *ticket;           // [2] per net, local memory
  int *err;                       // host-mapped, sticky: written on a timeout, read by the host only
  int *err_dev;                   // device-resident twin: what kernels read
  unsigned long long timeout_ns;
};
__device__ __forceinline__ bool spin_until_ge(const unsigned int *flag, unsigned int target, int *err, int *err_dev, unsigned long long timeout_ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  while (ld_acquire_sys_u32(flag) < target) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    // a peer is gone (or later than DQNB_P2P_TIMEOUT_MS): give up instead of hanging the GPU.  The flag is sticky and
    // host-visible: the optimiser kernels skip their update once it is set and dqnb_update / dqnb_results fail.
    if (t1 - t0 > timeout_ns) {
      *reinterpret_cast<volatile int *>(err_dev) = 1;
      *reinterpret_cast<volatile int *>(err) = 1;
      __threadfence_system();
      return false;
    }
    __nanosleep(64);
  }
  return true;
}
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(const P2PArgs a) {
  DQNB_PDL_PROLOGUE();
  trace_begin(a.trace);
  __shared__ int s_last;
  const int W = a.world;
  // a peer already missed its timeout (sticky, set before this launch): the replicas have stopped updating
  // (adam_kernel skips), so later exchanges return at once instead of waiting out the timeout again
  if (*reinterpret_cast<const volatile int *>(a.err_dev) != 0) { trace_end(a.trace); return; }
  const unsigned int e = a.epoch[a.net] + 1u;
  float *mine = a.tab->base[a.rank];
  unsigned int *my_flags = reinterpret_cast<unsigned int *>(mine + a.flag_off);
  unsigned int *my_flagA = my_flags + a.net * kMaxPeers, *my_flagB = my_flags + 2 * kMaxPeers + a.net * kMaxPeers;
  // A: flag A was released by the last block of the gradient reduction in front of this kernel
  if (threadIdx.x < W) spin_until_ge(my_flagA + threadIdx.x, e, a.err, a.err_dev, a.timeout_ns);
  __syncthreads();
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[3] = (long long)gtime_ns();
  // B
  const long long n4 = a.count / 4, per = (n4 + W - 1) / W;
  const long long lo = (long long)a.rank * per, hi = min(n4, lo + per);
  float ss = 0.f;                    // sum of squares of the reduced slice (ClipGradients numerator)
  const long long n4_real = (a.count - 4) / 4;   // the 4-float tail carries loss / avg-q, not a gradient
  // The exchange is latency-bound (a few hundred KB per rank over NVLink): each thread first issues every
  // remote load of a batch of kU float4 elements (kU x world loads in flight), then sums in rank order.
  constexpr int kU = 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += kU * stride) {
    float4 v[kU][kMaxPeers];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = i0 + u * stride;
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < W && i < hi) v[u][p] = ld_volatile_f4(a.tab->base[p] + a.in_off + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = i0 + u * stride;
      if (i >= hi) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < W) { acc.x += v[u][p].x; acc.y += v[u][p].y; acc.z += v[u][p].z; acc.w += v[u][p].w; }
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < W) *reinterpret_cast<float4 *>(a.tab->base[p] + a.out_off + 4 * i) = acc;
      if (i < n4_real) ss += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
    }
  }
  {
    __shared__ float red[16];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
      a.block_ss[a.net * gridDim.x + blockIdx.x] = s;
    }
  }
  // C
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[4] = (long long)gtime_ns();
  __syncthreads();
  if (threadIdx.x == 0) {
    // one system fence per block, after the barrier (cumulative over the block's remote stores): they are acknowledged
    // - one NVLink round trip behind the data - before the ticket
    __threadfence_system();
  }
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[5] = (long long)gtime_ns();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(a.ticket + a.net, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (a.trace && threadIdx.x == 0) a.trace[8] = (long long)gtime_ns();
  {
    // this rank's share of ||g||^2 travels with the completion flag.  The per-block partials are summed by a fixed
    // tree (thread b holds block b, b + 512, ..; warp shuffles; warps in order): the same bits on every run.  (A serial
    // loop of volatile loads in the signalling threads cost 6.6 us here.)
    __shared__ float red2[16];
    float s = 0.f;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += reinterpret_cast<volatile float *>(a.block_ss)[a.net * gridDim.x + b];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red2[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red2[w];
    if (threadIdx.x < W) {
      float *pn = a.tab->base[threadIdx.x] + a.flag_off + 4 * kMaxPeers + a.net * kMaxPeers;
      reinterpret_cast<volatile float *>(pn)[a.rank] = s;
      // st.release orders this thread's norm store (and, through the ticket and the fences above, every block's slice
      // stores) before the flag: no second system fence
      unsigned int *pf = reinterpret_cast<unsigned int *>(a.tab->base[threadIdx.x] + a.flag_off) + 2 * kMaxPeers + a.net * kMaxPeers;
      st_release_sys_u32(pf + a.rank, e);
      if (a.trace && threadIdx.x == 0) a.trace[9] = (long long)gtime_ns();
      spin_until_ge(my_flagB + threadIdx.x, e, a.err, a.err_dev, a.timeout_ns);
    }
  }
  __syncthreads();
  if (a.trace && threadIdx.x == 0) a.trace[10] = (long long)gtime_ns();
  if (threadIdx.x == 0) { a.ticket[a.net] = 0; a.epoch[a.net] = e; __threadfence(); }
}

// -------------------------------------------------------------------------------------------
// Act path: SelectActionGreedily on a handful of states (dqn.cpp:734-766; one call per environment step per agent,
// dqn_main.cpp:123).  Latency-bound, so it never touches the learner's stream: it runs as its own captured graph on a
// high-priority stream and reads a SNAPSHOT of the actor (plain fp32, internal padded layout) that the actor's
// adam_kernel writes beside the learner's split planes - into the buffer the act path is not reading - and publishes
// by flipping *cur when the update is complete.  Rows travel through host-mapped pinned memory in both directions and
// completion is a sequence number the host spins on: no stream synchronisation, no cudaMemcpy.
// Skinny M (1..kActMaxRows rows): one warp per output feature, lanes along K, 8 rows of accumulators per pass, the
// weights (3 MB, L2-resident) are read once per 8 rows.  FP32 FFMA: 0.75 MFLOP per row is far below any roofline.
constexpr int kActMaxRows = 64;
constexpr int kActRowGroup = 8;
constexpr int kActHdr = 4;         // floats of header in front of the staged rows: {n, seq, -, -} as 32-bit integers
struct ActCtl {                    // host-mapped pinned completion block (written by the device with posted PCIe writes)
  volatile unsigned int done;      // seq of the last finished call
  unsigned int cur_used;           // snapshot buffer that call read (diagnostics)
};
// The call's header and rows arrive in device memory by one DMA (a memcpy node at the head of the act graph): kernels
// never READ host memory - a sysmem read costs ~1 us and they serialise (measured: 128 blocks polling a mapped word
// made the first layer 0.9 ms long).
struct ActStageArgs {
  const unsigned int *cur; unsigned int *sel;   // latches the snapshot choice of this call for all its layers
};
__global__ void act_stage_kernel(const ActStageArgs a) {
  DQNB_PDL_PROLOGUE();
  if (threadIdx.x == 0) *a.sel = *reinterpret_cast<const volatile unsigned int *>(a.cur) & 1u;
}
struct ActLayerArgs {
  ActCtl *ctl; const int *hdr;             // hdr: device copy of {n, seq}
  const float *snap; long long snap_stride; const unsigned int *sel;   // weights: snap + sel * stride
  long long w_off, b_off; int Kp, N;       // W [N x Kp] row-major, b [N]
  const float *X; int ldx; float *Y; int ldy;
  int lrelu;
  int last; float *h_out;                  // last layer: rows go to mapped host memory [n][16], then ctl->done = seq
};
__global__ void __launch_bounds__(512) act_layer_kernel(const ActLayerArgs a) {
  DQNB_PDL_PROLOGUE();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * (int)(blockDim.x >> 5) + warp;
  const int n = min(max(a.hdr[0], 0), kActMaxRows);
  const float *P = a.snap + (long long)(*a.sel) * a.snap_stride;
  if (j < a.N) {
    const float *w = P + a.w_off + (long long)j * a.Kp;
    const float bias = P[a.b_off + j];
    for (int r0 = 0; r0 < n; r0 += kActRowGroup) {
      float acc[kActRowGroup];
#pragma unroll
      for (int r = 0; r < kActRowGroup; ++r) acc[r] = 0.f;
      for (int k = lane * 4; k < a.Kp; k += 128) {
        const float4 w4 = ld4(w + k);
#pragma unroll
        for (int r = 0; r < kActRowGroup; ++r)
          if (r0 + r < n) acc[r] = dot4(*reinterpret_cast<const float4 *>(a.X + (long long)(r0 + r) * a.ldx + k), w4, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < kActRowGroup; ++r) acc[r] = warp_sum(acc[r]);
      if (lane < kActRowGroup && r0 + lane < n) {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < kActRowGroup; ++r)
          if (r == lane) v = acc[r];
        v += bias;
        if (a.lrelu) v = fmaxf(v, 0.f) + kNegSlope * fminf(v, 0.f);
        if (a.last) a.h_out[(long long)(r0 + lane) * 16 + j] = v;
        else a.Y[(long long)(r0 + lane) * a.ldy + j] = v;
      }
    }
  }
  if (a.last) {          // launched as ONE block of 512 threads (N <= 16, checked by the host): publish completion
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      a.ctl->cur_used = *a.sel;
      __threadfence_system();
      a.ctl->done = (unsigned int)a.hdr[1];
    }
  }
}

// split an fp32 array into (hi, lo) planes / join it back (parameter import / export)
__global__ void split_kernel(const float *x, float *hi, float *lo, long long n) {
  DQNB_PDL_PROLOGUE();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float h = tf32_hi(x[i]);
  hi[i] = h; lo[i] = x[i] - h;
}
__global__ void join_kernel(const float *hi, const float *lo, float *x, long long n) {
  DQNB_PDL_PROLOGUE();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = hi[i] + lo[i];
}
// standalone gemm test: sum split planes
__global__ void sum_planes_kernel(const float *part, long long stride, int planes, float *out, long long n) {
  DQNB_PDL_PROLOGUE();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < planes; ++p) s += part[(long long)p * stride + i];
  out[i] = s;
}

}  // namespace dqnb
