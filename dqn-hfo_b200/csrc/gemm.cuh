// gemm.cuh — dense contraction kernels for the actor/critic towers (dqn.cpp:400-454) on sm_100a.
//
// Every GEMM operand lives in HBM as two fp32 planes [2][rows][ld] ("hi" = value rounded to TF32,
// "lo" = exact remainder), see common.cuh.  One launch computes
//      D[M x N] = sum_k A(m,k) * B(n,k)          (A, B given K-major or MN-major)
// as hi*hi + hi*lo + lo*hi (3xTF32 ~ fp32 accuracy) and applies a fused epilogue:
//   EPI_FWD   : InnerProduct forward + bias + leaky ReLU (Caffe InnerProduct/ReLU, dqn.cpp:340-356,
//               :292-301) -> split planes of the next layer's input
//   EPI_DX    : InnerProduct backward w.r.t. bottom + in-place leaky-ReLU backward of the layer
//               below (mask from that layer's saved activation) -> split planes of dZ
//   EPI_PLAIN : raw fp32 (weight-gradient partials of the split-K dW GEMMs, input diffs)
//
// gemm_tc_kernel   : TMA (cp.async.bulk.tensor.3d, 128B swizzle) -> 2-4-stage smem ring ->
//                    tcgen05.mma.cta_group::1.kind::tf32, two instructions per K = 8 step (A_hi x [B_hi | B_lo] with
//                    N = 2 BN, A_lo x B_hi with N = BN), three fp32 accumulators in TMEM, tiles M128 x N{32,64,128}
//                    -> tcgen05.ld epilogue -> swizzled staging -> TMA store.  Warp roles: 0 = TMA producer,
//                    1 = MMA issuer (+ TMEM alloc), 4-7 = A movers when the A operand is staged in tensor memory
//                    (tcgen05.st); all 8 warps run the epilogue (two per TMEM lane quarter).  The epilogue flavour
//                    is a template parameter.
// gemm_simt_kernel : plain FFMA tiles over the same operands/epilogues (verification mode).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace dqnb {

enum { EPI_FWD = 0, EPI_DX = 1, EPI_PLAIN = 2 };

struct GemmParams {
  int M, N, K;            // padded problem: M rows of D, N cols of D, K contraction
  int a_mn, b_mn;         // 0: K-major (contraction contiguous)  1: MN-major
  int splits;             // split-K factor (gridDim.z)
  int epi;
  // raw operand views (SIMT mode and descriptors): plane 0 = hi, plane 1 = lo
  const float *A; long long a_plane; int lda;
  const float *B; long long b_plane; int ldb;
  // outputs
  float *out_hi; float *out_lo; int ldo;     // EPI_FWD / EPI_DX
  float *out; long long out_split_stride;     // EPI_PLAIN: out + split*stride
  const float *bias_hi; const float *bias_lo; // EPI_FWD (nullable)
  const float *mask_hi; const float *mask_lo; int ldmask;  // EPI_DX, verification kernel: saved activation
  // tcgen05 kernel: the ReLU sign of every activation travels as ONE BIT (word (m, n/32), bit n%32 = y > 0).
  // EPI_FWD writes it next to the activation, EPI_DX reads 4 bytes per 32 columns instead of 256
  // (the fp32 mask tile was 64 KB of LSU loads per CTA competing with the TMA operand stream).
  uint32_t *relu_bits_out; const uint32_t *relu_bits_in; int ldbits;
  // tcgen05 kernel, EPI_DX: the bias gradient of the layer below is the column sum of the dZ tile this epilogue
  // produces (Caffe: gemv(dY^T, ones)).  Each CTA adds its 128 rows and writes one partial row:
  // colsum_out[m_tile * colsum_stride + n]; reduce_kernel sums the m_tile planes in fixed order (nullable).
  float *colsum_out; long long colsum_stride;
  int apply_lrelu;                            // EPI_FWD: 0 for linear layers
  int bn;                                     // N tile of the tcgen05 kernel: 32, 64 or 128
  int stages;                                 // depth of the TMA->MMA smem ring (2..4).  BN=64 with 2 stages is 98 KB
                                              // per CTA: two CTAs share an SM, one's epilogue under the other's mainloop
  int cluster_k;                              // 1: launched as 2-CTA clusters along z; the CTAs split K and the
                                              //    leader adds its peer's partial tile over DSMEM before the epilogue
  int pdl_early;                              // 1: release the dependent grid right after griddepcontrol.wait (its CTAs
                                              //    then sit on SMs - 197 KB of smem each - for this kernel's whole
                                              //    duration); 0: when the last MMA has been issued, so they only overlap
                                              //    the accumulator drain + epilogue (measured: profiles/r01c_*)
  int a_ts;                                   // 1 (K-major A only): the A tile is copied shared -> tensor memory by four mover
                                              // warps and tcgen05.mma reads it from there (see gemm_tc_body)
  int dbg;                                    // perf experiments (gemm_test only): bit0 skip MMA, bit1 skip TMA
  long long *dbg_clk;                         // optional: MMA-thread clock64 stamps {start, issued, complete}
};

struct alignas(64) GemmArgs {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmD;      // output: {ldo, M, 2 planes} (EPI_FWD / EPI_DX) or {ldo, M, splits} (EPI_PLAIN), box {32, 32, 2 | 1}
  GemmParams p;
};

// ---------------------------------------------------------------------------------------------
// shared epilogue: CNT consecutive columns [n0, n0+CNT) of row m
// ---------------------------------------------------------------------------------------------
template <int CNT>
__device__ __forceinline__ void epi_store(const GemmParams &p, int m, int n0, const float *acc,
                                          int split) {
  if (m >= p.M || n0 >= p.N) return;
  if (p.epi == EPI_PLAIN) {
    float *o = p.out + (long long)split * p.out_split_stride + (long long)m * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < CNT; j += 4)
      *reinterpret_cast<float4 *>(o + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    return;
  }
  float hi[CNT], lo[CNT];
  if (p.epi == EPI_FWD) {
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
      float v = acc[j];
      if (p.bias_hi) v += p.bias_hi[n0 + j] + p.bias_lo[n0 + j];
      if (p.apply_lrelu) v = fmaxf(v, 0.f) + kNegSlope * fminf(v, 0.f);
      hi[j] = tf32_hi(v);
      lo[j] = v - hi[j];
    }
  } else {  // EPI_DX
    const float *mh = p.mask_hi + (long long)m * p.ldmask + n0;
    const float *ml = p.mask_lo + (long long)m * p.ldmask + n0;
#pragma unroll
    for (int j = 0; j < CNT; j += 4) {
      const float4 a = *reinterpret_cast<const float4 *>(mh + j);
      const float4 b = *reinterpret_cast<const float4 *>(ml + j);
      const float y[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        // Caffe ReLU backward: diff * ((y > 0) + slope * (y <= 0)), y = in-place activation
        const float v = acc[j + t] * (y[t] > 0.f ? 1.f : kNegSlope);
        hi[j + t] = tf32_hi(v);
        lo[j + t] = v - hi[j + t];
      }
    }
  }
  float *oh = p.out_hi + (long long)m * p.ldo + n0;
  float *ol = p.out_lo + (long long)m * p.ldo + n0;
#pragma unroll
  for (int j = 0; j < CNT; j += 4) {
    *reinterpret_cast<float4 *>(oh + j) = make_float4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
    *reinterpret_cast<float4 *>(ol + j) = make_float4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
  }
}

// ---------------------------------------------------------------------------------------------
// verification-mode kernel: 64x64 tile, 256 threads, 4x4 per thread, generic operand strides
// ---------------------------------------------------------------------------------------------
constexpr int ST = 64, SK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
  DQNB_PDL_PROLOGUE();
  __shared__ float As[SK][ST + 4];
  __shared__ float Bs[SK][ST + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST, split = blockIdx.z;
  const int kchunk = (p.K / SK + p.splits - 1) / p.splits * SK;
  const int kb = split * kchunk, ke = min(p.K, kb + kchunk);
  // element (r, k) of A lives at r*a_rs + k*a_cs
  const long long a_rs = p.a_mn ? 1 : p.lda, a_cs = p.a_mn ? p.lda : 1;
  const long long b_rs = p.b_mn ? 1 : p.ldb, b_cs = p.b_mn ? p.ldb : 1;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int k0 = kb; k0 < ke; k0 += SK) {
    for (int i = tid; i < ST * SK; i += 256) {
      // thread -> (row, k) chosen so that the contiguous axis is the fast one
      int r, k;
      if (p.a_mn) { r = i % ST; k = i / ST; } else { k = i % SK; r = i / SK; }
      float v = 0.f;
      if (m0 + r < p.M && k0 + k < ke) {
        const long long o = (long long)(m0 + r) * a_rs + (long long)(k0 + k) * a_cs;
        v = p.A[o] + p.A[o + p.a_plane];
      }
      As[k][r] = v;
      if (p.b_mn) { r = i % ST; k = i / ST; } else { k = i % SK; r = i / SK; }
      v = 0.f;
      if (n0 + r < p.N && k0 + k < ke) {
        const long long o = (long long)(n0 + r) * b_rs + (long long)(k0 + k) * b_cs;
        v = p.B[o] + p.B[o + p.b_plane];
      }
      Bs[k][r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) epi_store<4>(p, m0 + ty * 4 + i, n0 + tx * 4, acc[i], split);
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMA / TMEM kernel
// ---------------------------------------------------------------------------------------------
constexpr int BM = 128, BK = 32;
constexpr int TC_THREADS = 256;   // warp 0: TMA producer, warp 1: MMA issuer (+ TMEM alloc); all 8 run the epilogue
constexpr int TC_SMEM_MAX = 232448;                   // 227 KB: the per-CTA maximum on sm_100

template <int BN_>
struct TcCfg {
  static constexpr int BN = BN_;
  static constexpr int A_PLANE_BYTES = BM * BK * 4;            // 16 KB: one plane of the A tile
  static constexpr int B_PLANE_BYTES = BN_ * BK * 4;           // 8 / 16 KB
  static constexpr int A_STAGE_BYTES = 2 * A_PLANE_BYTES;      // hi + lo
  static constexpr int B_STAGE_BYTES = 2 * B_PLANE_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;   // 48 / 64 KB
  static constexpr int MAX_STAGES = (TC_SMEM_MAX - 2048) / STAGE_BYTES > 4 ? 4 : (TC_SMEM_MAX - 2048) / STAGE_BYTES;
  static constexpr int MIN_STAGES = 2;          // also holds the epilogue's staging boxes (4 warps x BN/32 x 8 KB)
  static constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 512 /*bias tile*/ + 4 * BN_ * 4 /*column sums*/; }
  // A-in-TMEM mode: ring of A tiles behind the accumulators, [stage][hi: 32 columns | lo: 32 columns] (one k-block)
  static constexpr int A_TMEM_COL0 = 3 * BN_;
  static constexpr int A_TMEM_STAGES = (512 - 3 * BN_) / 64 >= 4 ? 4 : (512 - 3 * BN_) / 64;
  static constexpr int TMEM_USED = 3 * BN_;                    // three fp32 accumulators of BN columns: hi*hi, hi*lo, lo*hi
  static constexpr int TMEM_COLS = TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
};
constexpr int MN_GROUP_BYTES = 2 * 32 * BK * 4;       // MN-major: one 32-wide group, hi plane then lo plane

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
// one lane of a converged warp (deterministic for a given member mask)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred px;\n\t"
      "elect.sync _|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1,
                                            int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// TMA store of one staged box (128B-swizzled in shared memory) + bulk-group bookkeeping of the issuing thread
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// .read: wait until the staged boxes have been READ out of shared memory (it may then be released); the writes
// themselves complete before the grid does, which is what griddepcontrol.wait / stream order of the consumer wait for
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_smem_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand read from tensor memory (lane = row, 8 columns of 32-bit per K = 8 step), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> tensor memory: thread t of the warp writes 32 consecutive columns of lane (taddr.lane + t)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ld_smem_u4(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}

// UMMA shared-memory matrix descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46)
// | version=1 [46,48) | layout_type [61,64).
//  K-major tile  : rows of 128 B (32 fp32 of K), SWIZZLE_128B (type 2, 16-byte chunks XOR row%8),
//                  8-row atoms 1024 B apart (SBO); LBO unused.  k-step advance: +32 B.
//  MN-major tile : 128 B rows hold 32 consecutive M/N elements of one k.  32-bit MN-major operands
//                  must use the 128B swizzle with 32-byte atomicity (type 1; TMA
//                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of 4 k-rows, SBO = 512 B to the next
//                  k-group, LBO = distance between 32-wide M/N groups.  k-step (8 rows): +1024 B.
template <int MN>
struct DescHi {   // the constant upper 32 bits and the LBO part of the lower 32 bits
  static constexpr uint32_t hi = MN ? ((512u >> 4) | (1u << 14) | (1u << 29))      // SBO 512, v1, type 1
                                    : ((1024u >> 4) | (1u << 14) | (2u << 29));    // SBO 1024, v1, type 2
  static constexpr uint32_t lbo = MN ? ((uint32_t)(MN_GROUP_BYTES >> 4) << 16) : (1u << 16);
  static constexpr uint32_t kstep = MN ? (1024u >> 4) : (32u >> 4);                // per 8 of K
};
template <int MN>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return ((uint64_t)DescHi<MN>::hi << 32) | (uint64_t)(((saddr & 0x3FFFFu) >> 4) | DescHi<MN>::lbo);
}
// The B tile read as ONE operand of 2 BN columns, [B_hi | B_lo]:
//  K-major  : the lo plane's rows follow the hi plane's rows (plane = BN rows of 128 B) - same descriptor, N doubled;
//             accumulator columns [0, BN) = A*B_hi, [BN, 2 BN) = A*B_lo.
//  MN-major : every 32-wide group holds its hi plane (4 KB) then its lo plane (4 KB); with LBO = 4 KB the operand's
//             32-column groups are hi0, lo0, hi1, lo1, ..: accumulator columns 64 g + [0, 32) = A*B_hi of group g,
//             64 g + [32, 64) = A*B_lo of group g.
template <int MN>
__device__ __forceinline__ uint64_t make_desc_bcat(uint32_t saddr) {
  constexpr uint32_t lbo = MN ? ((4096u >> 4) << 16) : (1u << 16);
  return ((uint64_t)DescHi<MN>::hi << 32) | (uint64_t)(((saddr & 0x3FFFFu) >> 4) | lbo);
}
// TMEM column of the 32-column chunk ci of accumulator hh (hi*hi), hl (A_hi*B_lo), lh (A_lo*B_hi)
template <int B_MN, int BN_> __device__ __forceinline__ uint32_t acc_col_hh(int ci) { return B_MN ? 64u * ci : 32u * ci; }
template <int B_MN, int BN_> __device__ __forceinline__ uint32_t acc_col_hl(int ci) { return B_MN ? 64u * ci + 32u : (uint32_t)BN_ + 32u * ci; }
template <int B_MN, int BN_> __device__ __forceinline__ uint32_t acc_col_lh(int ci) { return 2u * BN_ + 32u * ci; }

// instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, a_major [15], b_major [16],
// N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int a_mn, int b_mn, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// timeline stamps of CTA (0,0,0) for gemm_test / DQNB_TRACE: [0] entry [2] pdl_wait passed [3] first operands
// landed [4] accumulators complete [5] epilogue stores issued; over all CTAs of the grid: [1] latest pass of
// pdl_wait, [6] latest exit, [8] latest first-operands, [9] latest accumulators complete, [10] latest epilogue done
#define DQNB_STAMP(i) do { if (p.dbg_clk && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg_clk[i] = (long long)gtime_ns(); } while (0)
#define DQNB_STAMP_MAX(i) do { if (p.dbg_clk) atomicMax(p.dbg_clk + (i), (long long)gtime_ns()); } while (0)

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_saddr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}

// The warp roles run with the whole warp converged: every value feeding TMA / MMA issue is
// warp-uniform (uniform registers, no per-thread descriptor arithmetic), and one elected lane
// issues.  A single divergent thread doing that arithmetic costs ~300 cycles per k-step, 3x the
// tensor time of the three MMAs it feeds (measured: profiles/r01_gemm_issue_loop.md).
template <int A_MN, int B_MN, int BN_, int EPI_>
__device__ __forceinline__ void gemm_tc_body(const GemmArgs &args, const int n_tile, const int m_tile, const int split) {
  using Cfg = TcCfg<BN_>;
  extern __shared__ uint8_t smem_raw[];
  const GemmParams &p = args.p;
  const int STAGES = p.stages;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // swizzle atoms need 1024 B alignment
  uint8_t *base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * Cfg::STAGE_BYTES;   // full[STAGES], empty[STAGES], tmem_full
  volatile uint32_t *tmem_slot =
      reinterpret_cast<volatile uint32_t *>(base_ptr + STAGES * Cfg::STAGE_BYTES + 192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = p.K / BK;
  const int per = (kblocks + p.splits - 1) / p.splits;
  const int kb0 = split * per;
  const int kb1 = min(kblocks, kb0 + per);
  const int iters = max(kb1 - kb0, 0);
  const uint32_t tfull = bars + 8 * (2 * STAGES);
  // A-in-TMEM mode (K-major A): a_full[AT] "tile copied into tensor memory" (one arrival per mover warp),
  // a_empty[AT] "the MMAs that read it have retired"
  constexpr int AT = Cfg::A_TMEM_STAGES;
  const bool a_ts = !A_MN && p.a_ts != 0;
  const uint32_t afull0 = bars + 8 * (2 * STAGES + 1), aempty0 = afull0 + 8 * AT;
  const bool clus = p.cluster_k != 0;          // split-K over a 2-CTA cluster (blockIdx.z = cluster rank)
  uint32_t crank = 0;
  if (clus) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (threadIdx.x == 0) DQNB_STAMP(0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&args.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&args.tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&args.tmD) : "memory");
    for (int i = 0; i < 2 * STAGES + 1; ++i) mbar_init(bars + 8 * i, 1);
    for (int i = 0; i < AT; ++i) { mbar_init(afull0 + 8 * i, 4); mbar_init(aempty0 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void *)tmem_slot)),
                 "r"(a_ts ? 512u : (uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the tail of the
  // previous kernel; from here on we read what it wrote.
  pdl_wait();
  if (p.pdl_early) pdl_launch_dependents();
  if (threadIdx.x == 0) DQNB_STAMP(2);
  if (p.dbg_clk && threadIdx.x == 0) atomicMax(p.dbg_clk + 1, (long long)gtime_ns());   // latest CTA to pass the wait

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t full = bars + 8 * s, empty = bars + 8 * (STAGES + s);
      mbar_wait(empty, ph ^ 1u);
      if (elect_one()) {
        if (p.dbg & 2) {   // experiment: no loads, just hand the (stale) stage to the MMA warp
          mbar_arrive(full);
        } else {
          mbar_expect_tx(full, Cfg::STAGE_BYTES);
          const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_STAGE_BYTES;
          const int k0 = (kb0 + it) * BK;
          if (!A_MN) {
            tma_load_3d(sa, &args.tmA, k0, m_tile * BM, 0, full);              // box {32, 128, 2}
          } else {
#pragma unroll
            for (int g = 0; g < BM / 32; ++g)                                   // box {32, 32, 2}
              tma_load_3d(sa + g * MN_GROUP_BYTES, &args.tmA, m_tile * BM + g * 32, k0, 0, full);
          }
          if (!B_MN) {
            tma_load_3d(sb, &args.tmB, k0, n_tile * BN_, 0, full);             // box {32, BN, 2}
          } else {
#pragma unroll
            for (int g = 0; g < BN_ / 32; ++g)
              tma_load_3d(sb + g * MN_GROUP_BYTES, &args.tmB, n_tile * BN_ + g * 32, k0, 0, full);
          }
        }
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = umma_idesc_tf32(A_MN, B_MN, BN_);
    constexpr uint32_t idesc_w = umma_idesc_tf32(A_MN, B_MN, 2 * BN_);
    constexpr uint32_t a_lo_off = A_MN ? 4096u : (uint32_t)Cfg::A_PLANE_BYTES;
    int s = 0, as = 0;
    uint32_t ph = 0, pa = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t full = bars + 8 * s, empty = bars + 8 * (STAGES + s);
      mbar_wait(full, ph);
      if (a_ts) mbar_wait(afull0 + 8 * as, pa);
      tc_fence_after();
      if (it == 0 && lane == 0) { DQNB_STAMP(3); DQNB_STAMP_MAX(8); }
      const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_STAGE_BYTES;
      const uint64_t a_hi = make_desc<A_MN>(sa), a_lo = make_desc<A_MN>(sa + a_lo_off);
      const uint64_t b_hi = make_desc<B_MN>(sb), b_cat = make_desc_bcat<B_MN>(sb);
      if (elect_one()) {
        if (!(p.dbg & 1)) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t ka = (uint64_t)(ks * DescHi<A_MN>::kstep), kb = (uint64_t)(ks * DescHi<B_MN>::kstep);
            // Two instructions per k-step instead of three: a tcgen05.mma costs ~60 cycles + 0.2 per column of N
            // here (measured, operands in shared memory: 67 / 73 / 87 cycles at N = 32 / 64 / 128), so A_hi is
            // multiplied with B_hi and B_lo in ONE instruction of N = 2 BN - the two planes of the B tile are
            // neighbours in shared memory and read as one operand - and A_lo with B_hi in a second one.
            // Three TMEM accumulators: the dominant hi*hi chain and the two small cross terms.  The tensor core
            // truncates when it adds into the fp32 accumulator, so keeping the 2^-11-scaled terms out of the long
            // chain cuts the accumulated rounding bias ~3x; the epilogue adds them.
            const uint32_t first = (it > 0 || ks > 0) ? 1u : 0u;
            if (a_ts) {
              // A from tensor memory: no 4 KB shared-memory read of the A slice per instruction (measured: 74 / 53
              // cycles at N = 128 / 64 instead of 87 / 73)
              const uint32_t a_t = tmem + (uint32_t)Cfg::A_TMEM_COL0 + (uint32_t)as * 64u + (uint32_t)ks * 8u;
              tc_mma_tf32_ts(tmem, a_t, b_cat + kb, idesc_w, first);
              tc_mma_tf32_ts(tmem + 2 * BN_, a_t + 32u, b_hi + kb, idesc, first);
            } else {
              tc_mma_tf32(tmem, a_hi + ka, b_cat + kb, idesc_w, first);
              tc_mma_tf32(tmem + 2 * BN_, a_lo + ka, b_hi + kb, idesc, first);
            }
          }
        }
        tc_commit(empty);                  // frees the smem stage when these MMAs retire
        if (a_ts) tc_commit(aempty0 + 8 * as);   // ... and the A tile in tensor memory
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
      if (++as == AT) { as = 0; pa ^= 1u; }
    }
    if (elect_one()) {
      tc_commit(tfull);                    // accumulators complete -> epilogue
    }
    __syncwarp();
    if (!p.pdl_early) pdl_launch_dependents();   // every MMA is issued: the next kernel's prologue overlaps our tail
  } else if (!A_MN && warp >= 4 && a_ts) {
    // ---------------- A movers (A-in-TMEM mode): shared memory -> registers -> tensor memory ----------------
    // Warp 4 + q owns TMEM lane quarter q = rows 32 q .. 32 q + 31 of the tile; a thread copies its row of the
    // k-block: 128 B of the hi plane and 128 B of the lo plane (16-byte chunk j of row r sits at j ^ (r & 7):
    // conflict-free LDS.128) into 32 + 32 columns of the stage's slot.
    const int q = warp & 3;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const uint32_t row_off = row * 128u, sw = row & 7u;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)Cfg::A_TMEM_COL0;
    int s = 0, as = 0;
    uint32_t ph = 0, pa = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(aempty0 + 8 * as, pa ^ 1u);
      mbar_wait(bars + 8 * s, ph);
      const uint32_t sa = base + s * Cfg::STAGE_BYTES + row_off;
      uint32_t rh[32], rl[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t o = ((uint32_t)j ^ sw) << 4;
        ld_smem_u4(sa + o, rh[4 * j], rh[4 * j + 1], rh[4 * j + 2], rh[4 * j + 3]);
        ld_smem_u4(sa + (uint32_t)Cfg::A_PLANE_BYTES + o, rl[4 * j], rl[4 * j + 1], rl[4 * j + 2], rl[4 * j + 3]);
      }
      tc_fence_after();                    // the slot's previous readers (MMAs) retired: ordered before our writes
      tmem_st32(t_lane + (uint32_t)as * 64u, rh);
      tmem_st32(t_lane + (uint32_t)as * 64u + 32u, rl);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(afull0 + 8 * as);
      if (++s == STAGES) { s = 0; ph ^= 1u; }
      if (++as == AT) { as = 0; pa ^= 1u; }
    }
  }
  {
    // ---------------- epilogue: TMEM -> registers -> swizzled smem -> TMA store ----------------
    // All eight warps take part (the producer and the MMA issuer join when their loops are done): warp w may
    // read TMEM lane quarter w % 4, so warps w and w + 4 share a quarter and split its 32-column chunks.
    // TMEM hands every thread one accumulator ROW.  All TMA loads have been consumed once tfull fires, so
    // the pipeline stages are reused as staging: per chunk a warp writes its 32 rows x 128 B (x 2 planes) in
    // the 128B-swizzle pattern (16-byte slot j of row r at slot j ^ (r & 7): conflict-free float4 stores)
    // and one lane hands the box to the TMA unit, which writes full lines to L2 and clips the rows / columns
    // outside the matrix.  (Row-per-thread global stores: 4.2 us per tile; staged and re-read by the threads
    // for coalesced stores: 2.5 us.)
    constexpr int LDS = BN_ + 4;             // row stride of the split-K peer's raw partial tile
    constexpr int NCH = BN_ / 32;            // 32-column chunks of the tile
    constexpr int MYCH = (NCH + 1) / 2;      // ... handled by this warp: chunks half, half + 2, .. (those below NCH)
    static_assert(BN_ == 32 || BN_ == 64 || BN_ == 128, "epilogue mapping assumes BN of 32, 64 or 128");
    // BN = 32 (layers with few output tiles: twice the CTAs, half the bytes each SM has to store): one chunk per TMEM
    // lane quarter, taken by warps 0-3; warps 4-7 only help with the bias tile
    static_assert(BM * LDS * 4 <= Cfg::MIN_STAGES * Cfg::STAGE_BYTES, "peer partial tile must fit in the stage ring");
    static_assert(4 * NCH * 2 * 4096 <= Cfg::MIN_STAGES * Cfg::STAGE_BYTES, "staging boxes must fit in the stage ring");
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int half = warp >> 2;
    float *s_bias = reinterpret_cast<float *>(base_ptr + STAGES * Cfg::STAGE_BYTES + 256);
    float *s_colsum = s_bias + 128;          // [4 row quarters][BN]
    float *st_hi = reinterpret_cast<float *>(base_ptr) + (q * 32) * LDS;
    const int n_base = n_tile * BN_;
    const bool peer = clus && crank == 1;
    if (warp >= 2 && !peer && EPI_ == EPI_FWD) {    // stage the tile's bias once (overlaps the mainloop)
      for (int t = threadIdx.x - 64; t < BN_; t += TC_THREADS - 64) {
        const int n = n_base + t;
        s_bias[t] = (p.bias_hi && n < p.N) ? p.bias_hi[n] + p.bias_lo[n] : 0.f;
      }
    }
    // EPI_DX: the ReLU' sign bits of this thread's row (one word per 32 columns) do not depend on the
    // MMAs: loaded before waiting for the accumulators.
    const int m_row = m_tile * BM + q * 32 + lane;
    uint32_t rbits[MYCH];
#pragma unroll
    for (int i = 0; i < MYCH; ++i) {
      const int c0 = (half + 2 * i) * 32;
      rbits[i] = 0u;
      if (!peer && EPI_ == EPI_DX && half + 2 * i < NCH && m_row < p.M && n_base + c0 < p.N)
        rbits[i] = __ldg(p.relu_bits_in + (long long)m_row * p.ldbits + ((n_base + c0) >> 5));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory");   // bias tile visible; warps 0/1 are done issuing
    if (iters > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    const uint32_t taddr0 = tmem + ((uint32_t)(q * 32) << 16);
    if (peer) {
      // split-K peer: publish the raw partial tile in this CTA's staging area; the leader reads it over
      // distributed shared memory after cluster barrier #1 (below) and runs the real epilogue
      float *mine = st_hi + lane * LDS;
#pragma unroll
      for (int i = 0; i < MYCH; ++i) {
        const int c0 = (half + 2 * i) * 32;
        if (half + 2 * i >= NCH) break;
        uint32_t r[32], r2[32], r3[32];
        if (iters > 0) {
          tmem_ld32(taddr0 + acc_col_hh<B_MN, BN_>(half + 2 * i), r);
          tmem_ld32(taddr0 + acc_col_hl<B_MN, BN_>(half + 2 * i), r2);
          tmem_ld32(taddr0 + acc_col_lh<B_MN, BN_>(half + 2 * i), r3);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) { r[j] = 0u; r2[j] = 0u; r3[j] = 0u; }
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + (__uint_as_float(r2[j]) + __uint_as_float(r3[j]));
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4 *>(mine + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    } else {
      // the TMEM loads of this warp's first chunk go out before the (cluster) hand-shake
      uint32_t ra[32], rb[32], rc[32];
      auto load_chunk = [&](int ci) {
        tmem_ld32(taddr0 + acc_col_hh<B_MN, BN_>(ci), ra);
        tmem_ld32(taddr0 + acc_col_hl<B_MN, BN_>(ci), rb);
        tmem_ld32(taddr0 + acc_col_lh<B_MN, BN_>(ci), rc);
      };
      if (iters > 0 && half < NCH) load_chunk(half);
      if (clus) { cluster_arrive_release(); cluster_wait_acquire(); }   // barrier #1: the peer's partial is visible
      if (warp == 2 && lane == 0) { DQNB_STAMP(4); DQNB_STAMP_MAX(9); }
      const float *peer_row = st_hi + lane * LDS;          // same offset inside the peer CTA's shared memory
      const uint32_t box0 = base + (uint32_t)(q * NCH) * 8192u;   // staging boxes: [quarter][chunk][plane][32 rows][128 B]
      const uint32_t row_off = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
      const int planes_out = EPI_ == EPI_PLAIN ? 1 : 2;
#pragma unroll
      for (int i = 0; i < MYCH; ++i) {
        const int c0 = (half + 2 * i) * 32;
        if (half + 2 * i >= NCH) break;
        float v[32];
        const uint32_t bits_in = rbits[i];
        uint32_t bits_out = 0u;
        if (iters > 0) {
          tmem_wait_ld();
          if (i == 0 && warp == 2 && lane == 0) DQNB_STAMP(11);       // first accumulator chunk in registers
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]) + (__uint_as_float(rb[j]) + __uint_as_float(rc[j]));
          if (i + 1 < MYCH && half + 2 * (i + 1) < NCH) load_chunk(half + 2 * (i + 1));   // next chunk in flight while this one is processed
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (clus) {                            // + the other half of K, from the peer CTA's shared memory
          const uint32_t mine = smem_u32(peer_row + c0);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 pp = ld_dsmem_f4(mine + j * 4, 1u);
            v[j] += pp.x; v[j + 1] += pp.y; v[j + 2] += pp.z; v[j + 3] += pp.w;
          }
        }
        const uint32_t box = box0 + (uint32_t)(c0 >> 5) * 8192u + row_off;
        if (EPI_ == EPI_PLAIN) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            st_smem_f4(box + ((((uint32_t)j >> 2) ^ sw) << 4), v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float o[4], l[4];
            float4 aux = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI_ == EPI_FWD) aux = *reinterpret_cast<const float4 *>(s_bias + c0 + j);
            const float ax[4] = {aux.x, aux.y, aux.z, aux.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float x = v[j + t];
              if (EPI_ == EPI_FWD) {
                x += ax[t];                                            // InnerProduct bias
                if (p.apply_lrelu) x = fmaxf(x, kNegSlope * x);        // == max(x,0) + slope * min(x,0) (Caffe ReLU), one op less
                bits_out |= (x > 0.f ? 1u : 0u) << (j + t);            // sign of the in-place activation
              } else {
                x *= ((bits_in >> (j + t)) & 1u) ? 1.f : kNegSlope;    // ReLU backward on the in-place activation
                v[j + t] = x;                                          // kept for the column sum below
              }
              o[t] = tf32_hi(x);
              l[t] = x - o[t];
            }
            const uint32_t slot = box + ((((uint32_t)j >> 2) ^ sw) << 4);
            st_smem_f4(slot, o[0], o[1], o[2], o[3]);
            st_smem_f4(slot + 4096u, l[0], l[1], l[2], l[3]);
          }
          if (EPI_ == EPI_FWD && p.relu_bits_out && m_row < p.M && n_base + c0 < p.N)
            p.relu_bits_out[(long long)m_row * p.ldbits + ((n_base + c0) >> 5)] = bits_out;
          if (EPI_ == EPI_DX && p.colsum_out) {
            // column sums over this warp's 32 rows: halving butterfly, 31 shuffles; lane L ends up with column c0 + L
            // (bit b of the lane selects bit b of the column).  Fixed tree -> the same bits on every run.
            if (m_row >= p.M) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
              const bool upper = (lane & sft) != 0;
#pragma unroll
              for (int j = 0; j < sft; ++j) {
                const float send = upper ? v[j] : v[j + sft];
                const float keep = upper ? v[j + sft] : v[j];
                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
              }
            }
            s_colsum[q * BN_ + c0 + lane] = v[0];
          }
        }
        // hand this chunk's box (both planes) to the TMA unit; the next chunk is staged meanwhile
        fence_proxy_async_smem();
        __syncwarp();
        if (half + 2 * (i + 1) >= NCH && warp == 2 && lane == 0) DQNB_STAMP(12);  // last box staged
        if (lane == 0 && n_base + c0 < p.N && m_tile * BM + q * 32 < p.M) {
          tma_store_3d(&args.tmD, box - row_off, n_base + c0, m_tile * BM + q * 32, planes_out == 1 ? split : 0);
          bulk_commit();
        }
      }
      if (EPI_ == EPI_DX && p.colsum_out) {
        asm volatile("bar.sync 2, %0;" ::"n"(TC_THREADS) : "memory");   // every warp of this (non-peer) CTA is here
        if (threadIdx.x < BN_ && n_base + (int)threadIdx.x < p.N) {
          const float *c = s_colsum + threadIdx.x;
          p.colsum_out[(long long)m_tile * p.colsum_stride + n_base + threadIdx.x] = ((c[0] + c[BN_]) + c[2 * BN_]) + c[3 * BN_];
        }
      }
      if (lane == 0) bulk_wait_read();         // the boxes have been read before shared memory goes away
      __syncwarp();
      if (warp == 2 && lane == 0) { DQNB_STAMP(5); DQNB_STAMP_MAX(10); }
    }
  }
  if (clus) {
    // every thread of both CTAs passes two cluster barriers: #1 publishes the peer's partial tile (the
    // leader's warps already passed it above), #2 keeps the peer's shared memory alive until the leader
    // has read it
    if (crank != 0) { cluster_arrive_release(); cluster_wait_acquire(); }
    cluster_arrive_release();
    cluster_wait_acquire();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DQNB_STAMP(13);                                                  // CTA (0,0,0) past its last barrier
  if (p.dbg_clk && threadIdx.x == 0) atomicMax(p.dbg_clk + 6, (long long)gtime_ns());   // latest CTA exit of the grid
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a_ts ? 512u : (uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// The epilogue flavour is a template parameter: with a run-time switch the compiler kept a branch per group of four
// columns and could not interleave the groups (12 instructions and ~60 cycles per element at two warps per scheduler;
// 1.0 us of the 2.3 us epilogue).
template <int A_MN, int B_MN, int BN_, int EPI_>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs args) {
  gemm_tc_body<A_MN, B_MN, BN_, EPI_>(args, blockIdx.x, blockIdx.y, blockIdx.z);
}

// host-side dispatch over the template instances: forward layers are K-major x K-major, dX is K-major x MN-major,
// raw outputs (weight gradients, input diffs, the unit test) exist for every operand layout
typedef void (*TcKernel)(const GemmArgs);
template <int BN_>
inline TcKernel tc_kernel_for_bn(int a_mn, int b_mn, int epi) {
  if (epi == EPI_FWD) return (!a_mn && !b_mn) ? gemm_tc_kernel<0, 0, BN_, EPI_FWD> : nullptr;
  if (epi == EPI_DX) return (!a_mn && b_mn) ? gemm_tc_kernel<0, 1, BN_, EPI_DX> : nullptr;
  if (!a_mn && !b_mn) return gemm_tc_kernel<0, 0, BN_, EPI_PLAIN>;
  if (!a_mn && b_mn) return gemm_tc_kernel<0, 1, BN_, EPI_PLAIN>;
  if (a_mn && !b_mn) return gemm_tc_kernel<1, 0, BN_, EPI_PLAIN>;
  return gemm_tc_kernel<1, 1, BN_, EPI_PLAIN>;
}
inline TcKernel tc_kernel_for(int a_mn, int b_mn, int bn, int epi) {
  return bn == 128 ? tc_kernel_for_bn<128>(a_mn, b_mn, epi) : bn == 32 ? tc_kernel_for_bn<32>(a_mn, b_mn, epi) : tc_kernel_for_bn<64>(a_mn, b_mn, epi);
}
inline int tc_max_stages(int bn) { return bn == 128 ? TcCfg<128>::MAX_STAGES : bn == 32 ? TcCfg<32>::MAX_STAGES : TcCfg<64>::MAX_STAGES; }
inline int tc_smem_for(int bn, int stages) {
  return bn == 128 ? TcCfg<128>::smem_bytes(stages) : bn == 32 ? TcCfg<32>::smem_bytes(stages) : TcCfg<64>::smem_bytes(stages);
}
inline cudaError_t tc_prepare(const void *fn, int bn) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_for(bn, tc_max_stages(bn)));
  if (e != cudaSuccess) return e;
  // two 98 KB CTAs only share an SM if the carve-out leaves room for both
  return cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
inline cudaError_t tc_prepare_all() {
  for (int bn : {32, 64, 128})
    for (int epi : {EPI_FWD, EPI_DX, EPI_PLAIN})
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          TcKernel k = tc_kernel_for(a, b, bn, epi);
          if (!k) continue;
          cudaError_t e = tc_prepare((const void *)k, bn);
          if (e != cudaSuccess) return e;
        }
  return cudaSuccess;
}

}  // namespace dqnb
