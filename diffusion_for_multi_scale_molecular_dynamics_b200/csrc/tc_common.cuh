// Shared machinery of the tcgen05 kernels (egnn_tc.cu: edge MLPs, egnn_node_tc.cu: node MLPs).
//
// A kernel processes tiles of 128 rows (edges or nodes) through a *program*: a fixed sequence of Linear(K -> N)
// layers whose A operand (the activations of the 128 rows) is produced chunk by chunk (32 K-columns at a time) by the
// "epilogue" warps into a 2-stage shared-memory ring, and whose B operand (the weights) is streamed by a producer warp
// from a pre-split, pre-swizzled global image into a 4-stage ring with cp.async.bulk + mbarrier complete_tx.
// One thread issues the MMAs into two 256-column fp32 TMEM accumulators that alternate between consecutive GEMMs.
//
// fp32 accuracy on the tensor cores (the parity contract is 1e-5 relative; one TF32 pass gives 2^-11): every operand is
// split as x = x_1 + x_2 and D += A_1 B_1^T + A_2 B_1^T + A_1 B_2^T (the dropped A_2 B_2 is 2^-24 relative):
//   mode kTf32x3 : x_1 = cvt.rna.tf32(x), x_2 = x - x_1 (exact), kind::tf32, K chunk = 32 elements (128-byte rows),
//                  8 bytes per weight and chunk streamed from L2;
//   mode kF16x3  : x_1 = fp16(x), x_2 = fp16(x - x_1): the same 11-bit significands as TF32 in 2-byte containers,
//                  kind::f16, K chunk = 64 elements, 4 bytes per weight -- half the L2 -> SM traffic (the measured
//                  bottleneck of the TF32 version) and half the tensor-pipe time.  fp16 has a 5-bit exponent, so the
//                  exponent is managed explicitly:
//                    * weights: every matrix is multiplied by a power of two 2^s chosen at pack time so that its
//                      largest magnitude lands in [2^14, 2^15) (tc_pack_kernel); w_2 = fp16(2^s w - w_1) is then a
//                      NORMAL fp16 number for every element within 2^-17 of the largest one, whatever the scale of a
//                      trained checkpoint's weights (1e-3 or 1e+2 alike).  The accumulator holds 2^s times the
//                      product; the epilogue undoes it exactly with one FMA, v = acc * 2^-s + bias (the factor is
//                      stored next to the image, see weight_unscale_offset).
//                    * activations: stored multiplied by kF16ActScale = 4 (see there): x_2 is subnormal only for
//                      |x| < 0.0625 (absolute error 7.5e-9 per element), and an activation above 16376 becomes inf: the
//                      epilogue raises B200MD_STATUS_TC_RANGE in the device status word and the host re-runs the call
//                      with the 3xTF32 kernels (fp32 exponent range).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace b200md {
namespace tc {

constexpr int kTileM = 128;            // rows per tile = UMMA M
constexpr int kRowBytes = 128;         // operand rows are one 128-byte swizzle row per K chunk
constexpr int kMaxN = 256;             // widest layer (UMMA N limit)
constexpr int kAStages = 2;
constexpr int kBStages = 4;
constexpr int kEpiWarps = 16;           // 4 per TMEM lane quarter: each owns a quarter of a chunk's columns
constexpr int kEpiParts = kEpiWarps / 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kMaxGemms = 10;
constexpr int kAPartBytes = kTileM * 128;            // one 128-byte-per-row operand of an A stage (16 KB)
constexpr int kAStageBytes = 2 * kAPartBytes;        // part 0: hi (fp32 bits), part 1: lo (tf32) or [lo | hi] (f16)
constexpr int kBStageBytes = kMaxN * 128;            // one weight part of a K chunk (32 KB)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSpinLimit = 1u << 24;            // turns a protocol deadlock into a trap instead of a hang

enum Precision { kTf32x3 = 0, kF16x3 = 1, kF16x3Ts = 2 };   // kF16x3Ts: edge kernel only, A operand in TMEM (egnn_tc.cu)
template <int PREC>
struct Mode {
  static constexpr int kChunk = (PREC == kF16x3) ? 64 : 32;   // K elements per chunk (128 bytes per row)
  static constexpr int kVals = kChunk / kEpiParts;            // activation values one epilogue thread handles per chunk
};
// CTAs per thread-block cluster.  The CTAs of a cluster walk the same weight stream in lock-step: each one fetches
// 1/kCluster of every weight part from L2 and multicasts it into the shared memory of all of them, which divides the
// L2 -> SM traffic (the measured bottleneck of the single-CTA version: 64 KB per K chunk and SM) by kCluster.
constexpr int kCluster = 1;
// Global weight image of one Linear(K -> N), every part already in the 128B-swizzled K-major shared-memory layout
// (N rows of 128 bytes): first the TF32 image -- per 32-element K chunk [w_1 (fp32 bits)] [w_2 (fp32 bits)], 2 K N
// floats -- then the fp16 image of 2^s w -- per 64-element K chunk [w_1 (fp16)] [w_2 (fp16)], K N floats -- then a
// 256-byte trailer: [max |w|] [2^-s] (the TF32 image is not scaled: TF32 has the fp32 exponent).
constexpr int kWeightTrailerFloats = 64;
__host__ __device__ inline int64_t weight_image_floats(int K, int N) { return (int64_t)3 * K * N + kWeightTrailerFloats; }
__host__ __device__ inline int64_t f16_image_offset(int K, int N) { return (int64_t)2 * K * N; }
__host__ __device__ inline int64_t weight_absmax_offset(int K, int N) { return (int64_t)3 * K * N; }
__host__ __device__ inline int64_t weight_unscale_offset(int K, int N) { return (int64_t)3 * K * N + 1; }

struct Program {
  const float* wtc;      // base of the weight images
  const float* packed;   // fp32 parameter image (biases)
  int n_gemms;
  int gK[kMaxGemms], gN[kMaxGemms];
  int64_t gW[kMaxGemms];   // offset (floats) of the layer's image in wtc
  int64_t gB[kMaxGemms];   // offset of the bias in packed, or -1
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait costs ~200 cycles even on a completed phase and may put the warp to sleep).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_test_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epilogue_barrier() { asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// Same, delivered to the same CTA-relative address (data and mbarrier) of every CTA in `cta_mask`.
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// The same descriptor as two words: the high word is a constant, the low word advances by 1 per 16 bytes of K offset
// inside the swizzle atom (shared-memory addresses are < 2^18, so the 14-bit address field never carries).
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
// One lane of the (converged) warp.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, issued by one thread for the CTA.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int PREC>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (PREC == kF16x3) umma_f16(d_tmem, a_desc, b_desc, idesc, accumulate);
  else umma_tf32(d_tmem, a_desc, b_desc, idesc, accumulate);
}
// Same with both operand descriptors given by their low words (desc_lo_sw128); the high word is a constant.
template <int PREC>
__device__ __forceinline__ void umma_lo(uint32_t d_tmem, uint32_t a_desc_lo, uint32_t b_desc_lo, uint32_t idesc,
                                        uint32_t accumulate) {
  if (PREC == kF16x3)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 ad, bd;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "mov.b64 ad, {%1, %5};\n"
        "mov.b64 bd, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_desc_lo), "r"(b_desc_lo), "r"(idesc), "r"(accumulate), "n"(kDescHiSw128)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 ad, bd;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "mov.b64 ad, {%1, %5};\n"
        "mov.b64 bd, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_desc_lo), "r"(b_desc_lo), "r"(idesc), "r"(accumulate), "n"(kDescHiSw128)
        : "memory");
}
// Arrive on an mbarrier when every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Same, arriving on the barrier at this CTA-relative address in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) { tmem_ld8(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }

// K-major, 128B-swizzled operand descriptor (cute::UMMA::SmemDescriptor): 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row core-matrix groups
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor: fp32 accumulate, both operands K-major, M = 128, N = n; format 2 = TF32, 0 = F16.
__device__ __forceinline__ uint32_t make_idesc(int n, uint32_t format) {
  return (1u << 4) | (format << 7) | (format << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// two fp32 -> packed f16x2 (first argument in the LOW half), round to nearest, saturating to the finite range
__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// Exponent management of the ACTIVATIONS in the 3xFP16 modes.  x_2 = fp16(x - x_1) is subnormal for |x| < 0.25: an
// absolute error of 2^-25 per activation, i.e. 6e-6 relative at |x| = 5e-3, which a following layer with large weights
// amplifies past the 1e-5 contract (test_weight_and_activation_scales_against_oracle[x1e-2] measured 1.3e-5).  Every A
// operand is therefore stored multiplied by kF16ActScale = 4 (exact; undone with the weight scale in the epilogue's
// FMA): the subnormal threshold drops to |x| < 0.0625 and the error floor to 7.5e-9, at the price of an fp16 overflow
// (-> B200MD_STATUS_TC_RANGE -> 3xTF32 fallback) at |x| > 16376 instead of 65504.  The window [0.0625, 16376] is
// centred on the O(0.1 .. 100) activations of SiLU networks.
// (Tried first: A_2 in BF16 -- fp32 exponent, 8 significant bits, no cliff at all -- with A_1 / B in FP16.  The
// instruction descriptor has separate a_format / b_format fields, but kind::f16 with a_format != b_format raises
// "illegal instruction" on B200; B200MD_F16X3_LO_BF16 = 1 keeps the experiment compilable.)
#ifndef B200MD_F16_ACT_SCALE
#define B200MD_F16_ACT_SCALE 4.0f
#endif
constexpr float kF16ActScale = B200MD_F16_ACT_SCALE;   // a power of two (1.0f = no scaling, for A/B builds)
#ifndef B200MD_F16X3_LO_BF16
#define B200MD_F16X3_LO_BF16 0
#endif
constexpr uint32_t kIdescAFormatBf16 = B200MD_F16X3_LO_BF16 ? (1u << 7) : 0u;
__device__ __forceinline__ uint32_t pack_lo_x2(float lo_elem, float hi_elem) {
#if B200MD_F16X3_LO_BF16
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
#else
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
#endif
}

// SiLU with the SFU exp2 / reciprocal (about 2 ulp each): far below the 1e-5 parity bound.
__device__ __forceinline__ float silu_fast(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return v * r;
}

// byte offset of (row, 16-byte chunk j) inside a 128B-swizzled K-major operand
__device__ __forceinline__ uint32_t sw128_offset(int row, int j) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
}

// inclusive segmented sum over the lanes of a warp; seg_start = lane of this lane's segment head
__device__ __forceinline__ float seg_scan(float v, int lane, int seg_start) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, v, off);
    if (lane - off >= seg_start) v += t;
  }
  return v;
}

// ---- the pipeline -----------------------------------------------------------------------------------------------
struct PipeSmem {
  uint32_t a, b, bars, tmem_slot, end;   // byte offsets from the 1024-aligned base
};
__host__ __device__ inline PipeSmem pipe_smem() {
  PipeSmem L;
  uint32_t o = 0;
  L.a = o; o += kAStages * kAStageBytes;
  L.b = o; o += kBStages * kBStageBytes;
  L.bars = o; o += 16 * 8;
  L.tmem_slot = o; o += 16;
  L.end = o;
  return L;
}

template <int PREC>
struct Pipe {
  uint8_t* smem;        // 1024-aligned base
  uint32_t s_base;      // same, shared-window address
  uint32_t bar0;
  uint32_t tmem_base;
  uint32_t ca, gc;      // epilogue-side counters: A chunks produced, GEMMs consumed

  __device__ __forceinline__ uint32_t a_full(uint32_t s) const { return bar0 + 8u * s; }
  __device__ __forceinline__ uint32_t a_empty(uint32_t s) const { return bar0 + 8u * (2 + s); }
  __device__ __forceinline__ uint32_t b_full(uint32_t s) const { return bar0 + 8u * (4 + s); }
  __device__ __forceinline__ uint32_t b_empty(uint32_t s) const { return bar0 + 8u * (8 + s); }
  __device__ __forceinline__ uint32_t acc_full(uint32_t s) const { return bar0 + 8u * (12 + s); }

  // All threads of the CTA.  Ends with a __syncthreads().
  __device__ __forceinline__ void setup(uint8_t* aligned_smem) {
    const PipeSmem L = pipe_smem();
    smem = aligned_smem;
    s_base = smem_u32(smem);
    bar0 = s_base + L.bars;
    ca = 0; gc = 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), kEpiWarps); mbar_init(a_empty(s), 1); }
      for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), kCluster); }
      for (int s = 0; s < 2; ++s) mbar_init(acc_full(s), 1);
      fence_barrier_init();
    }
    if ((threadIdx.x >> 5) == 1) tmem_alloc(s_base + L.tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync();  // every CTA's barriers exist before any multicast can reach them
    tc_fence_after();
    tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);
  }

  __device__ __forceinline__ void teardown() {
    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync();  // no CTA leaves while a peer may still multicast into it
    tc_fence_after();
    if ((threadIdx.x >> 5) == 1) {
      __syncwarp();
      tmem_dealloc(tmem_base, kTmemCols);
    }
  }

  // ===== weight producer: one thread =====
  // `n_iterations` = tiles this CTA walks through (identical for the CTAs of a cluster).
  __device__ void run_producer(const Program& p, int n_iterations) const {
    const PipeSmem L = pipe_smem();
    const uint32_t rank = kCluster > 1 ? cluster_ctarank() : 0u;
    const uint16_t all_ctas = (uint16_t)((1u << kCluster) - 1u);
    uint32_t cb = 0;
    for (int it = 0; it < n_iterations; ++it) {
      for (int g = 0; g < p.n_gemms; ++g) {
        const int nk = p.gK[g] / Mode<PREC>::kChunk;
        const uint32_t bytes = (uint32_t)p.gN[g] * kRowBytes;
        const uint32_t slice = bytes / kCluster;  // rows [rank N/kCluster, (rank+1) N/kCluster): whole swizzle atoms
        const int64_t part_floats = (int64_t)p.gN[g] * (kRowBytes / 4);
        const float* src = p.wtc + p.gW[g] + (PREC == kF16x3 ? f16_image_offset(p.gK[g], p.gN[g]) : 0);
        for (int part = 0; part < 2 * nk; ++part, ++cb) {  // chunk 0 part 1, chunk 0 part 2, chunk 1 part 1, ...
          const uint32_t s = cb % kBStages;
          mbar_wait(b_empty(s), ((cb / kBStages) & 1u) ^ 1u);   // stage s is free in EVERY CTA of the cluster
          mbar_arrive_expect_tx(b_full(s), bytes);
          const float* part_src = src + (int64_t)part * part_floats;
          if (kCluster > 1)
            bulk_g2s_multicast(s_base + L.b + s * kBStageBytes + rank * slice, part_src + rank * (slice / 4), slice,
                               b_full(s), all_ctas);
          else
            bulk_g2s(s_base + L.b + s * kBStageBytes, part_src, bytes, b_full(s));
        }
      }
    }
  }

  // ===== MMA issuer: the whole warp walks the loop (converged), one elected lane issues =====
  // With a single thread inside the loop every operand sits in a vector register and the compiler wraps each
  // tcgen05.mma in an ELECT / R2UR / branch "waterfall" (~40 instructions per MMA): the issuing thread, not the tensor
  // pipe, then paces the kernel.  Converged, all addresses and descriptors live in uniform registers.
  __device__ void run_mma(const Program& p, int n_iterations) const {
    const PipeSmem L = pipe_smem();
    const uint16_t all_ctas = (uint16_t)((1u << kCluster) - 1u);
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = elect_one();
    uint32_t mca = 0, cb = 0, mgc = 0;
    for (int it = 0; it < n_iterations; ++it) {
      for (int g = 0; g < p.n_gemms; ++g, ++mgc) {
        const int nk = p.gK[g] / Mode<PREC>::kChunk;
        const uint32_t idesc = make_idesc(p.gN[g], PREC == kF16x3 ? 0u : 2u);
        const uint32_t d_tmem = tmem + (mgc & 1u) * kMaxN;
        for (int c = 0; c < nk; ++c, ++mca) {
          const uint32_t sa = mca % kAStages;
          mbar_wait(a_full(sa), (mca / kAStages) & 1u);
          const uint32_t a1 = desc_lo_sw128(s_base + L.a + sa * kAStageBytes), a2 = desc_lo_sw128(s_base + L.a + sa * kAStageBytes + kAPartBytes);
          uint32_t sb = cb % kBStages;
          mbar_wait(b_full(sb), (cb / kBStages) & 1u);
          tc_fence_after();
          uint32_t bd = desc_lo_sw128(s_base + L.b + sb * kBStageBytes);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 4 K-steps of 32 bytes: 8 tf32 or 16 fp16 elements
              umma_lo<PREC>(d_tmem, a1 + 2 * k, bd + 2 * k, idesc, (c | k) != 0);  // A_1 B_1
              umma_lo<PREC>(d_tmem, a2 + 2 * k, bd + 2 * k, PREC == kF16x3 ? (idesc | kIdescAFormatBf16) : idesc, 1u);   // A_2 B_1
            }
            if (kCluster > 1) umma_commit_multicast(b_empty(sb), all_ctas); else umma_commit(b_empty(sb));
          }
          __syncwarp();
          ++cb;
          sb = cb % kBStages;
          mbar_wait(b_full(sb), (cb / kBStages) & 1u);
          tc_fence_after();
          bd = desc_lo_sw128(s_base + L.b + sb * kBStageBytes);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // A_1 B_2
              umma_lo<PREC>(d_tmem, a1 + 2 * k, bd + 2 * k, idesc, 1u);
            if (kCluster > 1) umma_commit_multicast(b_empty(sb), all_ctas); else umma_commit(b_empty(sb));
            umma_commit(a_empty(sa));
          }
          __syncwarp();
          ++cb;
        }
        if (leader) umma_commit(acc_full(mgc & 1u));
        __syncwarp();
      }
    }
  }

  // ===== epilogue side =====
  // Wait for the accumulator of the next GEMM in program order; returns its TMEM column base.
  __device__ __forceinline__ uint32_t wait_accumulator() {
    mbar_wait(acc_full(gc & 1u), (gc >> 1) & 1u);
    tc_fence_after();
    const uint32_t col = (gc & 1u) * kMaxN;
    ++gc;
    return col;
  }

  // Publish kVals activation values of `row` (columns hh*kVals .. of the next A chunk).  Called by every epilogue
  // thread (all kEpiWarps warps) exactly once per chunk, in program order.  Returns max |v| (fp16 range guard).
  __device__ __forceinline__ float push_chunk(int row, int hh, int lane, const float (&v)[Mode<PREC>::kVals]) {
    const PipeSmem L = pipe_smem();
    const uint32_t sa = ca % kAStages;
    mbar_wait(a_empty(sa), ((ca / kAStages) & 1u) ^ 1u);
    uint8_t* p1 = smem + L.a + sa * kAStageBytes;
    uint8_t* p2 = p1 + kAPartBytes;
    float vmax = 0.f;
    // both modes: this thread owns 32 bytes of the 128-byte row = 16-byte units hh*2 and hh*2+1 of each part
    if (PREC == kTf32x3) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float4 a1, a2;
        a1.x = tf32_rna(v[4 * j + 0]); a2.x = v[4 * j + 0] - a1.x;
        a1.y = tf32_rna(v[4 * j + 1]); a2.y = v[4 * j + 1] - a1.y;
        a1.z = tf32_rna(v[4 * j + 2]); a2.z = v[4 * j + 2] - a1.z;
        a1.w = tf32_rna(v[4 * j + 3]); a2.w = v[4 * j + 3] - a1.w;
        const uint32_t off = sw128_offset(row, hh * 2 + j);
        *reinterpret_cast<float4*>(p1 + off) = a1;
        *reinterpret_cast<float4*>(p2 + off) = a2;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) {  // 8 values = 16 bytes of fp16 per iteration
        uint32_t w1[4], w2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x0 = v[8 * j + 2 * i] * kF16ActScale, x1 = v[8 * j + 2 * i + 1] * kF16ActScale;   // see kF16ActScale
          vmax = fmaxf(vmax, fmaxf(fabsf(x0), fabsf(x1)));
          w1[i] = pack_f16x2(x0, x1);
          const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&w1[i]));
          w2[i] = pack_lo_x2(x0 - back.x, x1 - back.y);
        }
        const uint32_t off = sw128_offset(row, hh * 2 + j);
        *reinterpret_cast<uint4*>(p1 + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<uint4*>(p2 + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_full(sa));
    ++ca;
    return vmax;
  }
};

constexpr float kF16Safe = 65000.0f;   // activations above this saturate in fp16: raise B200MD_STATUS_TC_RANGE

// Tile schedule: cluster k of n_clusters takes tiles (i * n_clusters + k) * kCluster + rank for i = 0, 1, ...; every
// CTA of a cluster runs the same number of iterations (a tile index >= n_tiles is an empty tile).
struct TileSchedule {
  int n_iterations, first, stride;
  __device__ __forceinline__ int tile(int it) const { return first + it * stride; }
};
__device__ __forceinline__ TileSchedule make_schedule(int n_tiles) {
  const int n_clusters = gridDim.x / kCluster;
  const int cluster = blockIdx.x / kCluster;
  const int rank = kCluster > 1 ? (int)cluster_ctarank() : 0;
  const int groups = (n_tiles + kCluster - 1) / kCluster;  // groups of kCluster consecutive tiles
  TileSchedule s;
  s.n_iterations = groups > cluster ? (groups - cluster + n_clusters - 1) / n_clusters : 0;
  s.first = cluster * kCluster + rank;
  s.stride = n_clusters * kCluster;
  return s;
}

// Host side: launch `kernel` with thread-block clusters of kCluster CTAs; the grid is rounded to whole clusters.
template <typename Args>
inline cudaError_t launch_clustered(void (*kernel)(const Args), unsigned wanted_ctas, unsigned max_ctas, size_t smem,
                                    cudaStream_t st, const Args& args) {
  unsigned grid = wanted_ctas < max_ctas ? wanted_ctas : max_ctas;
  grid = ((grid + kCluster - 1) / kCluster) * kCluster;
  if (grid > max_ctas) grid = (max_ctas / kCluster) * kCluster;
  if (grid < (unsigned)kCluster) grid = kCluster;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args);
}

// Pack one Linear's weights ([K][N] fp32, transposed storage) into its tensor-core image.
int pack_weight_image(const float* wt, int K, int N, float* dst, cudaStream_t st);

}  // namespace tc
}  // namespace b200md
