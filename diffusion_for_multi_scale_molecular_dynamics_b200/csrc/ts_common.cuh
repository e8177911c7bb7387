// Shared pieces of the "activations in tensor memory" kernels (egnn_edge_ts.cuh, egnn_node_ts.cuh): geometry of the
// epilogue (two fat warps per TMEM lane quarter, 32 accumulator columns per thread and chunk), the tcgen05.mma .ts /
// tcgen05.ld / tcgen05.st wrappers, the hi / -lo split, packed fp32 arithmetic and the SiLU.
// Include inside namespace b200md (after tc_common.cuh, with `using namespace tc`).
#pragma once

#ifndef B200MD_TS_NSPLIT
#define B200MD_TS_NSPLIT 2
#endif
constexpr int kTsNSplit = B200MD_TS_NSPLIT;       // column groups of an accumulator, computed one after the other
constexpr int kTsBStageBytes = (kMaxN / kTsNSplit) * kRowBytes;   // one weight part of a K chunk for ONE column group
constexpr int kTsBStages = 3 * kTsNSplit * 2;     // 192 KB in all
constexpr int kTsSlots = 4;                       // K-chunk slots of the "A ready" barriers
constexpr int kTsParts = 2;                       // epilogue warps per lane quarter
constexpr int kTsCols = 64 / kTsParts;            // accumulator columns one epilogue thread converts per chunk
constexpr int kTsEpiWarps = 4 * kTsParts;
// Warp roles.  Default (B200MD_TS_WARPGROUPS = 0): 10 warps, 0 = weight producer, 1 = MMA issuer, 2-9 = epilogue; 168
// registers per thread (a 320-thread block cannot be launched with more: the register file is handed out in units of
// four warps, which is also why __maxnreg__(200) fails with "too many resources requested").
// B200MD_TS_WARPGROUPS = 1 lays the roles out by WARPGROUP so that registers can be moved between them with setmaxnreg:
//   warps 0-3  : 0 = weight producer, 1 = MMA issuer, 2-3 idle     -> shrink to kTsRegsControl registers per thread
//   warps 4-11 : the 8 epilogue warps (two warpgroups)             -> grow to kTsRegsEpilogue
// (384 threads are compiled for 168 registers; the control group gives back (168 - 56) x 128 registers and each epilogue
// group takes (224 - 168) x 128; the setmaxnreg must sit at the top of each group's branch, ptxas budgets the code it
// dominates.)  Measured on one box, 4096 x 64 atoms: edge launch 18.8 ms (default) vs 19.2 ms (warpgroups) vs 19.5 ms
// (warpgroups + the next tile's P / Q rows requested one chunk ahead, B200MD_TS_PQ_AHEAD): with 227 KB of shared memory
// there is almost no L1 left, and 64 KB of gathers in flight per SM thrash it -- more registers do not buy deeper
// prefetch here.  Kept as a compile-time option.
#ifndef B200MD_TS_WARPGROUPS
#define B200MD_TS_WARPGROUPS 0
#endif
constexpr int kTsFirstEpiWarp = B200MD_TS_WARPGROUPS ? 4 : 2;
constexpr int kTsThreads = 32 * (kTsFirstEpiWarp + kTsEpiWarps);
constexpr int kTsRegsControl = 56, kTsRegsEpilogue = 224;
template <int N>
__device__ __forceinline__ void ts_regs_release() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void ts_regs_acquire() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// Index of the last column group (of `nh` columns each) that accumulator chunk c (columns [64 c, 64 c + 64)) reads.
// Compares instead of (64 c + 64 + nh - 1) / nh - 1: an integer division is an I2F / MUFU.RCP / F2I chain on the SFU
// pipe, where it queues behind the SiLU phase of the other epilogue warp (measured: 440 cycles per chunk).
template <int NSPLIT = kTsNSplit>
__device__ __forceinline__ int ts_last_group_of_chunk(int c, int nh) {
  int need = 0;
#pragma unroll
  for (int h = 1; h < NSPLIT; ++h) need += (c * 64 + 64 > h * nh) ? 1 : 0;
  return need;
}

// ---- CTA pair (tcgen05 cta_group::2) ----
// M = 256 over the two CTAs of the pair (128 rows each), N = n in all (each CTA holds n / 2 rows of B)
__device__ __forceinline__ uint32_t make_idesc_pair(int n, uint32_t format) {
  return (1u << 4) | (format << 7) | (format << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((2 * kTileM) >> 4) << 24);
}
// executed by the same warp of BOTH CTAs (same slot address)
__device__ __forceinline__ void tmem_alloc_pair(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrive on the barrier at this CTA-relative address in every CTA of `cta_mask` when the pair's MMAs issued so far are complete
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
// arrive on the barrier at this CTA-relative address in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar),
      "r"(cta)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T
// `b_desc_lo`: low word of the K-major SWIZZLE_128B descriptor (desc_lo_sw128); the high word is a constant.
template <bool PAIR = false>
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t idesc,
                                            uint32_t accumulate) {
  if (PAIR) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 bd;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "mov.b64 bd, {%2, %5};\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], bd, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_desc_lo), "r"(idesc), "r"(accumulate), "n"(kDescHiSw128)
        : "memory");
    return;
  }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 bd;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 bd, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_desc_lo), "r"(idesc), "r"(accumulate), "n"(kDescHiSw128)
      : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// tcgen05.ld in two halves, so independent work can be placed between the issue and the wait; the wait names the
// registers as in/out operands, which keeps every use of them after it
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// kTsCols accumulator columns of this lane -> registers (complete on return)
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, uint32_t (&r)[kTsCols]) {
#pragma unroll
  for (int i = 0; i < kTsCols / 16; ++i) tmem_ld16_issue(taddr + 16 * i, r + 16 * i);
#pragma unroll
  for (int i = 0; i < kTsCols / 16; ++i) tmem_ld_wait16(r + 16 * i);
}
// two fp32 -> packed f16x2 (first argument in the LOW half); out-of-range values become +-inf (no saturation): they
// poison the chain and are caught by the finite check on the tile's result
__device__ __forceinline__ uint32_t pack_f16x2_inf(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
constexpr uint32_t kIdescNegateA = 1u << 13;
constexpr uint32_t kIdescA2 = kIdescNegateA | kIdescAFormatBf16;   // A_2 is stored negated, in BF16 (tc_common.cuh)
// kTsCols activations -> per K step (16 values) [hi x 8 | -lo x 8] packed fp16 pairs: the A-operand image of those
// columns.  hi = fp16(x); the second part holds fp16(hi - x) = -lo, one mixed-precision subtract per element
// (sub.f32.f16, SASS FHADD) instead of an unpack and a subtract; the A_2 B_1 MMA negates it back (a_negate bit of the
// instruction descriptor).
__device__ __forceinline__ void split_hi_lo(const float (&v)[kTsCols], uint32_t (&w)[kTsCols]) {
#pragma unroll
  for (int s = 0; s < kTsCols / 16; ++s)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x0 = v[16 * s + 2 * i], x1 = v[16 * s + 2 * i + 1];
      const uint32_t hi = pack_f16x2_inf(x0, x1);
      float n0, n1;
      asm("{\n"
          ".reg .b16 h0, h1;\n"
          "mov.b32 {h0, h1}, %2;\n"
          "sub.rn.f32.f16 %0, h0, %3;\n"
          "sub.rn.f32.f16 %1, h1, %4;\n"
          "}\n"
          : "=f"(n0), "=f"(n1)
          : "r"(hi), "f"(x0), "f"(x1));
      w[16 * s + i] = hi;
      w[16 * s + 8 + i] = pack_lo_x2(n0, n1);
    }
}
// Packed fp32 pairs (FADD2 / FMUL2 / FFMA2: two lanes per instruction) for the bias add and the SiLU arithmetic.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// The TS kernels keep every activation multiplied by kF16ActScale (tc_common.cuh): `x` below is kF16ActScale times the
// pre-activation and the result kF16ActScale times its SiLU -- the factor rides on the multiplier and the bias of the
// FMA that produces x and on the exp2 constant, so it costs no instruction.
constexpr float kTsExpConst = -1.4426950408889634f / kF16ActScale;
// (silu(x0), silu(x1)) of a packed pair: x / (1 + 2^(-x log2 e)) with the SFU exp2 / reciprocal (about 2 ulp each, far
// below the 1e-5 parity bound).  Moving the reciprocal to three Newton steps on the FMA pipe halves the SFU load and
// saves 2.6 % of the cycles, but the extra FMA work lowers the clock the 1000 W power cap allows (1657 -> 1605 MHz):
// +-1 % on the launch, within noise, so the SFU version stays.
__device__ __forceinline__ void silu2(uint64_t x, float& o0, float& o1) {
  float t0, t1, e0, e1, r0, r1;
  unpack2(mul2(x, pack2(kTsExpConst, kTsExpConst)), t0, t1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  unpack2(add2(pack2(e0, e1), pack2(1.0f, 1.0f)), t0, t1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
  unpack2(mul2(x, pack2(r0, r1)), o0, o1);
}
// The same with an SFU "token" dependency: `dep` collects the bits of every reciprocal, so sfu_release(.., dep) cannot
// be scheduled before the last SFU result exists.
__device__ __forceinline__ void silu2_dep(uint64_t x, float& o0, float& o1, uint32_t& dep) {
  float t0, t1, e0, e1, r0, r1;
  unpack2(mul2(x, pack2(kTsExpConst, kTsExpConst)), t0, t1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  unpack2(add2(pack2(e0, e1), pack2(1.0f, 1.0f)), t0, t1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
  dep |= __float_as_uint(r0) | __float_as_uint(r1);
  unpack2(mul2(x, pack2(r0, r1)), o0, o1);
}
// Shared reciprocals (batch inversion).  With a_i = 1 + 2^(t_i):
//   pair : r = 1 / (a_0 a_1),  1 / a_0 = r a_1,  1 / a_1 = r a_0                       3 MUFU per 2 elements instead of 4
//   quad : r = 1 / (a_0 a_1 a_2 a_3),  1 / (a_0 a_1) = r (a_2 a_3), ... then as a pair  5 MUFU per 4 elements instead of 8
// for a few FMUL / FMNMX more.  The SiLU phases hold the SFU pipe (8.1 cycles per warp instruction, pipe_rates), and
// under the power cap a MUFU costs more than the multiplies that replace it: A/B on one box, pair: edge launch 19.2 ->
// 18.9 ms.  t is clamped so that the product of the a_i stays finite (0 x inf would poison the OTHER elements of the
// group): 2^63 per element for a pair, 2^31 for a quad -- beyond the clamp silu(x) is |x| 2^-31 at most in magnitude, below
// half an ulp of the unit-scale activations next to it.
__device__ __forceinline__ void silu2_pair(uint64_t x, float& o0, float& o1, uint32_t& dep) {
  float t0, t1, e0, e1, a0, a1, r;
  unpack2(mul2(x, pack2(kTsExpConst, kTsExpConst)), t0, t1);
  t0 = fminf(t0, 63.0f);
  t1 = fminf(t1, 63.0f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  unpack2(add2(pack2(e0, e1), pack2(1.0f, 1.0f)), a0, a1);
  const float prod = a0 * a1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(prod));
  dep |= __float_as_uint(r);
  unpack2(mul2(x, mul2(pack2(r, r), pack2(a1, a0))), o0, o1);
}
__device__ __forceinline__ void silu4_quad(uint64_t x01, uint64_t x23, float& o0, float& o1, float& o2, float& o3, uint32_t& dep) {
  float t0, t1, t2, t3, e0, e1, e2, e3, a0, a1, a2, a3, p01, p23, r;
  unpack2(mul2(x01, pack2(kTsExpConst, kTsExpConst)), t0, t1);
  unpack2(mul2(x23, pack2(kTsExpConst, kTsExpConst)), t2, t3);
  t0 = fminf(t0, 31.0f); t1 = fminf(t1, 31.0f); t2 = fminf(t2, 31.0f); t3 = fminf(t3, 31.0f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(t2));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(t3));
  unpack2(add2(pack2(e0, e1), pack2(1.0f, 1.0f)), a0, a1);
  unpack2(add2(pack2(e2, e3), pack2(1.0f, 1.0f)), a2, a3);
  unpack2(mul2(pack2(a0, a2), pack2(a1, a3)), p01, p23);
  const float prod = p01 * p23;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(prod));
  dep |= __float_as_uint(r);
  float r01, r23;
  unpack2(mul2(pack2(r, r), pack2(p23, p01)), r01, r23);   // 1 / (a_0 a_1), 1 / (a_2 a_3)
  unpack2(mul2(x01, mul2(pack2(r01, r01), pack2(a1, a0))), o0, o1);
  unpack2(mul2(x23, mul2(pack2(r23, r23), pack2(a3, a2))), o2, o3);
}
// four elements, MODE 0: one reciprocal each, 1: per pair, 2: per quad
template <int MODE>
__device__ __forceinline__ void silu4(uint64_t x01, uint64_t x23, float& o0, float& o1, float& o2, float& o3, uint32_t& dep) {
  if (MODE == 2) {
    silu4_quad(x01, x23, o0, o1, o2, o3, dep);
  } else if (MODE == 1) {
    silu2_pair(x01, o0, o1, dep);
    silu2_pair(x23, o2, o3, dep);
  } else {
    silu2_dep(x01, o0, o1, dep);
    silu2_dep(x23, o2, o3, dep);
  }
}
// SFU token between the two epilogue warps that share a scheduler (named barriers of 64 threads, strict alternation):
// both warps are released by the same accumulator barrier and would otherwise run their SiLU phases (SFU-bound) at the
// same time and their split / store phases (no SFU) at the same time.
__device__ __forceinline__ void sfu_acquire(uint32_t bar_id) { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); }
__device__ __forceinline__ void sfu_arrive(uint32_t bar_id) { asm volatile("bar.arrive %0, 64;" ::"r"(bar_id) : "memory"); }
// (dep >> 31) + (~dep >> 31) == 1 for every dep, which the assembler cannot see: the arrive stays behind the SFU results
__device__ __forceinline__ void sfu_release(uint32_t bar_id, uint32_t dep) {
  asm volatile(
      "{\n"
      ".reg .b32 t1, t2;\n"
      "shr.u32 t1, %1, 31;\n"
      "not.b32 t2, %1;\n"
      "shr.u32 t2, t2, 31;\n"
      "add.u32 t1, t1, t2;\n"
      "add.u32 t1, t1, %0;\n"
      "sub.u32 t1, t1, 1;\n"
      "bar.arrive t1, 64;\n"
      "}\n" ::"r"(bar_id),
      "r"(dep)
      : "memory");
}
// shared-memory load by window address (keeps the access an LDS: pointers derived from the aligned dynamic-smem base
// lose their address space and become generic loads)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// the same, ordered against the other volatile shared-memory / cp.async statements (slots that are refilled)
__device__ __forceinline__ float4 lds128_ordered(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
// 16 bytes global -> shared, L2 only (LDGSTS.BYPASS: no register, no L1 line); complete for the issuing thread after
// cp_async_wait_all
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
// this thread's arrival on `bar` happens when all its cp.async issued so far have landed (the barrier's count includes it)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* addr, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float x) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(x) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t saddr, float x) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(x) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void ts_epilogue_barrier() { asm volatile("bar.sync 1, %0;" ::"r"(kTsEpiWarps * 32) : "memory"); }
