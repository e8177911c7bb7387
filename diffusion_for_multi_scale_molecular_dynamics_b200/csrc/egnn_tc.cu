// EGNN edge kernel on the 5th-generation tensor cores (tcgen05 / TMEM), fp32-accurate through a 3xTF32 split.
//
// Reference arithmetic: models/egnn.py:136-160 (message MLP), :189-215 (coordinate MLP and update), :217-238
// (coord2radial), models/egnn_utils.py:11-70 (segment sum / mean).  Same fusion as the fp32 CUDA-core kernel of
// egnn.cu -- gather, distance featurisation, message MLP, coordinate MLP and the segmented reductions in one
// launch, no per-edge tensor ever reaches HBM -- but every Linear(K -> N) of the chain runs as
//
//     D[128 edges, N] (TMEM, fp32)  +=  A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T        (tcgen05.mma kind::tf32)
//
// with x = x_hi + x_lo, x_hi = rna_tf32(x), x_lo = x - x_hi (exact in fp32; the tensor core reads its top 11 bits).
// The dropped term A_lo B_lo is 2^-24 relative, i.e. below fp32 rounding, which is what the 1e-5 parity bound needs;
// a single TF32 pass (2^-11) would not meet it.
//
// One persistent CTA per SM, tile = 128 consecutive edges of the (src-sorted) edge list.  Warp roles:
//   warp 0      bulk-copy producer: streams the pre-split, pre-swizzled weight chunks (L2-resident) into a 4-stage
//               shared-memory ring with cp.async.bulk + mbarrier complete_tx (no tensor map needed: the global
//               image already has the 128B-swizzled K-major layout the UMMA descriptor expects);
//   warp 1      TMEM allocator and single-thread MMA issuer;
//   warps 2-9   "epilogue" warps: build the first activation from P[src] + Q[dst] + r w_r, and between two Linear
//               layers read the accumulator (tcgen05.ld), add the bias, apply SiLU, split into hi/lo and write the
//               next layer's A operand chunk by chunk into a 2-stage ring, so that the MMA of layer g+1 on chunk c
//               overlaps the epilogue of layer g on chunk c+1.  They also do the segmented warp reductions
//               (messages -> agg, coordinate updates -> xagg) with one atomic per (segment, column).
// TMEM: two 256-column fp32 accumulators (all 512 columns), alternating between consecutive GEMMs.
#include <math_constants.h>

#include "egnn.cuh"

namespace b200md {

namespace {

constexpr int kTileM = 128;            // edges per tile = UMMA M
constexpr int kChunkK = 32;            // tf32 elements per K chunk: one 128-byte swizzle row
constexpr int kUmmaK = 8;              // K of one tcgen05.mma kind::tf32
constexpr int kAStages = 2;
constexpr int kBStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 32 * (2 + kEpiWarps);
constexpr int kMaxGemms = 10;
constexpr int kAPartBytes = kTileM * 128;            // one hi or lo operand of an A stage (16 KB)
constexpr int kAStageBytes = 2 * kAPartBytes;        // hi + lo
constexpr int kBStageBytes = kMaxWidth * 128;        // one hi or lo weight chunk (32 KB)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSpinLimit = 1u << 24;            // turns a protocol deadlock into a trap instead of a hang

struct TcArgs {
  const float* packed;   // fp32 parameter image (biases, w_r, last coordinate layer)
  const float* wtc;      // this layer's tensor-core weight image
  const float *xcur, *P, *Q;
  float *agg, *xagg;
  const int32_t *edge_src, *edge_dst, *n_edges;
  int D, M;
  int n_gemms, n_msg;    // GEMMs in the chain; the first n_msg belong to the message MLP
  int normalize, tanh_out;
  int64_t msg_wr, coord_last_w;
  int gK[kMaxGemms], gN[kMaxGemms];
  int64_t gW[kMaxGemms], gB[kMaxGemms];   // offsets (floats) into wtc / packed
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, issued by one thread for the CTA.
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
// Arrive on an mbarrier when every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
// cute::UMMA::InstrDescriptor: fp32 accumulate, tf32 x tf32, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// SiLU with the SFU exp2 / reciprocal (2 ulp each): the error is far below the 1e-5 parity bound.
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

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

struct SmemLayout {
  uint32_t a, b, bias, wr, wlast, dot, bars, tmem_slot, total;
};
__host__ __device__ inline SmemLayout smem_layout(int n_gemms) {
  SmemLayout L;
  uint32_t o = 0;
  L.a = o; o += kAStages * kAStageBytes;
  L.b = o; o += kBStages * kBStageBytes;
  L.bias = o; o += (uint32_t)n_gemms * kMaxWidth * 4;
  L.wr = o; o += kMaxWidth * 4;
  L.wlast = o; o += kMaxWidth * 4;
  L.dot = o; o += kTileM * 4;
  L.bars = o; o += 16 * 8;
  L.tmem_slot = o; o += 16;
  L.total = o;
  return L;
}

// ---- the kernel ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1) egnn_edge_tc_kernel(const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const SmemLayout L = smem_layout(a.n_gemms);
  const uint32_t s_base = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  float* s_wr = reinterpret_cast<float*>(smem + L.wr);
  float* s_wlast = reinterpret_cast<float*>(smem + L.wlast);
  float* s_dot = reinterpret_cast<float*>(smem + L.dot);
  volatile uint32_t* s_tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);
  const uint32_t bar0 = s_base + L.bars;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (8 + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (12 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = *a.n_edges;
  const int n_tiles = (E + kTileM - 1) / kTileM;
  const int M = a.M;

  // ---- one-time setup ----
  for (int t = threadIdx.x; t < a.n_gemms * kMaxWidth; t += kTcThreads) {
    const int g = t / kMaxWidth, n = t - g * kMaxWidth;
    s_bias[t] = (n < a.gN[g]) ? a.packed[a.gB[g] + n] : 0.f;
  }
  for (int t = threadIdx.x; t < kMaxWidth; t += kTcThreads) {
    s_wr[t] = (t < M) ? a.packed[a.msg_wr + t] : 0.f;
    s_wlast[t] = (t < a.gN[a.n_gemms - 1]) ? a.packed[a.coord_last_w + t] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), kEpiWarps); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) mbar_init(acc_full(s), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_base + L.tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_slot;

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t cb = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int g = 0; g < a.n_gemms; ++g) {
          const int nk = a.gK[g] / kChunkK;
          const uint32_t bytes = (uint32_t)a.gN[g] * 128u;
          const float* src = a.wtc + a.gW[g];
          for (int part = 0; part < 2 * nk; ++part, ++cb) {  // (chunk 0 hi, chunk 0 lo, chunk 1 hi, ...)
            const uint32_t s = cb % kBStages;
            mbar_wait(b_empty(s), ((cb / kBStages) & 1u) ^ 1u);
            mbar_arrive_expect_tx(b_full(s), bytes);
            bulk_g2s(s_base + L.b + s * kBStageBytes, src + (int64_t)part * a.gN[g] * kChunkK, bytes, b_full(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t ca = 0, cb = 0, gc = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int g = 0; g < a.n_gemms; ++g, ++gc) {
          const int nk = a.gK[g] / kChunkK;
          const uint32_t idesc = make_idesc_tf32(a.gN[g]);
          const uint32_t d_tmem = tmem_base + (gc & 1u) * kMaxWidth;
          for (int c = 0; c < nk; ++c, ++ca) {
            const uint32_t sa = ca % kAStages;
            mbar_wait(a_full(sa), (ca / kAStages) & 1u);
            const uint32_t a_hi = s_base + L.a + sa * kAStageBytes, a_lo = a_hi + kAPartBytes;
            uint32_t sb = cb % kBStages;
            mbar_wait(b_full(sb), (cb / kBStages) & 1u);
            tc_fence_after();
            uint32_t b_addr = s_base + L.b + sb * kBStageBytes;
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k) {  // A_hi B_hi + A_lo B_hi
              const uint64_t bd = make_desc_sw128(b_addr + k * kUmmaK * 4);
              umma_tf32(d_tmem, make_desc_sw128(a_hi + k * kUmmaK * 4), bd, idesc, (c | k) != 0);
              umma_tf32(d_tmem, make_desc_sw128(a_lo + k * kUmmaK * 4), bd, idesc, 1u);
            }
            umma_commit(b_empty(sb));
            ++cb;
            sb = cb % kBStages;
            mbar_wait(b_full(sb), (cb / kBStages) & 1u);
            tc_fence_after();
            b_addr = s_base + L.b + sb * kBStageBytes;
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k)  // A_hi B_lo
              umma_tf32(d_tmem, make_desc_sw128(a_hi + k * kUmmaK * 4), make_desc_sw128(b_addr + k * kUmmaK * 4), idesc, 1u);
            umma_commit(b_empty(sb));
            ++cb;
            umma_commit(a_empty(sa));
          }
          umma_commit(acc_full(gc & 1u));
        }
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;        // which 16-column half of a 32-column chunk
    const int row = q * 32 + lane;        // row of the tile = TMEM lane
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ca = 0, gc = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int e = tile * kTileM + row;
      const bool valid = e < E;
      const int src = valid ? a.edge_src[e] : -1;
      const int dst = valid ? a.edge_dst[e] : -1;
      // segments of equal source index inside this warp's 32 rows (edges are sorted by source)
      const int src_prev = __shfl_up_sync(0xffffffffu, src, 1);
      const int src_next = __shfl_down_sync(0xffffffffu, src, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || src != src_prev);
      const int seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
      const bool is_tail = (lane == 31) || (src != src_next);
      // coord2radial (models/egnn.py:230-238)
      float radial = 0.f;
      if (valid)
        for (int d = 0; d < a.D; ++d) {
          const float df = a.xcur[(int64_t)src * a.D + d] - a.xcur[(int64_t)dst * a.D + d];
          radial += df * df;
        }
      if (h == 0) s_dot[row] = 0.f;

      // ---- first activation: silu(P[src] + Q[dst] + radial * w_r), chunk by chunk ----
      const int nk0 = M / kChunkK;
      for (int c = 0; c < nk0; ++c, ++ca) {
        const int col0 = c * kChunkK + h * 16;
        float v[16];
        if (valid) {
          const float4* p4 = reinterpret_cast<const float4*>(a.P + (int64_t)src * M + col0);
          const float4* q4 = reinterpret_cast<const float4*>(a.Q + (int64_t)dst * M + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 pv = __ldg(p4 + j), qv = __ldg(q4 + j);
            v[4 * j + 0] = silu_fast(pv.x + qv.x + radial * s_wr[col0 + 4 * j + 0]);
            v[4 * j + 1] = silu_fast(pv.y + qv.y + radial * s_wr[col0 + 4 * j + 1]);
            v[4 * j + 2] = silu_fast(pv.z + qv.z + radial * s_wr[col0 + 4 * j + 2]);
            v[4 * j + 3] = silu_fast(pv.w + qv.w + radial * s_wr[col0 + 4 * j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (a.n_msg == 0) {  // no hidden message layer: this IS the message
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float sum = seg_scan(v[j], lane, seg_start);
            if (is_tail && valid) atomicAdd(a.agg + (int64_t)src * M + col0 + j, sum);
          }
        }
        const uint32_t sa = ca % kAStages;
        mbar_wait(a_empty(sa), ((ca / kAStages) & 1u) ^ 1u);
        uint8_t* hi = smem + L.a + sa * kAStageBytes;
        uint8_t* lo = hi + kAPartBytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 vh, vl;
          vh.x = tf32_rna(v[4 * j + 0]); vl.x = v[4 * j + 0] - vh.x;
          vh.y = tf32_rna(v[4 * j + 1]); vl.y = v[4 * j + 1] - vh.y;
          vh.z = tf32_rna(v[4 * j + 2]); vl.z = v[4 * j + 2] - vh.z;
          vh.w = tf32_rna(v[4 * j + 3]); vl.w = v[4 * j + 3] - vh.w;
          const uint32_t off = sw128_offset(row, h * 4 + j);
          *reinterpret_cast<float4*>(hi + off) = vh;
          *reinterpret_cast<float4*>(lo + off) = vl;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(sa));
      }

      // ---- the chain of Linear layers ----
      float dot = 0.f;
      for (int g = 0; g < a.n_gemms; ++g, ++gc) {
        const bool is_message = (g == a.n_msg - 1);
        const bool is_last = (g == a.n_gemms - 1);
        mbar_wait(acc_full(gc & 1u), (gc >> 1) & 1u);
        tc_fence_after();
        const int nchunks = a.gN[g] / kChunkK;
        const float* bias = s_bias + g * kMaxWidth;
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = c * kChunkK + h * 16;
          float v[16];
          tmem_ld16(lane_taddr + (gc & 1u) * kMaxWidth + col0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu_fast(v[j] + bias[col0 + j]);
          if (is_message) {  // agg[src] += m  (models/egnn.py:178-181)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float sum = seg_scan(valid ? v[j] : 0.f, lane, seg_start);
              if (is_tail && valid) atomicAdd(a.agg + (int64_t)src * M + col0 + j, sum);
            }
          }
          if (is_last) {  // last coordinate layer: Linear(Cx -> 1, no bias)
#pragma unroll
            for (int j = 0; j < 16; ++j) dot = fmaf(v[j], s_wlast[col0 + j], dot);
          } else {
            const uint32_t sa = ca % kAStages;
            mbar_wait(a_empty(sa), ((ca / kAStages) & 1u) ^ 1u);
            uint8_t* hi = smem + L.a + sa * kAStageBytes;
            uint8_t* lo = hi + kAPartBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 vh, vl;
              vh.x = tf32_rna(v[4 * j + 0]); vl.x = v[4 * j + 0] - vh.x;
              vh.y = tf32_rna(v[4 * j + 1]); vl.y = v[4 * j + 1] - vh.y;
              vh.z = tf32_rna(v[4 * j + 2]); vl.z = v[4 * j + 2] - vh.z;
              vh.w = tf32_rna(v[4 * j + 3]); vl.w = v[4 * j + 3] - vh.w;
              const uint32_t off = sw128_offset(row, h * 4 + j);
              *reinterpret_cast<float4*>(hi + off) = vh;
              *reinterpret_cast<float4*>(lo + off) = vl;
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(sa));
            ++ca;
          }
        }
        tc_fence_before();
      }

      // ---- coordinate update: xagg[src] += (x_src - x_dst) * phi_x(m)   (models/egnn.py:189-215) ----
      atomicAdd(&s_dot[row], dot);
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
      if (h == 0) {
        float phi = s_dot[row];
        if (a.tanh_out) phi = tanhf(phi);
        float norm = 1.0f;
        if (a.normalize) norm = tanhf(radial) / sqrtf(radial + 1e-16f);  // models/egnn.py:240-264, epsilon = 1e-8
        for (int d = 0; d < a.D; ++d) {
          float t = 0.f;
          if (valid) t = (a.xcur[(int64_t)src * a.D + d] - a.xcur[(int64_t)dst * a.D + d]) * norm * phi;
          const float sum = seg_scan(t, lane, seg_start);
          if (is_tail && valid) atomicAdd(a.xagg + (int64_t)src * a.D + d, sum);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- weight image: per GEMM, per K chunk, [hi | lo] blocks of N x 32 floats in the swizzled smem layout -------
__global__ void tc_pack_kernel(const float* __restrict__ wt /* [K][N] */, int K, int N, float* __restrict__ dst) {
  const int total = K * N;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int n = t / K, k = t - n * K;
    const float w = wt[(int64_t)k * N + n];
    uint32_t hi_bits;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi_bits) : "f"(w));
    const float hi = __uint_as_float(hi_bits);
    const float lo = w - hi;
    const int c = k / kChunkK, kk = k - c * kChunkK;
    const int64_t block = (int64_t)N * kChunkK;  // floats per hi or lo chunk
    const int64_t inner = (int64_t)(n >> 3) * 256 + (n & 7) * 32 + (((kk >> 2) ^ (n & 7)) << 2) + (kk & 3);
    dst[(2 * c) * block + inner] = hi;
    dst[(2 * c + 1) * block + inner] = lo;
  }
}

struct GemmPlan {
  int n;
  int K[kMaxGemms], N[kMaxGemms];
  int64_t wt[kMaxGemms], bias[kMaxGemms];  // offsets into the fp32 parameter image
  int64_t tc[kMaxGemms];                   // offsets into the layer's tensor-core image
  int64_t floats;
};

GemmPlan make_plan(const b200md_egnn_config& cfg, const EgnnLayerOffsets& Lo) {
  GemmPlan p{};
  const int M = cfg.message_hidden, Cx = cfg.coord_hidden;
  int g = 0;
  int64_t o = 0;
  auto add = [&](int K, int N, int64_t wt, int64_t b) {
    p.K[g] = K; p.N[g] = N; p.wt[g] = wt; p.bias[g] = b; p.tc[g] = o;
    o += 2ll * K * N;
    ++g;
  };
  for (int i = 0; i < cfg.message_n_hidden; ++i) add(M, M, Lo.msg_hidden_wt + (int64_t)i * M * M, Lo.msg_hidden_b + (int64_t)i * M);
  add(M, Cx, Lo.coord_w0_t, Lo.coord_b0);
  for (int i = 0; i < cfg.coord_n_hidden; ++i) add(Cx, Cx, Lo.coord_hidden_wt + (int64_t)i * Cx * Cx, Lo.coord_hidden_b + (int64_t)i * Cx);
  p.n = g;
  p.floats = o;
  return p;
}

}  // namespace

bool tc_edge_supported(const b200md_egnn_config& cfg) {
  const int M = cfg.message_hidden, Cx = cfg.coord_hidden;
  if (M % kChunkK || Cx % kChunkK || M > kMaxWidth || Cx > kMaxWidth || M < kChunkK || Cx < kChunkK) return false;
  if (cfg.flags & B200MD_EGNN_ATTENTION) return false;  // the gate needs the whole message row before the next layer
  if (cfg.message_n_hidden + 1 + cfg.coord_n_hidden > kMaxGemms) return false;
  return true;
}

int64_t tc_packed_floats(const b200md_egnn_config& cfg) {
  const int64_t M = cfg.message_hidden, Cx = cfg.coord_hidden;
  return 2 * (cfg.message_n_hidden * M * M + M * Cx + cfg.coord_n_hidden * Cx * Cx);
}

int tc_pack_layer(const b200md_egnn_config& cfg, const EgnnLayout& lay, int layer, float* packed, cudaStream_t st) {
  const GemmPlan plan = make_plan(cfg, lay.layer[layer]);
  float* base = packed + lay.tc_base + (int64_t)layer * lay.tc_layer_stride;
  for (int g = 0; g < plan.n; ++g) {
    { ProfileScope ps__(B200MD_KERNEL_OTHER, st); tc_pack_kernel<<<128, 256, 0, st>>>(packed + plan.wt[g], plan.K[g], plan.N[g], base + plan.tc[g]); }
    B200MD_CHECK_LAUNCH();
  }
  return B200MD_OK;
}

int tc_launch_edge_layer(const b200md_egnn_config& cfg, const EgnnLayout& lay, int layer, const float* packed,
                         const EgnnWorkspace& ws, cudaStream_t st) {
  const EgnnLayerOffsets& Lo = lay.layer[layer];
  const GemmPlan plan = make_plan(cfg, Lo);
  TcArgs a{};
  a.packed = packed;
  a.wtc = packed + lay.tc_base + (int64_t)layer * lay.tc_layer_stride;
  a.xcur = ws.xcur; a.P = ws.P; a.Q = ws.Q; a.agg = ws.agg; a.xagg = ws.xagg;
  a.edge_src = ws.edge_src; a.edge_dst = ws.edge_dst; a.n_edges = ws.n_edges;
  a.D = lay.D; a.M = lay.M;
  a.n_gemms = plan.n; a.n_msg = cfg.message_n_hidden;
  a.normalize = (cfg.flags & B200MD_EGNN_NORMALIZE) ? 1 : 0;
  a.tanh_out = (cfg.flags & B200MD_EGNN_TANH) ? 1 : 0;
  a.msg_wr = Lo.msg_wr; a.coord_last_w = Lo.coord_last_w;
  for (int g = 0; g < plan.n; ++g) { a.gK[g] = plan.K[g]; a.gN[g] = plan.N[g]; a.gW[g] = plan.tc[g]; a.gB[g] = plan.bias[g]; }
  const size_t smem = smem_layout(plan.n).total + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(egnn_edge_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_set = true;
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  { ProfileScope ps__(B200MD_KERNEL_EDGE, st); egnn_edge_tc_kernel<<<n_sm, kTcThreads, smem, st>>>(a); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

}  // namespace b200md
