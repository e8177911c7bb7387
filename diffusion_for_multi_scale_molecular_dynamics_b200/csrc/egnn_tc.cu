// EGNN edge kernel on the 5th-generation tensor cores (tcgen05 / TMEM), fp32-accurate through a hi/lo split.
//
// Reference arithmetic: models/egnn.py:136-160 (message MLP), :189-215 (coordinate MLP and update), :217-238
// (coord2radial), models/egnn_utils.py:11-70 (segment sum / mean).  Same fusion as the fp32 CUDA-core kernel of
// egnn.cu -- gather, distance featurisation, message MLP, coordinate MLP and the segmented reductions in one
// launch, no per-edge tensor ever reaches HBM -- but every Linear(K -> N) of the chain runs on the tensor cores
// (see tc_common.cuh for the pipeline and the two precision modes).
//
// One persistent CTA per SM, tile = 128 consecutive edges of the (src-sorted) edge list.  Warp roles:
//   warp 0      bulk-copy producer of the weight chunks (L2-resident image -> 4-stage shared-memory ring);
//   warp 1      TMEM allocator and single-thread MMA issuer;
//   warps 2-9   "epilogue" warps: build the first activation from P[src] + Q[dst] + r w_r, and between two Linear
//               layers read the accumulator (tcgen05.ld), add the bias, apply SiLU, split into hi/lo and write the
//               next layer's A operand chunk by chunk, so that the MMA of layer g+1 on chunk c overlaps the epilogue
//               of layer g on chunk c+1.  They also do the segmented warp reductions (messages -> agg, coordinate
//               updates -> xagg) with one atomic per (segment, column).
#include <math_constants.h>

#include "egnn.cuh"
#include "tc_common.cuh"

namespace b200md {

using namespace tc;

namespace {

struct EdgeTcArgs {
  Program prog;
  const float *xcur, *P, *Q;
  float *agg, *xagg;
  const int32_t *edge_src, *edge_dst, *n_edges;
  int D, M;
  int n_msg;             // the first n_msg GEMMs belong to the message MLP
  int normalize, tanh_out;
  int64_t msg_wr, coord_last_w;
  int32_t* status;
};

struct EdgeSmem {
  uint32_t bias, wr, wlast, dot, total;
};
__host__ __device__ inline EdgeSmem edge_smem(int n_gemms) {
  EdgeSmem L;
  uint32_t o = pipe_smem().end;
  L.bias = o; o += (uint32_t)n_gemms * kMaxN * 4;
  L.wr = o; o += kMaxN * 4;
  L.wlast = o; o += kMaxN * 4;
  L.dot = o; o += kTileM * 4;
  L.total = o;
  return L;
}

template <int PREC>
__global__ void __launch_bounds__(kThreads, 1) egnn_edge_tc_kernel(const EdgeTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Program& p = a.prog;
  const EdgeSmem L = edge_smem(p.n_gemms);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  float* s_wr = reinterpret_cast<float*>(smem + L.wr);
  float* s_wlast = reinterpret_cast<float*>(smem + L.wlast);
  float* s_dot = reinterpret_cast<float*>(smem + L.dot);

  constexpr int kChunk = Mode<PREC>::kChunk, kVals = Mode<PREC>::kVals;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = *a.n_edges;
  const int n_tiles = (E + kTileM - 1) / kTileM;
  const int M = a.M;

  for (int t = threadIdx.x; t < p.n_gemms * kMaxN; t += kThreads) {
    const int g = t / kMaxN, n = t - g * kMaxN;
    s_bias[t] = (n < p.gN[g] && p.gB[g] >= 0) ? p.packed[p.gB[g] + n] : 0.f;
  }
  for (int t = threadIdx.x; t < kMaxN; t += kThreads) {
    s_wr[t] = (t < M) ? p.packed[a.msg_wr + t] : 0.f;
    s_wlast[t] = (t < p.gN[p.n_gemms - 1]) ? p.packed[a.coord_last_w + t] : 0.f;
  }
  Pipe<PREC> pipe;
  pipe.setup(smem);
  const TileSchedule sched = make_schedule(n_tiles);

  if (warp == 0) {
    if (lane == 0) pipe.run_producer(p, sched.n_iterations);
  } else if (warp == 1) {
    pipe.run_mma(p, sched.n_iterations);
  } else {
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int hh = (warp - 2) >> 2;       // which quarter of a chunk's columns
    float vmax = 0.f;                     // largest activation handed to the tensor cores (fp16 range guard)
    const int row = q * 32 + lane;        // row of the tile = TMEM lane
    const uint32_t lane_taddr = pipe.tmem_base + ((uint32_t)(q * 32) << 16);
    for (int it = 0; it < sched.n_iterations; ++it) {
      const int tile = sched.tile(it);
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
      if (hh == 0) s_dot[row] = 0.f;

      // ---- first activation: silu(P[src] + Q[dst] + radial * w_r), chunk by chunk ----
      const int nk0 = M / kChunk;
      for (int c = 0; c < nk0; ++c) {
        const int col0 = c * kChunk + hh * kVals;
        float v[kVals];
        if (valid) {
          const float4* p4 = reinterpret_cast<const float4*>(a.P + (int64_t)src * M + col0);
          const float4* q4 = reinterpret_cast<const float4*>(a.Q + (int64_t)dst * M + col0);
#pragma unroll
          for (int j = 0; j < kVals / 4; ++j) {
            const float4 pv = __ldg(p4 + j), qv = __ldg(q4 + j);
            v[4 * j + 0] = silu_fast(pv.x + qv.x + radial * s_wr[col0 + 4 * j + 0]);
            v[4 * j + 1] = silu_fast(pv.y + qv.y + radial * s_wr[col0 + 4 * j + 1]);
            v[4 * j + 2] = silu_fast(pv.z + qv.z + radial * s_wr[col0 + 4 * j + 2]);
            v[4 * j + 3] = silu_fast(pv.w + qv.w + radial * s_wr[col0 + 4 * j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < kVals; ++j) v[j] = 0.f;
        }
        if (a.n_msg == 0) {  // no hidden message layer: this IS the message
#pragma unroll
          for (int j = 0; j < kVals; ++j) {
            const float sum = seg_scan(v[j], lane, seg_start);
            if (is_tail && valid) atomicAdd(a.agg + (int64_t)src * M + col0 + j, sum);
          }
        }
        vmax = fmaxf(vmax, pipe.push_chunk(row, hh, lane, v));
      }

      // ---- the chain of Linear layers ----
      float dot = 0.f;
      for (int g = 0; g < p.n_gemms; ++g) {
        const bool is_message = (g == a.n_msg - 1);
        const bool is_last = (g == p.n_gemms - 1);
        const uint32_t acc_col = pipe.wait_accumulator();
        const int nchunks = p.gN[g] / kChunk;
        const float* bias = s_bias + g * kMaxN;
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = c * kChunk + hh * kVals;
          float v[kVals];
          tmem_ld<kVals>(lane_taddr + acc_col + col0, v);
#pragma unroll
          for (int j = 0; j < kVals; ++j) v[j] = silu_fast(v[j] + bias[col0 + j]);
          if (is_message) {  // agg[src] += m  (models/egnn.py:178-181)
#pragma unroll
            for (int j = 0; j < kVals; ++j) {
              const float sum = seg_scan(valid ? v[j] : 0.f, lane, seg_start);
              if (is_tail && valid) atomicAdd(a.agg + (int64_t)src * M + col0 + j, sum);
            }
          }
          if (is_last) {  // last coordinate layer: Linear(Cx -> 1, no bias)
#pragma unroll
            for (int j = 0; j < kVals; ++j) dot = fmaf(v[j], s_wlast[col0 + j], dot);
          } else {
            vmax = fmaxf(vmax, pipe.push_chunk(row, hh, lane, v));
          }
        }
        tc_fence_before();
      }

      // ---- coordinate update: xagg[src] += (x_src - x_dst) * phi_x(m)   (models/egnn.py:189-215) ----
      atomicAdd(&s_dot[row], dot);
      epilogue_barrier();
      if (hh == 0) {
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
      epilogue_barrier();
    }
    if (PREC == kF16x3 && vmax > kF16Safe && a.status) atomicOr(a.status, (int)B200MD_STATUS_TC_RANGE);
  }
  pipe.teardown();
}

// =====================================================================================================================
// 3xFP16 with the A operand in TENSOR MEMORY (tcgen05.mma ".ts" form): the epilogue converts an accumulator IN PLACE into
// the next layer's A operand.  Accumulator columns [64c + 16j, +16) of a row (fp32) become, for K chunk c and K step j,
// 8 columns of packed fp16 pairs `hi` followed by 8 columns `lo` -- exactly the (lane = row, column = k / 2) layout
// tcgen05.mma reads for a K = 16 step -- so no activation ever touches shared memory: no swizzled stores, no
// generic->async proxy fence, no A ring (its 64 KB go to a 12-stage weight ring), and the tensor pipe reads only B from
// shared memory.  Because A(g+1) stays in TMEM for the whole GEMM, the MMA order is N-half-major: the first half of the
// output columns over all K chunks (as the epilogue delivers them), then the second half back to back; the epilogue of
// layer g+1 starts on the first half while the tensor pipe still computes the second -- the accumulator-complete ->
// first-chunk bubble of the shared-memory pipeline (one whole K chunk of MMA) disappears.
//   GEMM index n (running over tiles): D = buffer n & 1, A = buffer (n & 1) ^ 1 (2 x 256 columns = all of TMEM).
constexpr int kTsBStageBytes = 128 * kRowBytes;   // one weight part of a K chunk for HALF of the output columns (<= 16 KB)
constexpr int kTsBStages = 12;
constexpr int kTsSlots = 4;                       // K-chunk slots of the "A ready" barriers (x 4 K steps each)

struct TsSmem {
  uint32_t b, bars, tmem_slot, bias, wr, wlast, dot, total;
};
__host__ __device__ inline TsSmem ts_smem(int n_gemms) {
  TsSmem L;
  uint32_t o = 0;
  L.b = o; o += kTsBStages * kTsBStageBytes;
  L.bars = o; o += (2 * kTsBStages + kTsSlots * 4 + 4) * 8;
  L.tmem_slot = o; o += 16;
  L.bias = o; o += (uint32_t)n_gemms * kMaxN * 4;
  L.wr = o; o += kMaxN * 4;
  L.wlast = o; o += kMaxN * 4;
  L.dot = o; o += 2 * 4 * kTileM * 4;   // [tile parity][K step warp][row]
  L.total = o;
  return L;
}

// D[tmem] (+)= A[tmem] * B[smem desc]^T
// `b_desc_lo`: low word of the K-major SWIZZLE_128B descriptor (desc_lo_sw128); the high word is a constant.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t idesc,
                                            uint32_t accumulate) {
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// two fp32 -> packed f16x2 (first argument in the LOW half); out-of-range values become +-inf (no saturation): they
// poison the chain and are caught by the finite check on the tile's result
__device__ __forceinline__ uint32_t pack_f16x2_inf(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// shared-memory load by window address (keeps the access an LDS: pointers derived from the aligned dynamic-smem base
// lose their address space and become generic loads)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void red_add_v4(float* addr, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

struct TsPipe {
  uint32_t s_base, bar0, tmem_base;
  __device__ __forceinline__ uint32_t b_full(uint32_t s) const { return bar0 + 8u * s; }
  __device__ __forceinline__ uint32_t b_empty(uint32_t s) const { return bar0 + 8u * (kTsBStages + s); }
  __device__ __forceinline__ uint32_t a_full(uint32_t slot, uint32_t j) const { return bar0 + 8u * (2 * kTsBStages + slot * 4 + j); }
  __device__ __forceinline__ uint32_t acc_full(uint32_t buf, uint32_t half) const {
    return bar0 + 8u * (2 * kTsBStages + kTsSlots * 4 + buf * 2 + half);
  }
};

__global__ void __launch_bounds__(kThreads, 1) egnn_edge_ts_kernel(const EdgeTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Program& p = a.prog;
  const TsSmem L = ts_smem(p.n_gemms);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  float* s_wr = reinterpret_cast<float*>(smem + L.wr);
  float* s_wlast = reinterpret_cast<float*>(smem + L.wlast);
  float* s_dot = reinterpret_cast<float*>(smem + L.dot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = *a.n_edges;
  const int n_tiles = (E + kTileM - 1) / kTileM;
  const int M = a.M;

  for (int t = threadIdx.x; t < p.n_gemms * kMaxN; t += kThreads) {
    const int g = t / kMaxN, n = t - g * kMaxN;
    s_bias[t] = (n < p.gN[g] && p.gB[g] >= 0) ? p.packed[p.gB[g] + n] : 0.f;
  }
  for (int t = threadIdx.x; t < kMaxN; t += kThreads) {
    s_wr[t] = (t < M) ? p.packed[a.msg_wr + t] : 0.f;
    s_wlast[t] = (t < p.gN[p.n_gemms - 1]) ? p.packed[a.coord_last_w + t] : 0.f;
  }
  TsPipe pipe;
  pipe.s_base = smem_u32(smem);
  pipe.bar0 = pipe.s_base + L.bars;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTsBStages; ++s) { mbar_init(pipe.b_full(s), 1); mbar_init(pipe.b_empty(s), 1); }
    for (int s = 0; s < kTsSlots; ++s)
      for (int j = 0; j < 4; ++j) mbar_init(pipe.a_full(s, j), 4);   // the 4 lane-quarter warps of K step j
    for (int s = 0; s < 4; ++s) mbar_init(pipe.acc_full(s >> 1, s & 1), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(pipe.s_base + L.tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pipe.tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);

  const int n_iterations = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t cb = 0;
      for (int it = 0; it < n_iterations; ++it)
        for (int g = 0; g < p.n_gemms; ++g) {
          const int nk = p.gK[g] / 64, nh = p.gN[g] / 2;
          const uint32_t bytes = (uint32_t)nh * kRowBytes;
          const int64_t part_floats = (int64_t)p.gN[g] * (kRowBytes / 4);
          const float* src = p.wtc + p.gW[g] + f16_image_offset(p.gK[g], p.gN[g]);
          for (int h = 0; h < 2; ++h)
            for (int part = 0; part < 2 * nk; ++part, ++cb) {   // chunk 0 hi, chunk 0 lo, chunk 1 hi, ...
              const uint32_t s = cb % kTsBStages;
              mbar_wait(pipe.b_empty(s), ((cb / kTsBStages) & 1u) ^ 1u);
              mbar_arrive_expect_tx(pipe.b_full(s), bytes);
              bulk_g2s(pipe.s_base + L.b + s * kTsBStageBytes, src + (int64_t)part * part_floats + (int64_t)h * nh * (kRowBytes / 4),
                       bytes, pipe.b_full(s));
            }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The WHOLE warp walks the loop (converged), so every address and descriptor lives in uniform registers; one
    // elected lane issues.  (With only lane 0 inside the loop the compiler wraps each tcgen05.mma in an
    // ELECT / R2UR / branch "waterfall": ~40 instructions per MMA, and the issuing thread -- not the tensor pipe --
    // paces the kernel.)
    {
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, pipe.tmem_base, 0);
      const uint32_t b_ring = pipe.s_base + L.b;
      const bool leader = elect_one();
      uint32_t cb = 0, mca = 0, mgc = 0;
      for (int it = 0; it < n_iterations; ++it)
        for (int g = 0; g < p.n_gemms; ++g, ++mgc) {
          const int nk = p.gK[g] / 64, nh = p.gN[g] / 2;
          const uint32_t idesc = make_idesc(nh, 0u);
          const uint32_t d_col = (mgc & 1u) * kMaxN, a_col = d_col ^ kMaxN;
          for (int h = 0; h < 2; ++h) {
            const uint32_t d_tmem = tmem_base + d_col + (uint32_t)(h * nh);
            for (int c = 0; c < nk; ++c) {
              const uint32_t s1 = cb % kTsBStages, s2 = (cb + 1) % kTsBStages;
              mbar_wait(pipe.b_full(s1), (cb / kTsBStages) & 1u);
              mbar_wait(pipe.b_full(s2), ((cb + 1) / kTsBStages) & 1u);
              const uint32_t bd1 = desc_lo_sw128(b_ring + s1 * kTsBStageBytes), bd2 = desc_lo_sw128(b_ring + s2 * kTsBStageBytes);
              const uint32_t chunk = mca + (uint32_t)c;
              const uint32_t a_chunk = tmem_base + a_col + (uint32_t)(c * 64);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (h == 0) mbar_wait(pipe.a_full(chunk % kTsSlots, j), (chunk / kTsSlots) & 1u);
                tc_fence_after();
                if (leader) {
                  umma_f16_ts(d_tmem, a_chunk + j * 16, bd1 + 2 * j, idesc, (c | j) != 0);   // A_1 B_1
                  umma_f16_ts(d_tmem, a_chunk + j * 16 + 8, bd1 + 2 * j, idesc, 1u);         // A_2 B_1
                  umma_f16_ts(d_tmem, a_chunk + j * 16, bd2 + 2 * j, idesc, 1u);             // A_1 B_2
                }
              }
              if (leader) {
                umma_commit(pipe.b_empty(s1));
                umma_commit(pipe.b_empty(s2));
              }
              __syncwarp();
              cb += 2;
            }
            if (leader) umma_commit(pipe.acc_full(mgc & 1u, h));
            __syncwarp();
          }
          mca += nk;
        }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int hh = (warp - 2) >> 2;       // K step of a chunk (16 columns) this warp converts
    const int row = q * 32 + lane;
    const uint32_t lane_taddr = pipe.tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sa_bias = pipe.s_base + L.bias, sa_wr = pipe.s_base + L.wr, sa_wlast = pipe.s_base + L.wlast;
    uint32_t ca = 0, gc = 0;
    bool poisoned = false;

    // convert 16 activations to [hi x 8 | lo x 8] packed fp16 pairs, store them at `col` of this lane, publish K step hh
    auto push = [&](uint32_t col, const float (&v)[16]) {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x0 = v[2 * i], x1 = v[2 * i + 1];
        w[i] = pack_f16x2_inf(x0, x1);
        const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        w[8 + i] = pack_f16x2_inf(x0 - back.x, x1 - back.y);
      }
      tmem_st16(lane_taddr + col, w);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pipe.a_full(ca % kTsSlots, hh));
      ++ca;
    };

    // per-tile state of this thread's edge (row of the tile)
    struct Edge {
      int src, dst, seg_start;
      bool valid, is_tail;
      float radial;
    };
    auto load_edge = [&](int it) {
      Edge ed;
      const int e = ((int)blockIdx.x + it * (int)gridDim.x) * kTileM + row;
      ed.valid = e < E;
      ed.src = ed.valid ? a.edge_src[e] : -1;
      ed.dst = ed.valid ? a.edge_dst[e] : -1;
      // segments of equal source index inside this warp's 32 rows (edges are sorted by source)
      const int src_prev = __shfl_up_sync(0xffffffffu, ed.src, 1);
      const int src_next = __shfl_down_sync(0xffffffffu, ed.src, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || ed.src != src_prev);
      ed.seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
      ed.is_tail = (lane == 31) || (ed.src != src_next);
      ed.radial = 0.f;   // coord2radial (models/egnn.py:230-238)
      if (ed.valid)
        for (int d = 0; d < a.D; ++d) {
          const float df = a.xcur[(int64_t)ed.src * a.D + d] - a.xcur[(int64_t)ed.dst * a.D + d];
          ed.radial += df * df;
        }
      return ed;
    };
    // 16 columns at once: inclusive segmented scans over the warp's rows (interleaved, so the shuffle latencies
    // overlap), then one vector reduction per segment tail:  target[src, col0 .. col0 + 16) += sum over the segment
    auto segment_add16 = [&](const Edge& ed, const float (&v)[16], float* target_row) {
      float sum[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) sum[j] = ed.valid ? v[j] : 0.f;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        float t[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = __shfl_up_sync(0xffffffffu, sum[j], off);
        const bool take = lane - off >= ed.seg_start;
#pragma unroll
        for (int j = 0; j < 16; ++j) sum[j] += take ? t[j] : 0.f;
      }
      if (ed.is_tail && ed.valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) red_add_v4(target_row + 4 * j, sum[4 * j], sum[4 * j + 1], sum[4 * j + 2], sum[4 * j + 3]);
      }
    };
    // chunk c of the first activation of a tile: silu(P[src] + Q[dst] + radial * w_r) -> A operand at a_col
    auto prefetch_first = [&](const Edge& ed, int c) {
      if (ed.valid) {
        prefetch_l1(a.P + (int64_t)ed.src * M + c * 64 + hh * 16);
        prefetch_l1(a.Q + (int64_t)ed.dst * M + c * 64 + hh * 16);
      }
    };
    auto first_activation_chunk = [&](const Edge& ed, int c, uint32_t a_col) {
      const int col0 = c * 64 + hh * 16;
      float v[16];
      if (ed.valid) {
        const float4* p4 = reinterpret_cast<const float4*>(a.P + (int64_t)ed.src * M + col0);
        const float4* q4 = reinterpret_cast<const float4*>(a.Q + (int64_t)ed.dst * M + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 pv = __ldg(p4 + j), qv = __ldg(q4 + j), wv = lds128(sa_wr + (uint32_t)(col0 + 4 * j) * 4u);
          v[4 * j + 0] = silu_fast(pv.x + qv.x + ed.radial * wv.x);
          v[4 * j + 1] = silu_fast(pv.y + qv.y + ed.radial * wv.y);
          v[4 * j + 2] = silu_fast(pv.z + qv.z + ed.radial * wv.z);
          v[4 * j + 3] = silu_fast(pv.w + qv.w + ed.radial * wv.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (a.n_msg == 0) segment_add16(ed, v, a.agg + (int64_t)ed.src * M + col0);  // no hidden layer: this IS the message
      push(a_col + (uint32_t)col0, v);
    };

    const int nk0 = M / 64;
    const int n_last = p.gN[p.n_gemms - 1] / 64;   // chunks of the last accumulator
    Edge cur{};
    if (n_iterations > 0) {
      cur = load_edge(0);
      const uint32_t a_col = ((gc & 1u) ^ 1u) * kMaxN;
      prefetch_first(cur, 0);
      for (int c = 0; c < nk0; ++c) {
        if (c + 1 < nk0) prefetch_first(cur, c + 1);
        first_activation_chunk(cur, c, a_col);
      }
    }
    for (int it = 0; it < n_iterations; ++it) {
      // ---- the chain of Linear layers ----
      float dot = 0.f;
      Edge next{};
      const bool has_next = it + 1 < n_iterations;
      for (int g = 0; g < p.n_gemms; ++g, ++gc) {
        const bool is_message = (g == a.n_msg - 1);
        const bool is_last = (g == p.n_gemms - 1);
        const uint32_t d_col = (gc & 1u) * kMaxN;
        const int nchunks = p.gN[g] / 64, nh = p.gN[g] / 2;
        const uint32_t bias = sa_bias + (uint32_t)(g * kMaxN) * 4u;
        if (is_last && has_next) {   // the next tile's edge and the first lines of its P / Q rows, while the last MMAs run
          next = load_edge(it + 1);
          prefetch_first(next, 0);
        }
        int waited = -1;
        for (int c = 0; c < nchunks; ++c) {
          const int need = (c * 64 + 64 <= nh) ? 0 : 1;
          if (need > waited) {
            mbar_wait(pipe.acc_full(gc & 1u, need), (gc >> 1) & 1u);
            tc_fence_after();
            waited = need;
          }
          const int col0 = c * 64 + hh * 16;
          float v[16];
          tmem_ld16(lane_taddr + d_col + col0, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b = lds128(bias + (uint32_t)(col0 + 4 * j) * 4u);
            v[4 * j + 0] = silu_fast(v[4 * j + 0] + b.x);
            v[4 * j + 1] = silu_fast(v[4 * j + 1] + b.y);
            v[4 * j + 2] = silu_fast(v[4 * j + 2] + b.z);
            v[4 * j + 3] = silu_fast(v[4 * j + 3] + b.w);
          }
          if (is_message) segment_add16(cur, v, a.agg + (int64_t)cur.src * M + col0);  // agg[src] += m (models/egnn.py:178-181)
          if (is_last) {  // last coordinate layer: Linear(Cx -> 1, no bias)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 w = lds128(sa_wlast + (uint32_t)(col0 + 4 * j) * 4u);
              dot = fmaf(v[4 * j + 0], w.x, dot);
              dot = fmaf(v[4 * j + 1], w.y, dot);
              dot = fmaf(v[4 * j + 2], w.z, dot);
              dot = fmaf(v[4 * j + 3], w.w, dot);
            }
            // these accumulator columns are free now: the next tile's first activation goes there IN PLACE, so its
            // first GEMM starts while this tile is being finished
            if (has_next && c < nk0) {
              if (c + 1 < nk0) prefetch_first(next, c + 1);
              first_activation_chunk(next, c, d_col);
            }
          } else {
            push(d_col + (uint32_t)col0, v);
          }
        }
        if (is_last && has_next)
          for (int c = n_last; c < nk0; ++c) {   // message width > coordinate width: columns no accumulator used
            if (c + 1 < nk0) prefetch_first(next, c + 1);
            first_activation_chunk(next, c, d_col);
          }
      }
      // an fp16 overflow anywhere in the chain has turned into inf / NaN by now (see pack_f16x2_inf)
      if (cur.valid && !(fabsf(dot) <= CUDART_MAX_NORMAL_F)) poisoned = true;

      // ---- coordinate update: xagg[src] += (x_src - x_dst) * phi_x(m)   (models/egnn.py:189-215) ----
      float* dot_slot = s_dot + (it & 1) * 4 * kTileM;   // [parity][hh][row]: one barrier per tile is enough
      dot_slot[hh * kTileM + row] = dot;
      epilogue_barrier();
      {
        float phi = (dot_slot[row] + dot_slot[kTileM + row]) + (dot_slot[2 * kTileM + row] + dot_slot[3 * kTileM + row]);
        if (a.tanh_out) phi = tanhf(phi);
        float norm = 1.0f;
        if (a.normalize) norm = tanhf(cur.radial) / sqrtf(cur.radial + 1e-16f);  // models/egnn.py:240-264, epsilon = 1e-8
        for (int d = hh; d < a.D; d += kEpiParts) {   // the coordinate columns are dealt over the four warps of a row
          float t = 0.f;
          if (cur.valid) t = (a.xcur[(int64_t)cur.src * a.D + d] - a.xcur[(int64_t)cur.dst * a.D + d]) * norm * phi;
          const float sum = seg_scan(t, lane, cur.seg_start);
          if (cur.is_tail && cur.valid) atomicAdd(a.xagg + (int64_t)cur.src * a.D + d, sum);
        }
      }
      cur = next;
    }
    if (poisoned && a.status) atomicOr(a.status, (int)B200MD_STATUS_TC_RANGE);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(pipe.tmem_base, kTmemCols);
  }
}

// ---- weight image (tc_common.cuh): TF32 parts [w_1 | w_2] per 32-K chunk, then fp16 parts [w_1 | w_2] per 64-K chunk
__global__ void tc_pack_kernel(const float* __restrict__ wt /* [K][N] */, int K, int N, float* __restrict__ dst) {
  const int total = K * N;
  __half* dst16 = reinterpret_cast<__half*>(dst + f16_image_offset(K, N));
  const int64_t part = (int64_t)N * 32;  // floats per part (N rows of 128 bytes)
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int n = t / K, k = t - n * K;
    const float w = wt[(int64_t)k * N + n];
    const int sw = n & 7;
    {  // TF32 image: 4-byte elements, 32 per row, 4 per 16-byte unit
      const float w1 = tf32_rna(w);
      const int c = k >> 5, kk = k & 31;
      const int64_t inner = (int64_t)(n >> 3) * 256 + sw * 32 + (((kk >> 2) ^ sw) << 2) + (kk & 3);
      dst[((int64_t)c * 2 + 0) * part + inner] = w1;
      dst[((int64_t)c * 2 + 1) * part + inner] = w - w1;
    }
    if ((K & 63) == 0) {  // fp16 image: 2-byte elements, 64 per row, 8 per 16-byte unit
      const __half h1 = __float2half_rn(w);
      const __half h2 = __float2half_rn(w - __half2float(h1));
      const int c = k >> 6, kk = k & 63;
      const int64_t inner = (int64_t)(n >> 3) * 512 + sw * 64 + (((kk >> 3) ^ sw) << 3) + (kk & 7);
      dst16[((int64_t)c * 2 + 0) * part * 2 + inner] = h1;
      dst16[((int64_t)c * 2 + 1) * part * 2 + inner] = h2;
    }
  }
}

struct EdgePlan {
  int n;
  int K[kMaxGemms], N[kMaxGemms];
  int64_t wt[kMaxGemms], bias[kMaxGemms];  // offsets into the fp32 parameter image
  int64_t tc[kMaxGemms];                   // offsets into the layer's tensor-core image
  int64_t floats;
};

EdgePlan make_edge_plan(const b200md_egnn_config& cfg, const EgnnLayerOffsets& Lo) {
  EdgePlan p{};
  const int M = cfg.message_hidden, Cx = cfg.coord_hidden;
  int g = 0;
  int64_t o = 0;
  auto add = [&](int K, int N, int64_t wt, int64_t b) {
    p.K[g] = K; p.N[g] = N; p.wt[g] = wt; p.bias[g] = b; p.tc[g] = o;
    o += weight_image_floats(K, N);
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

namespace tc {
int pack_weight_image(const float* wt, int K, int N, float* dst, cudaStream_t st) {
  { ProfileScope ps__(B200MD_KERNEL_OTHER, st); tc_pack_kernel<<<128, 256, 0, st>>>(wt, K, N, dst); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}
}  // namespace tc

int tc_sm_count() {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  return n_sm;
}

// `precision`: the K chunk (and so the width granularity) is 32 for kTf32x3 and 64 for kF16x3
bool tc_edge_supported(const b200md_egnn_config& cfg, int precision) {
  const int M = cfg.message_hidden, Cx = cfg.coord_hidden;
  const int chunk = precision == kTf32x3 ? 32 : 64;
  if (M % chunk || Cx % chunk || M > kMaxN || Cx > kMaxN || M < chunk || Cx < chunk) return false;
  if (cfg.flags & B200MD_EGNN_ATTENTION) return false;  // the gate needs the whole message row before the next layer
  if (cfg.message_n_hidden + 1 + cfg.coord_n_hidden > kMaxGemms) return false;
  return true;
}

int64_t tc_edge_packed_floats(const b200md_egnn_config& cfg) {
  const int64_t M = cfg.message_hidden, Cx = cfg.coord_hidden;
  return 3 * (cfg.message_n_hidden * M * M + M * Cx + cfg.coord_n_hidden * Cx * Cx);
}

int tc_edge_pack_layer(const b200md_egnn_config& cfg, const EgnnLayout& lay, int layer, float* packed, cudaStream_t st) {
  const EdgePlan plan = make_edge_plan(cfg, lay.layer[layer]);
  float* base = packed + lay.tc_edge_base + (int64_t)layer * lay.tc_edge_stride;
  for (int g = 0; g < plan.n; ++g) {
    int rc = pack_weight_image(packed + plan.wt[g], plan.K[g], plan.N[g], base + plan.tc[g], st);
    if (rc != B200MD_OK) return rc;
  }
  return B200MD_OK;
}

int tc_launch_edge_layer(const b200md_egnn_config& cfg, const EgnnLayout& lay, int layer, const float* packed,
                         const EgnnWorkspace& ws, int32_t* status, int precision, cudaStream_t st) {
  const EgnnLayerOffsets& Lo = lay.layer[layer];
  const EdgePlan plan = make_edge_plan(cfg, Lo);
  EdgeTcArgs a{};
  a.prog.packed = packed;
  a.prog.wtc = packed + lay.tc_edge_base + (int64_t)layer * lay.tc_edge_stride;
  a.prog.n_gemms = plan.n;
  for (int g = 0; g < plan.n; ++g) { a.prog.gK[g] = plan.K[g]; a.prog.gN[g] = plan.N[g]; a.prog.gW[g] = plan.tc[g]; a.prog.gB[g] = plan.bias[g]; }
  a.xcur = ws.xcur; a.P = ws.P; a.Q = ws.Q; a.agg = ws.agg; a.xagg = ws.xagg;
  a.edge_src = ws.edge_src; a.edge_dst = ws.edge_dst; a.n_edges = ws.n_edges;
  a.D = lay.D; a.M = lay.M;
  a.n_msg = cfg.message_n_hidden;
  a.normalize = (cfg.flags & B200MD_EGNN_NORMALIZE) ? 1 : 0;
  a.tanh_out = (cfg.flags & B200MD_EGNN_TANH) ? 1 : 0;
  a.msg_wr = Lo.msg_wr; a.coord_last_w = Lo.coord_last_w;
  a.status = status;
  const bool ts = (precision == kF16x3Ts);
  const size_t smem = (ts ? ts_smem(plan.n).total : edge_smem(plan.n).total) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e0 = cudaFuncSetAttribute(egnn_edge_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e0 != cudaSuccess) return set_cuda_error(e0);
    cudaError_t e1 = cudaFuncSetAttribute(egnn_edge_tc_kernel<kTf32x3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaError_t e2 = cudaFuncSetAttribute(egnn_edge_tc_kernel<kF16x3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e1 != cudaSuccess) return set_cuda_error(e1);
    if (e2 != cudaSuccess) return set_cuda_error(e2);
    attr_set = true;
  }
  if (ts) {
    ProfileScope ps__(B200MD_KERNEL_EDGE, st);
    egnn_edge_ts_kernel<<<(unsigned)tc_sm_count(), kThreads, smem, st>>>(a);
  } else {
    ProfileScope ps__(B200MD_KERNEL_EDGE, st);
    cudaError_t e = launch_clustered(precision == kF16x3 ? egnn_edge_tc_kernel<kF16x3> : egnn_edge_tc_kernel<kTf32x3>,
                                     (unsigned)tc_sm_count(), (unsigned)tc_sm_count(), smem, st, a);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

}  // namespace b200md
