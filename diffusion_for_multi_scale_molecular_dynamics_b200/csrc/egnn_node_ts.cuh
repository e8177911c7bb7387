// EGNN node kernel with the activations in TENSOR MEMORY (the node-side sibling of egnn_edge_ts.cuh).  Included by
// egnn_node_tc.cu inside its anonymous namespace.
//
// Reference arithmetic: models/egnn.py:162-187 (node_model), :212-215 (coord += agg), :376-379 (embedding, read-out),
// models/score_networks/egnn_score_network.py:165-224,276-281, models/score_networks/score_network.py:183-185.
//
// The node MLP's first Linear acts on [h | agg] (K = H + M = 512): too wide to sit beside a 256-column accumulator.
// It is split like the first message Linear already is: U = h W0a^T + b0 is produced by the call that produces h (next
// to P and Q, same A operand), and the update call computes agg W0b^T and adds U in the epilogue.  Every GEMM is then
// K, N <= 256 and the program of a tile is
//     chain : A0 = agg / deg -> [+U, SiLU] -> n x [Linear, SiLU] -> [Linear + bias + residual = h', stored]     (update)
//     out   : P = h' Wa^T + b, Q = h' Wb^T, U = h' W0a^T + b0  (three GEMMs sharing the A operand h')        (a layer follows)
// Chain GEMMs alternate the two 256-column TMEM buffers (accumulator converted IN PLACE into the next A operand, as in
// the edge kernel); the three output GEMMs keep h' in one buffer and reuse the other as accumulator, so each waits for
// the epilogue to have drained the previous one ("d_free").
//
// Warps (10; ts_common.cuh): 0 = weight producer, 1 = MMA issuer (converged, elected lane), 2-9 = epilogue (two per TMEM
// lane quarter, 32 columns per thread and chunk).
#pragma once

// The node kernel runs the 12-warp layout of ts_common.cuh (control warpgroup shrunk to 56 registers, the two epilogue
// warpgroups grown to 224 with setmaxnreg): with 168 registers its epilogue spills 164 bytes per thread, and with 227 KB of
// shared memory in use there is hardly any L1 for local memory -- A/B on one box 11.2 -> 9.5 ms per step.  (The edge
// kernel does not spill and is 0.3 ms per launch slower in that layout: it keeps the 10 warps.)
#ifndef B200MD_NODE_TS_WARPGROUPS
#define B200MD_NODE_TS_WARPGROUPS 1
#endif
#ifndef B200MD_TS_WARPGROUPS
#define B200MD_TS_WARPGROUPS B200MD_NODE_TS_WARPGROUPS
#endif
#include "ts_common.cuh"

#ifndef B200MD_NODE_TS_EXACT_DIV
#define B200MD_NODE_TS_EXACT_DIV 0
#endif
// reciprocals of the SiLU (ts_common.cuh: silu4): 0 = one per element, 1 = one per pair (9.51 -> 9.40 ms per step)
#ifndef B200MD_NODE_TS_SILU
#define B200MD_NODE_TS_SILU 1
#endif
constexpr int kNodeSilu = B200MD_NODE_TS_SILU;
#ifndef B200MD_NODE_TS_PIPELINED
#define B200MD_NODE_TS_PIPELINED 1
#endif

enum NodeGemmKind {
  kNodeChain = 0,       // v = silu(acc + bias)                      -> next A
  kNodeChainAddU = 1,   // v = silu(acc + U[node])                   -> next A
  kNodeChainLast = 2,   // v = acc + bias (+ h) ; h <- v             -> next A when an output group follows
  kNodeOut = 3,         // out[node] = acc + bias                     (P, Q, U)
};

constexpr int kNodeTensorMaps = 4;   // B200MD_NODE_TS_TMA_STORE: h, and the (at most three) outputs of a call in GEMM order

struct NodeTsGemm {
  int K, N, kind;
  int fresh_a;          // the MMA waits for the epilogue to publish the A chunks (else: same A as the previous GEMM)
  int wait_d_free;      // the accumulator buffer still holds the previous output: each column group waits until the
                        // epilogue has drained that group (so the MMAs of one output run under the drain of the previous)
  int signal_d_free;    // ... and the epilogue says so, group by group, while draining this one
  int64_t w16;          // offset (floats) in wtc of the fp16 image of the first K chunk
  int64_t unscale;      // offset in wtc of 2^-s: the image holds 2^s w (tc_common.cuh), the epilogue multiplies it back
  int64_t bias;         // offset in packed, or -1
  float* out;           // kNodeOut: destination [n_nodes, N]
  int tm;               // kNodeOut: index of its tensor map in NodeTsArgs::tm (B200MD_NODE_TS_TMA_STORE)
};

struct NodeTsArgs {
  alignas(64) CUtensorMap tm[kNodeTensorMaps];   // [n_nodes, ld] fp32, box 32 x 32, SWIZZLE_128B: 0 = h, 1 .. = the outputs
  const float* wtc;
  const float* packed;
  int n_gemms;
  NodeTsGemm g[kMaxGemms];
  int do_embed, do_update, do_readout;
  const int64_t* atom_types; const float* x_rel; const float* sigma; float sigma_scalar; int sigma_stride;
  float *z, *xcur, *xagg, *h, *agg;
  const float* U;
  const int32_t* deg;
  float* logits_out; float* score_out; int32_t* status;
  int64_t n_nodes; int32_t n_atoms;
  int C, F0, D, H, M, nb;
  int residual, coords_mean, message_mean, mask_logit;
  int64_t bloch, proj, embed_w, embed_b, readout_w, readout_b;
  int trace;   // B200MD_TS_TRACE builds: record the cycle trace of this launch
};

constexpr int kNodeMaxClasses = 8;

// Optional cycle trace of CTA 0 (build with -DB200MD_TS_TRACE; read back with b200md_debug_node_trace, scripts/node_trace.py):
// one record (warp, event, argument, clock) per NODE_TRACE call of lane 0 of the MMA warp and of the two epilogue warps
// of lane quarter 0, for the launch that updates layer 1 (update + P / Q / U outputs).
#ifdef B200MD_TS_TRACE
constexpr unsigned kNodeTraceRecords = 4096;
__device__ uint4 g_node_trace[3 * kNodeTraceRecords];
__device__ unsigned g_node_trace_count;
#define NODE_TRACE(ev, arg)                                                                                     \
  do {                                                                                                          \
    if (trace_on && trace_idx < kNodeTraceRecords) {                                                            \
      g_node_trace[trace_slot * kNodeTraceRecords + trace_idx] = make_uint4((unsigned)warp, (unsigned)(ev), (unsigned)(arg), (unsigned)clock()); \
      ++trace_idx;                                                                                              \
    }                                                                                                           \
  } while (0)
#else
#define NODE_TRACE(ev, arg) do { } while (0)
#endif

// Weight ring of the node kernel: 10 stages of 16 KB (the 12 of ts_common.cuh minus the 32 KB of row staging below; the
// producer is always ahead of the MMAs -- it waits on "stage empty" for most of its life -- so two stages less cost nothing).
// B200MD_NODE_TS_INPUT_WARPS = 1: the two idle warps of the control warpgroup bring every row the epilogue READS from global
// memory -- the agg rows of the tile's first A operand, the U rows of the first layer, the residual h rows of the last one, in
// that order, one 64-column chunk of the tile's 128 rows at a time -- into a ring of shared-memory stages with warp-coalesced
// cp.async (16 lanes x 16 bytes = one row piece; XOR-swizzled 16-byte groups, conflict-free for the row-per-lane reads),
// completion on an mbarrier.  The chunks run ahead of their use by the depth of the ring, across layer and TILE boundaries
// (the next tile's first agg chunks arrive while the outputs of this one are drained), with no register of the epilogue
// threads and no exposed L2 / HBM round trip on the tile's serial chain.  Paid for by the weight ring.
#ifndef B200MD_NODE_TS_INPUT_WARPS
#define B200MD_NODE_TS_INPUT_WARPS 1
#endif
#ifndef B200MD_NODE_TS_IN_STAGES
#define B200MD_NODE_TS_IN_STAGES 2
#endif
// B200MD_NODE_TS_DRAIN_PIPELINED = 1 (needs the registers the input warps free): in the layers WITHOUT a SiLU (the last chain
// layer and the P / Q / U outputs) the accumulator columns of chunk c + 1 are requested (tcgen05.ld) before chunk c is stored to
// global memory, a column group is reported drained as soon as it is in registers (not after its rows have been stored), and
// h' is published as the next A operand before it is stored.
#ifndef B200MD_NODE_TS_DRAIN_PIPELINED
#define B200MD_NODE_TS_DRAIN_PIPELINED 1
#endif
static_assert(!B200MD_NODE_TS_DRAIN_PIPELINED || (B200MD_NODE_TS_INPUT_WARPS && B200MD_NODE_TS_PIPELINED), "the pipelined drain reads its side rows from the input ring and handles the layers without a SiLU only");
// B200MD_NODE_TS_EARLY_X = 1: the loads of the coordinate update (x += agg_x: first 8 coordinates of the row) and the NEXT
// tile's degree are issued at the top of the tile, so that their round trips run under the first A operand instead of on
// the critical path of the part-0 warps (cycle trace r02e: 4.5 k cycles between the first A operand and their first epilogue).
#ifndef B200MD_NODE_TS_EARLY_X
#define B200MD_NODE_TS_EARLY_X 1
#endif
// B200MD_NODE_TS_L2_AHEAD = 1: when the input warps have requested the last chunk of a tile, they prefetch into the L2 the
// chunks of the NEXT tile that the ring cannot hold yet (agg chunks 2, 3, ...): requested later, behind this tile's
// output stores, they took 4 - 6 k cycles to arrive (trace r02e).  2: also the next tile's U rows.
#ifndef B200MD_NODE_TS_L2_AHEAD
#define B200MD_NODE_TS_L2_AHEAD 0
#endif
// B200MD_NODE_TS_TMA_STORE = 1: the rows the epilogue WRITES (h', P, Q, U: 32 rows x 128 bytes per warp and chunk) leave through
// the TMA: the warp stages its tile exactly as before (XOR-swizzled 16-byte groups = the TMA's SWIZZLE_128B pattern), and one
// lane issues ONE 2-D tensor store (cp.async.bulk.tensor.2d, SASS UTMASTG) of the 32 x 32 box instead of 8 shared loads and 8
// global stores per thread; rows beyond the end of the array are clipped by the tensor map.  The store is asynchronous: the
// warp goes on to the next chunk and only waits (bulk wait_group.read) before it refills the staging tile -- the cycle trace
// r02e showed 1.2 - 1.5 k of a chunk's 2.6 k cycles in the drain spent behind the back-pressure of the global stores.
#ifndef B200MD_NODE_TS_LATE_ZERO
#define B200MD_NODE_TS_LATE_ZERO 0
#endif
#ifndef B200MD_NODE_TS_STAGGER_GROUPS
#define B200MD_NODE_TS_STAGGER_GROUPS 1
#endif
#ifndef B200MD_NODE_TS_STAGGER_CYCLES
#define B200MD_NODE_TS_STAGGER_CYCLES 50000
#endif
#ifndef B200MD_NODE_TS_TMA_STORE
#define B200MD_NODE_TS_TMA_STORE 0
#endif
constexpr bool kNodeInputWarps = B200MD_NODE_TS_INPUT_WARPS != 0;
static_assert(!kNodeInputWarps || B200MD_TS_WARPGROUPS, "the input warps are warps 2 and 3 of the control warpgroup");
constexpr int kNodeInStages = kNodeInputWarps ? B200MD_NODE_TS_IN_STAGES : 0;
constexpr int kNodeInStageBytes = kTileM * 64 * 4;   // 128 rows x one 64-column chunk
#ifndef B200MD_NODE_TS_B_STAGES
#define B200MD_NODE_TS_B_STAGES (B200MD_NODE_TS_INPUT_WARPS ? 10 - 2 * B200MD_NODE_TS_IN_STAGES : 10)
#endif
constexpr int kNodeBStages = B200MD_NODE_TS_B_STAGES;
static_assert(kNodeBStages % 2 == 0, "the MMA loop consumes the hi / lo stages of a K chunk in pairs");
// Row staging: 32 rows x 128 B per epilogue warp.  A TMEM lane is a ROW, so an epilogue thread owns 128 contiguous bytes
// of one row and a warp-wide 128-bit access touches 32 different lines: 32 L1 wavefronts per instruction, 2048 per 64-column
// chunk and SM, and the tile's 32 chunk-sized row accesses (agg, U, h in; agg reset, h, P, Q, U out) cost 65 k of its
// 149 k cycles (cycle trace, scripts/node_trace.py).  Through the staging tile every global access is made by 8 lanes per
// 128-byte line (4 lines per instruction): 8 x fewer wavefronts.  XOR-swizzled 16-byte pieces, conflict-free both ways.
#ifndef B200MD_NODE_TS_COALESCED
#define B200MD_NODE_TS_COALESCED 1
#endif
constexpr int kNodeStageBytes = 32 * 128;
struct NodeTsSmem {
  uint32_t b, stage, in, bars, tmem_slot, bias, scale, readout, logits, total;
};
__host__ __device__ constexpr NodeTsSmem node_ts_smem(int n_gemms) {
  NodeTsSmem L{};
  uint32_t o = 0;
  L.b = o; o += kNodeBStages * kTsBStageBytes;
  L.stage = o; o += kTsEpiWarps * kNodeStageBytes;
  L.in = o; o += kNodeInStages * kNodeInStageBytes;
  L.bars = o; o += (2 * kNodeBStages + kTsSlots * kTsParts + 2 * kTsNSplit + kTsNSplit + 2 * kNodeInStages) * 8;
  L.tmem_slot = o; o += 16;
  o = (o + 15u) & ~15u;   // the bias rows are read with 128-bit shared loads
  L.bias = o; o += (uint32_t)(n_gemms > 0 ? n_gemms : 1) * kMaxN * 4;
  L.scale = o; o += 16 * 4;
  L.readout = o; o += kNodeMaxClasses * kMaxN * 4;            // classification weights [C][H]
  L.logits = o; o += kTsParts * kTileM * kNodeMaxClasses * 4; // partial logits [part][row][C]
  L.total = o;
  return L;
}

static_assert(node_ts_smem(kMaxGemms).total + 1024 <= 227 * 1024, "node kernel: shared memory (weight ring + staging + input ring) beyond 227 KB");

struct NodeTsBars {
  uint32_t bar0;
  __device__ __forceinline__ uint32_t b_full(uint32_t s) const { return bar0 + 8u * s; }
  __device__ __forceinline__ uint32_t b_empty(uint32_t s) const { return bar0 + 8u * (kNodeBStages + s); }
  __device__ __forceinline__ uint32_t a_full(uint32_t slot, uint32_t part) const {
    return bar0 + 8u * (2 * kNodeBStages + slot * kTsParts + part);
  }
  __device__ __forceinline__ uint32_t acc_full(uint32_t buf, uint32_t group) const {
    return bar0 + 8u * (2 * kNodeBStages + kTsSlots * kTsParts + buf * kTsNSplit + group);
  }
  // "column group `group` of the output accumulator has been drained"
  __device__ __forceinline__ uint32_t d_free(uint32_t group) const {
    return bar0 + 8u * (2 * kNodeBStages + kTsSlots * kTsParts + 2 * kTsNSplit + group);
  }
  // input ring (B200MD_NODE_TS_INPUT_WARPS): "stage s holds a chunk" (64 input lanes), "stage s has been read" (8 epilogue warps)
  __device__ __forceinline__ uint32_t in_full(uint32_t s) const {
    return bar0 + 8u * (2 * kNodeBStages + kTsSlots * kTsParts + 3 * kTsNSplit + s);
  }
  __device__ __forceinline__ uint32_t in_empty(uint32_t s) const { return in_full(s) + 8u * kNodeInStages; }
};

__device__ __forceinline__ float4 lds128_staged(uint32_t saddr) {   // ordered against the staging stores (volatile, clobber)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void load_cols(const float* p, float (&v)[kTsCols]) {
#pragma unroll
  for (int j = 0; j < kTsCols / 4; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void store_cols(float* p, const float (&v)[kTsCols]) {
#pragma unroll
  for (int j = 0; j < kTsCols / 4; ++j)
    *reinterpret_cast<float4*>(p + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_tile, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_tile), "r"(col), "r"(row)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(kTsThreads, 1) egnn_node_ts_kernel(const __grid_constant__ NodeTsArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const NodeTsSmem L = node_ts_smem(a.n_gemms);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  const float* Wp = a.packed;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (int)((a.n_nodes + kTileM - 1) / kTileM);
  const int H = a.H, M = a.M, D = a.D;

  for (int t = threadIdx.x; t < a.n_gemms * kMaxN; t += kTsThreads) {
    const int g = t / kMaxN, n = t - g * kMaxN;
    // the bias of a SiLU layer is stored x kF16ActScale, like the activation it produces (ts_common.cuh)
    s_bias[t] = (n < a.g[g].N && a.g[g].bias >= 0) ? Wp[a.g[g].bias + n] * (a.g[g].kind == kNodeChain ? kF16ActScale : 1.0f) : 0.f;
  }
  if (threadIdx.x < a.n_gemms) reinterpret_cast<float*>(smem + L.scale)[threadIdx.x] = a.wtc[a.g[threadIdx.x].unscale];
  float* s_readout = reinterpret_cast<float*>(smem + L.readout);
  float* s_logits = reinterpret_cast<float*>(smem + L.logits);
  if (a.do_readout)
    for (int t = threadIdx.x; t < a.C * H; t += kTsThreads) s_readout[(t / H) * kMaxN + (t % H)] = Wp[a.readout_w + t];
  const uint32_t s_base = smem_u32(smem);
  NodeTsBars bars;
  bars.bar0 = s_base + L.bars;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kNodeBStages; ++s) { mbar_init(bars.b_full(s), 1); mbar_init(bars.b_empty(s), 1); }
    for (int s = 0; s < kTsSlots; ++s)
      for (int j = 0; j < kTsParts; ++j) mbar_init(bars.a_full(s, j), 4);
    for (int s = 0; s < 2 * kTsNSplit; ++s) mbar_init(bars.acc_full(s / kTsNSplit, s % kTsNSplit), 1);
    for (int s = 0; s < kTsNSplit; ++s) mbar_init(bars.d_free(s), kTsEpiWarps);
    for (int s = 0; s < kNodeInStages; ++s) { mbar_init(bars.in_full(s), 64); mbar_init(bars.in_empty(s), kTsEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_base + L.tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base_all = *reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);
  const int n_iterations = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
#if B200MD_NODE_TS_STAGGER_GROUPS > 1
  // Experiment: CTAs start out of phase (group g of G waits g x B200MD_NODE_TS_STAGGER_CYCLES), so that the output drains of
  // the 148 CTAs -- all tiles take the same time, the CTAs run in lockstep -- do not hit the memory system together.
  if (n_iterations > 2) {
    const long long wait_cycles = (long long)(blockIdx.x % B200MD_NODE_TS_STAGGER_GROUPS) * B200MD_NODE_TS_STAGGER_CYCLES;
    const long long t0 = clock64();
    while (clock64() - t0 < wait_cycles) __nanosleep(256);
  }
#endif
#ifdef B200MD_TS_TRACE
  bool trace_on = false;
  unsigned trace_idx = 0;
  const unsigned trace_slot = warp == 1 ? 0u : (warp == kTsFirstEpiWarp ? 1u : 2u);
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) g_node_trace_count = 3 * kNodeTraceRecords;
#endif

  if (warp < kTsFirstEpiWarp) {   // setmaxnreg at the top of each warpgroup's branch (see egnn_edge_ts.cuh)
#if B200MD_TS_WARPGROUPS
  ts_regs_release<kTsRegsControl>();
#endif
  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t cb = 0;
      for (int it = 0; it < n_iterations; ++it)
        for (int g = 0; g < a.n_gemms; ++g) {
          const NodeTsGemm& G = a.g[g];
          const int nk = G.K / 64, nh = G.N / kTsNSplit;
          const uint32_t bytes = (uint32_t)nh * kRowBytes;
          const int64_t part_floats = (int64_t)G.N * (kRowBytes / 4);
          const float* src = a.wtc + G.w16;
          for (int h = 0; h < kTsNSplit; ++h)
            for (int part = 0; part < 2 * nk; ++part, ++cb) {
              const uint32_t s = cb % kNodeBStages;
              mbar_wait(bars.b_empty(s), ((cb / kNodeBStages) & 1u) ^ 1u);
              mbar_arrive_expect_tx(bars.b_full(s), bytes);
              bulk_g2s(s_base + L.b + s * kTsBStageBytes, src + (int64_t)part * part_floats + (int64_t)h * nh * (kRowBytes / 4),
                       bytes, bars.b_full(s));
            }
        }
    }
  } else if (kNodeInputWarps && warp >= 2) {
    // ===== input warps: rows [64 hw, 64 hw + 64) of every chunk the epilogue reads, in the epilogue's order =====
    const int hw = warp - 2;
    const uint32_t piece = (uint32_t)(lane & 15);
    const uint32_t in_base = s_base + L.in;
    uint32_t n = 0;
    auto fill = [&](const float* rows, int ld, int col, int64_t row0, int rows_valid_tile) {
      const uint32_t s = n % (uint32_t)(kNodeInStages > 0 ? kNodeInStages : 1);
      mbar_wait(bars.in_empty(s), ((n / (uint32_t)(kNodeInStages > 0 ? kNodeInStages : 1)) & 1u) ^ 1u);
      const uint32_t dst0 = in_base + s * kNodeInStageBytes;
      const float* src0 = rows + row0 * ld + col + piece * 4;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {   // two rows per instruction: lanes 0-15 row 2 i, lanes 16-31 row 2 i + 1
        const int r = 64 * hw + 2 * i + (lane >> 4);
        if (r < rows_valid_tile) cp_async16(dst0 + (uint32_t)r * 256u + ((piece ^ (uint32_t)(r & 7)) << 4), src0 + (int64_t)r * ld);
      }
      cp_async_mbar_arrive_noinc(bars.in_full(s));
      ++n;
    };
    for (int it = 0; it < n_iterations; ++it) {
      const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kTileM;
      const int rows_valid_tile = (int)max((int64_t)0, min((int64_t)kTileM, a.n_nodes - row0));
      if (!a.do_embed)
        for (int c = 0; c < M / 64; ++c) fill(a.agg, M, c * 64, row0, rows_valid_tile);
      for (int g = 0; g < a.n_gemms; ++g) {
        const NodeTsGemm& G = a.g[g];
        const float* side_rows = nullptr;
        if (G.kind == kNodeChainAddU) side_rows = a.U;
        if (G.kind == kNodeChainLast && a.residual) side_rows = a.h;
        if (side_rows != nullptr)
          for (int c = 0; c < G.N / 64; ++c) fill(side_rows, H, c * 64, row0, rows_valid_tile);
      }
#if B200MD_NODE_TS_L2_AHEAD
      if (it + 1 < n_iterations) {
        const int64_t next0 = row0 + (int64_t)gridDim.x * kTileM;
        const int next_valid = (int)max((int64_t)0, min((int64_t)kTileM, a.n_nodes - next0));
        // one 128-byte line per lane and instruction: lanes 0-15 / 16-31 cover rows 2 i, 2 i + 1, lane & 15 the line of the row
        auto l2_ahead = [&](const float* rows, int ld, int first_col, int n_cols) {
          if ((lane & 15) * 32 >= n_cols) return;   // 128-byte lines per row: n_cols / 32
          const int r0 = 64 * hw + (lane >> 4);
          const float* line = rows + (next0 + r0) * ld + first_col + (lane & 15) * 32;
#pragma unroll 1
          for (int r = r0; r < 64 * hw + 64 && r < next_valid; r += 2, line += 2 * (int64_t)ld) prefetch_l2(line);
        };
        if (!a.do_embed && M > kNodeInStages * 64) l2_ahead(a.agg, M, kNodeInStages * 64, M - kNodeInStages * 64);
#if B200MD_NODE_TS_L2_AHEAD > 1
        if (a.do_update) l2_ahead(a.U, H, 0, H);
#endif
      }
#endif
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp converged, elected lane issues; see egnn_edge_ts.cuh) =====
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_all, 0);
    const uint32_t b_ring = s_base + L.b;
    const bool leader = elect_one();
    uint32_t cb = 0, mca = 0, abuf = 0, df = 0;
    for (int it = 0; it < n_iterations; ++it)
      for (int g = 0; g < a.n_gemms; ++g) {
#ifdef B200MD_TS_TRACE
        trace_on = (a.trace && blockIdx.x == 0 && lane == 0 && it >= 2 && it < 4);
#endif
        const NodeTsGemm& G = a.g[g];
        const int nk = G.K / 64, nh = G.N / kTsNSplit;
        const uint32_t idesc = make_idesc(nh, 0u);
        const uint32_t a_col = abuf * kMaxN, d_col = a_col ^ kMaxN;
        for (int h = 0; h < kTsNSplit; ++h) {
          NODE_TRACE(130 + h, g);
          if (G.wait_d_free) { mbar_wait(bars.d_free(h), df & 1u); tc_fence_after(); }
          NODE_TRACE(140 + h, g);
          const uint32_t d_tmem = tmem_base + d_col + (uint32_t)(h * nh);
          for (int c = 0; c < nk; ++c) {
            const uint32_t s1 = cb % kNodeBStages, s2 = (cb + 1) % kNodeBStages;
            mbar_wait(bars.b_full(s1), (cb / kNodeBStages) & 1u);
            mbar_wait(bars.b_full(s2), ((cb + 1) / kNodeBStages) & 1u);
            const uint32_t bd1 = desc_lo_sw128(b_ring + s1 * kTsBStageBytes), bd2 = desc_lo_sw128(b_ring + s2 * kTsBStageBytes);
            const uint32_t chunk = mca + (uint32_t)c;
            const uint32_t a_chunk = tmem_base + a_col + (uint32_t)(c * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (G.fresh_a && h == 0 && (j % (4 / kTsParts)) == 0) {
                NODE_TRACE(100 + j, g * 16 + c);
                mbar_wait(bars.a_full(chunk % kTsSlots, j / (4 / kTsParts)), (chunk / kTsSlots) & 1u);
                NODE_TRACE(110 + j, g * 16 + c);
              }
              tc_fence_after();
              if (leader) {
                umma_f16_ts(d_tmem, a_chunk + j * 16, bd1 + 2 * j, idesc, (c | j) != 0);
                umma_f16_ts(d_tmem, a_chunk + j * 16 + 8, bd1 + 2 * j, idesc | kIdescA2, 1u);
                umma_f16_ts(d_tmem, a_chunk + j * 16, bd2 + 2 * j, idesc, 1u);
              }
            }
            if (leader) {
              umma_commit(bars.b_empty(s1));
              umma_commit(bars.b_empty(s2));
            }
            __syncwarp();
            cb += 2;
          }
          if (leader) umma_commit(bars.acc_full(d_col / kMaxN, h));
          __syncwarp();
          NODE_TRACE(120 + h, g);
        }
        if (G.wait_d_free) ++df;
        if (G.fresh_a) mca += nk;
        const bool becomes_a = (G.kind == kNodeChain || G.kind == kNodeChainAddU ||
                                (G.kind == kNodeChainLast && g + 1 < a.n_gemms));
        if (becomes_a) abuf ^= 1u;
      }
  }
  } else {
#if B200MD_TS_WARPGROUPS
    ts_regs_acquire<kTsRegsEpilogue>();
#endif
    // ===== epilogue warps =====
    const int q = warp & 3;
    const int part = (warp - kTsFirstEpiWarp) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_taddr = tmem_base_all + ((uint32_t)(q * 32) << 16);
    const uint32_t sa_bias = s_base + L.bias;
    const uint32_t stage = s_base + L.stage + (uint32_t)(warp - kTsFirstEpiWarp) * kNodeStageBytes;
    uint32_t ca = 0, abuf = 0;
    int32_t deg_next = 1;   // B200MD_NODE_TS_EARLY_X: the next tile's degree of this row
    (void)deg_next;
    uint32_t inq = 0;   // chunks taken from the input ring
    (void)inq;
    uint32_t uses[2] = {0u, 0u};
    unsigned local_status = 0;
    float checksum = 0.f;   // inf / NaN of an fp16 overflow propagates into everything stored: one test at the end
#if B200MD_NODE_TS_PIPELINED
    const uint32_t tok_mine = 2u + 2u * (uint32_t)q + (uint32_t)part, tok_peer = 2u + 2u * (uint32_t)q + (uint32_t)(part ^ 1);
    if (part == 1) sfu_arrive(tok_peer);   // part 0 goes first
#endif

    // `scale`: kF16ActScale for values that are not scaled yet (every A operand is stored x kF16ActScale), 1 for the
    // outputs of the SiLU layers, which already are
    auto publish = [&](uint32_t col, const float (&v_in)[kTsCols], float scale) {
      float v[kTsCols];
#pragma unroll
      for (int j = 0; j < kTsCols; ++j) v[j] = v_in[j] * scale;
      uint32_t w[kTsCols];
      split_hi_lo(v, w);
#pragma unroll
      for (int i = 0; i < kTsCols / 16; ++i) tmem_st16(lane_taddr + col + 16 * i, w + 16 * i);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars.a_full(ca % kTsSlots, part));
      ++ca;
    };

    for (int it = 0; it < n_iterations; ++it) {
#ifdef B200MD_TS_TRACE
      trace_on = (a.trace && blockIdx.x == 0 && lane == 0 && q == 0 && it >= 2 && it < 4);
#endif
      NODE_TRACE(20, it);
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int64_t node = (int64_t)tile * kTileM + row;
      const bool valid = node < a.n_nodes;
#if B200MD_NODE_TS_EARLY_X
      if (it == 0) deg_next = valid ? a.deg[node] : 1;
      const float degf = fmaxf((float)deg_next, 1.0f);
      {
        const int64_t node_next = node + (int64_t)gridDim.x * kTileM;
        deg_next = (it + 1 < n_iterations && node_next < a.n_nodes) ? a.deg[node_next] : 1;
      }
#else
      const float degf = valid ? fmaxf((float)a.deg[node], 1.0f) : 1.0f;
#endif
      // mean = sum x (1 / max(deg, 1)) with the correctly rounded reciprocal: at most 1 ulp from the reference's division
      // (models/egnn_utils.py:53-70), and 32 IEEE divisions per thread and chunk were 2.4 k of the 4 k cycles each chunk
      // of the tile's first A operand took (cycle trace)
      const float inv_deg = __frcp_rn(degf);
      // ---- warp-cooperative row access through the staging tile (see kNodeStageBytes) ----
      const int64_t warp_row0 = (int64_t)tile * kTileM + q * 32;   // node of lane 0's row
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.n_nodes - warp_row0));   // warp-uniform
      const int crow = lane >> 3, cpiece = lane & 7;   // cooperative role: instruction i covers rows 4 i + crow, piece cpiece
#if !B200MD_NODE_TS_INPUT_WARPS   // (with the input warps the epilogue reads its rows from the shared-memory ring: input_collect)
      struct RowTile { float4 p[8]; };
      // request columns [col, col + 32) of the warp's 32 rows of `rows` (row stride ld floats) ...
      auto rows_request = [&](const float* rows, int ld, int col, RowTile& t) {
#if B200MD_NODE_TS_COALESCED
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int R = 4 * i + crow;
          t.p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rows != nullptr && R < rows_valid) t.p[i] = *(reinterpret_cast<const float4*>(rows + (warp_row0 + R) * ld + col) + cpiece);
        }
#else
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t.p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rows != nullptr && valid) t.p[i] = *(reinterpret_cast<const float4*>(rows + node * ld + col) + i);
        }
#endif
      };
      // ... and turn the requested tile into this lane's own row (rows beyond the end read as zero)
      auto rows_collect = [&](const RowTile& t, float (&v)[kTsCols]) {
#if B200MD_NODE_TS_COALESCED
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int R = 4 * i + crow;
          sts128(stage + (uint32_t)R * 128u + (uint32_t)((cpiece ^ (R & 7)) << 4), t.p[i].x, t.p[i].y, t.p[i].z, t.p[i].w);
        }
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const float4 f = lds128_staged(stage + (uint32_t)lane * 128u + (uint32_t)((p ^ (lane & 7)) << 4));
          v[4 * p] = f.x; v[4 * p + 1] = f.y; v[4 * p + 2] = f.z; v[4 * p + 3] = f.w;
        }
        __syncwarp();
#else
#pragma unroll
        for (int p = 0; p < 8; ++p) { v[4 * p] = t.p[p].x; v[4 * p + 1] = t.p[p].y; v[4 * p + 2] = t.p[p].z; v[4 * p + 3] = t.p[p].w; }
#endif
      };
#endif
      // store this lane's row values as columns [col, col + 32) of the warp's rows of `rows`
      auto rows_store = [&](float* rows, int ld, int col, const float (&v)[kTsCols], int tm) {
#if B200MD_NODE_TS_TMA_STORE
        (void)rows; (void)ld;
        if (lane == 0) bulk_wait_read();   // this warp's previous store has read the staging tile
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 8; ++p)
          sts128(stage + (uint32_t)lane * 128u + (uint32_t)((p ^ (lane & 7)) << 4), v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]);
        fence_proxy_async();               // the tile is read by the async proxy
        __syncwarp();
        NODE_TRACE(35, col);
        if (lane == 0 && rows_valid > 0) {
          tma_store_2d(&a.tm[tm], stage, col, (int)warp_row0);
          bulk_commit();
        }
#elif B200MD_NODE_TS_COALESCED
#pragma unroll
        for (int p = 0; p < 8; ++p)
          sts128(stage + (uint32_t)lane * 128u + (uint32_t)((p ^ (lane & 7)) << 4), v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]);
        __syncwarp();
        NODE_TRACE(35, col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int R = 4 * i + crow;
          const float4 f = lds128_staged(stage + (uint32_t)R * 128u + (uint32_t)((cpiece ^ (R & 7)) << 4));
          if (R < rows_valid) *(reinterpret_cast<float4*>(rows + (warp_row0 + R) * ld + col) + cpiece) = f;
        }
        __syncwarp();
#else
        if (valid) store_cols(rows + node * ld + col, v);
#endif
      };
      auto rows_zero = [&](float* rows, int ld, int col) {
#if B200MD_NODE_TS_COALESCED
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int R = 4 * i + crow;
          if (R < rows_valid) *(reinterpret_cast<float4*>(rows + (warp_row0 + R) * ld + col) + cpiece) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#else
        if (valid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) *(reinterpret_cast<float4*>(rows + node * ld + col) + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#endif
      };
#if B200MD_NODE_TS_INPUT_WARPS
      // this lane's row of the next chunk of the input ring (rows beyond the end read as zero)
      auto input_collect = [&](float (&v)[kTsCols]) {
        const uint32_t s = inq % (uint32_t)kNodeInStages;
        mbar_wait(bars.in_full(s), (inq / (uint32_t)kNodeInStages) & 1u);
        const uint32_t at = s_base + L.in + s * kNodeInStageBytes + (uint32_t)row * 256u + (uint32_t)part * 128u;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) f = lds128_staged(at + (uint32_t)((p ^ (lane & 7)) << 4));
          v[4 * p] = f.x; v[4 * p + 1] = f.y; v[4 * p + 2] = f.z; v[4 * p + 3] = f.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars.in_empty(s));
        ++inq;
      };
#endif
      float sg = 0.f;
      int atc = 0;
      if (a.do_embed && valid) {
        const int64_t at = a.atom_types[node];
        sg = a.sigma ? a.sigma[(node / a.n_atoms) * a.sigma_stride] : a.sigma_scalar;
        if (part == 0) {
          if (at < 0 || at >= a.C) local_status |= B200MD_STATUS_A_OUT_OF_RANGE;
          if (!(sg > 0.f)) local_status |= B200MD_STATUS_SIGMA_OUT_OF_RANGE;
        }
        atc = (int)min(max(at, (int64_t)0), (int64_t)(a.C - 1));
      }
      if (part == 0 && valid) {
        if (a.do_embed) {
          // uplift z = [cos(2 pi k.x), sin(2 pi k.x)] per Bloch vector   egnn_score_network.py:196-224
          const float two_pi = 6.283185307179586f;
          float xr[3];
          for (int s = 0; s < 3; ++s) {
            xr[s] = a.x_rel[node * 3 + s];
            if (!(xr[s] >= 0.0f && xr[s] < 1.0f)) local_status |= B200MD_STATUS_X_OUT_OF_RANGE;
          }
          for (int b = 0; b < a.nb; ++b) {
            float kr = 0.f;
            for (int s = 0; s < 3; ++s) kr += Wp[a.bloch + b * 3 + s] * (two_pi * xr[s]);
            float sn, cs;
            sincosf(kr, &sn, &cs);
            a.z[node * D + 2 * b] = cs; a.z[node * D + 2 * b + 1] = sn;
            a.xcur[node * D + 2 * b] = cs; a.xcur[node * D + 2 * b + 1] = sn;
            a.xagg[node * D + 2 * b] = 0.f; a.xagg[node * D + 2 * b + 1] = 0.f;
          }
        }
      }
      // coord += agg (models/egnn.py:212-215).  Nothing in this tile's GEMMs depends on it, so it runs BEHIND the first A
      // operand (while the first MMAs execute), and all its loads are issued before the first store: written as one
      // load-add-store-store per coordinate it is a chain of 2 D dependent memory round trips (the compiler cannot prove
      // that xagg and xcur do not alias) that sat in front of the tile's first MMA for 20 k cycles (cycle trace).
#if B200MD_NODE_TS_EARLY_X
      float xa_early[8], xc_early[8];
      if (a.do_update && part == 0 && valid) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          xa_early[k] = (k < D) ? a.xagg[node * D + k] : 0.f;
          xc_early[k] = (k < D) ? a.xcur[node * D + k] : 0.f;
        }
      }
#endif
      auto coordinate_update = [&]() {
        if (!(a.do_update && part == 0 && valid)) return;
        for (int d0 = 0; d0 < D; d0 += 8) {
          float xa[8], xc[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
#if B200MD_NODE_TS_EARLY_X
            if (d0 == 0) { xa[k] = xa_early[k]; xc[k] = xc_early[k]; continue; }
#endif
            xa[k] = (d0 + k < D) ? a.xagg[node * D + d0 + k] : 0.f;
            xc[k] = (d0 + k < D) ? a.xcur[node * D + d0 + k] : 0.f;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (d0 + k < D) {
              const float v = a.coords_mean ? xa[k] / degf : xa[k];
              a.xcur[node * D + d0 + k] = xc[k] + v;
              a.xagg[node * D + d0 + k] = 0.f;
            }
        }
      };

      // ---- first A operand of the tile ----
      const uint32_t a_col0 = abuf * kMaxN;
      if (a.do_embed) {
        // h = W_in [sigma, onehot(a)] + b   egnn_score_network.py:165-194, egnn.py:376
        for (int c = 0; c < H / 64; ++c) {
          const int col0 = c * 64 + part * kTsCols;
          float v[kTsCols];
#pragma unroll
          for (int j = 0; j < kTsCols; ++j) {
            const int o = col0 + j;
            v[j] = valid ? Wp[a.embed_b + o] + (Wp[a.embed_w + o * a.F0] * sg + Wp[a.embed_w + o * a.F0 + 1 + atc]) : 0.f;
          }
          rows_store(a.h, H, col0, v, 0);
          publish(a_col0 + (uint32_t)col0, v, kF16ActScale);
        }
        for (int c = 0; c < M / 64; ++c) rows_zero(a.agg, M, c * 64 + part * kTsCols);
      } else {
        // aggregated messages (mean: / max(deg, 1)), reset for the next layer   models/egnn.py:178-183
        // (the rows of chunk c + 1 are requested before chunk c is converted and published)
#if !B200MD_NODE_TS_INPUT_WARPS
        RowTile ahead;
        rows_request(a.agg, M, part * kTsCols, ahead);
#endif
        for (int c = 0; c < M / 64; ++c) {
          const int col0 = c * 64 + part * kTsCols;
          float v[kTsCols];
          NODE_TRACE(22, c);
#if B200MD_NODE_TS_INPUT_WARPS
          input_collect(v);
          NODE_TRACE(23, c);
#else
          rows_collect(ahead, v);   // the loads have returned here: the reset below may follow (a store issued right
                                    // behind loads of the same addresses doubles the time of the whole kernel, measured)
          NODE_TRACE(23, c);
          if (c + 1 < M / 64) rows_request(a.agg, M, col0 + 64, ahead);
#endif
          if (a.message_mean) {
#pragma unroll
            for (int j = 0; j < kTsCols; ++j) v[j] = B200MD_NODE_TS_EXACT_DIV ? v[j] / degf : v[j] * inv_deg;
          }
          NODE_TRACE(24, c);
#if !B200MD_NODE_TS_LATE_ZERO
          rows_zero(a.agg, M, col0);
#endif
          NODE_TRACE(25, c);
          publish(a_col0 + (uint32_t)col0, v, kF16ActScale);
          NODE_TRACE(21, c);
        }
#if B200MD_NODE_TS_LATE_ZERO
        // the reset of the agg rows waits until the whole A operand is out: issued chunk by chunk, its 32 KB of stores per
        // chunk sat in the SM's memory queue in front of the input warps' requests for chunks 2 and 3 (4 - 6 k cycles from
        // request to arrival in the trace r02e); here they drain while the first GEMM runs
        for (int c = 0; c < M / 64; ++c) rows_zero(a.agg, M, c * 64 + part * kTsCols);
#endif
      }
      coordinate_update();

      // ---- the GEMMs ----
      float logit[kNodeMaxClasses];
#pragma unroll
      for (int cl = 0; cl < kNodeMaxClasses; ++cl) logit[cl] = 0.f;
      for (int g = 0; g < a.n_gemms; ++g) {
        const NodeTsGemm& G = a.g[g];
        const uint32_t dbuf = abuf ^ 1u;
        const uint32_t d_col = dbuf * kMaxN;
        const int nchunks = G.N / 64, nh = G.N / kTsNSplit;
        const uint32_t bias = sa_bias + (uint32_t)(g * kMaxN) * 4u;
        // accumulator = 2^s (weights) x kF16ActScale (A operand) x product
        const float unscale = lds32(s_base + L.scale + (uint32_t)g * 4u);   // 2^-s: gives kF16ActScale x the product
        const float unscale_true = unscale * (1.0f / kF16ActScale);         // gives the product itself
        const uint64_t unscale2 = pack2(unscale, unscale);
        const bool becomes_a = (G.kind == kNodeChain || G.kind == kNodeChainAddU ||
                                (G.kind == kNodeChainLast && g + 1 < a.n_gemms));
        int waited = -1;
        // the rows this epilogue adds to the accumulator (U, or the residual h) travel one chunk ahead of their use:
        // the L2 / HBM round trip of chunk c + 1 is hidden behind the conversion of chunk c (and, for chunk 0,
        // behind the wait for the accumulator)
        const float* side_rows = nullptr;   // warp-uniform
        if (G.kind == kNodeChainAddU) side_rows = a.U;
        if (G.kind == kNodeChainLast && a.residual) side_rows = a.h;
#if !B200MD_NODE_TS_INPUT_WARPS
        RowTile side_tile;
        rows_request(side_rows, H, part * kTsCols, side_tile);
#endif
        auto wait_for_chunk = [&](int c) {
          const int need = ts_last_group_of_chunk(c, nh);
          if (need > waited) {
            NODE_TRACE(1, g * 16 + c);
            mbar_wait(bars.acc_full(dbuf, need), uses[dbuf] & 1u);
            tc_fence_after();
            waited = need;
            NODE_TRACE(2, g * 16 + c);
          }
        };
#if B200MD_NODE_TS_PIPELINED
        if (G.kind == kNodeChain || G.kind == kNodeChainAddU) {
          // The SiLU layers bound this kernel (the MMA warp spends most of its time waiting for A chunks, the cycle samples
          // of profiles/ncu_full_r02a say): same chunk pipeline as the edge kernel.  The accumulator columns of chunk
          // c + 1 are requested (tcgen05.ld) before chunk c is split and published and collected afterwards, and the two
          // epilogue warps of a scheduler alternate in the SFU-bound SiLU phase (token: ts_common.cuh).
          float v[kTsCols];
          uint32_t r[kTsCols];
          auto activate = [&](int c) {
            const int col0 = c * 64 + part * kTsCols;
            uint32_t dep = 0;
            if (G.kind == kNodeChain) {
#pragma unroll
              for (int j = 0; j < kTsCols / 4; ++j) {
                const float4 b = lds128(bias + (uint32_t)(col0 + 4 * j) * 4u);
                silu4<kNodeSilu>(fma2(pack2(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])), unscale2, pack2(b.x, b.y)),
                                 fma2(pack2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), unscale2, pack2(b.z, b.w)),
                                 v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dep);
              }
            } else {
              float side[kTsCols];
              NODE_TRACE(26, g * 16 + c);
#if B200MD_NODE_TS_INPUT_WARPS
              input_collect(side);
              NODE_TRACE(27, g * 16 + c);
#else
              rows_collect(side_tile, side);
              NODE_TRACE(27, g * 16 + c);
              if (c + 1 < nchunks) rows_request(side_rows, H, col0 + 64, side_tile);   // U of the next chunk
#endif
#pragma unroll
              for (int j = 0; j < kTsCols / 4; ++j)
                silu4<kNodeSilu>(fma2(pack2(side[4 * j], side[4 * j + 1]), pack2(kF16ActScale, kF16ActScale),
                                      mul2(pack2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), unscale2)),
                                 fma2(pack2(side[4 * j + 2], side[4 * j + 3]), pack2(kF16ActScale, kF16ActScale),
                                      mul2(pack2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), unscale2)),
                                 v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dep);
            }
            sfu_release(tok_peer, dep);
          };
          auto load_acc = [&](int c) {
#pragma unroll
            for (int i = 0; i < kTsCols / 16; ++i) tmem_ld16_issue(lane_taddr + d_col + c * 64 + part * kTsCols + 16 * i, r + 16 * i);
          };
          auto collect_acc = [&]() {
            sfu_acquire(tok_mine);
#pragma unroll
            for (int i = 0; i < kTsCols / 16; ++i) tmem_ld_wait16(r + 16 * i);
          };
          wait_for_chunk(0);
          load_acc(0);
          collect_acc();
          activate(0);
          for (int c = 0; c < nchunks; ++c) {
            const bool more = c + 1 < nchunks;
            if (more) { wait_for_chunk(c + 1); load_acc(c + 1); }
            publish(d_col + (uint32_t)(c * 64 + part * kTsCols), v, 1.0f);
            NODE_TRACE(6, g * 16 + c);
            if (more) { collect_acc(); activate(c + 1); NODE_TRACE(5, g * 16 + c + 1); }
          }
          ++uses[dbuf];
          abuf ^= 1u;
          continue;
        }
#endif
#if B200MD_NODE_TS_DRAIN_PIPELINED
        {
          uint32_t r[kTsCols];
          wait_for_chunk(0);
#pragma unroll
          for (int i = 0; i < kTsCols / 16; ++i) tmem_ld16_issue(lane_taddr + d_col + part * kTsCols + 16 * i, r + 16 * i);
          for (int c = 0; c < nchunks; ++c) {
            const int col0 = c * 64 + part * kTsCols;
            const bool more = c + 1 < nchunks;
            float v[kTsCols];
#pragma unroll
            for (int i = 0; i < kTsCols / 16; ++i) tmem_ld_wait16(r + 16 * i);
            NODE_TRACE(28, g * 16 + c);
            if (G.signal_d_free) {   // every column group that ends inside this chunk is in registers: the next output's MMAs may overwrite it
#pragma unroll
              for (int h = 0; h < kTsNSplit; ++h)
                if ((h + 1) * nh > c * 64 && (h + 1) * nh <= (c + 1) * 64) {
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(bars.d_free(h));
                }
            }
#pragma unroll
            for (int j = 0; j < kTsCols / 4; ++j) {
              const float4 b = lds128(bias + (uint32_t)(col0 + 4 * j) * 4u);
              v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), unscale_true, b.x);
              v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), unscale_true, b.y);
              v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), unscale_true, b.z);
              v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), unscale_true, b.w);
            }
            if (more) {   // the next chunk's accumulator columns travel under this chunk's stores
              wait_for_chunk(c + 1);
#pragma unroll
              for (int i = 0; i < kTsCols / 16; ++i) tmem_ld16_issue(lane_taddr + d_col + col0 + 64 + 16 * i, r + 16 * i);
            }
            if (side_rows != nullptr) {   // warp-uniform: the residual h
              float u[kTsCols];
              input_collect(u);
              NODE_TRACE(29, g * 16 + c);
#pragma unroll
              for (int j = 0; j < kTsCols; ++j) v[j] = u[j] + v[j];
            }
            if (becomes_a) publish(d_col + (uint32_t)col0, v, kF16ActScale);   // h' is the A operand of the outputs
            NODE_TRACE(30, g * 16 + c);
            if (G.kind == kNodeChainLast) {
              rows_store(a.h, H, col0, v, 0);
              if (a.do_readout) {   // logits = W_c h + b_c (egnn.py:379): partial dot products while h' is in registers
#pragma unroll
                for (int cl = 0; cl < kNodeMaxClasses; ++cl)
                  if (cl < a.C) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < kTsCols / 4; ++j) {
                      const float4 w = *reinterpret_cast<const float4*>(s_readout + cl * kMaxN + col0 + 4 * j);
                      acc = fmaf(v[4 * j + 0], w.x, acc); acc = fmaf(v[4 * j + 1], w.y, acc);
                      acc = fmaf(v[4 * j + 2], w.z, acc); acc = fmaf(v[4 * j + 3], w.w, acc);
                    }
                    logit[cl] += acc;
                  }
              }
            } else {
              rows_store(G.out, G.N, col0, v, G.tm);
            }
            NODE_TRACE(34, g * 16 + c);
            if (valid) {
              float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
              for (int j = 0; j < kTsCols; j += 4) { c0 += v[j]; c1 += v[j + 1]; c2 += v[j + 2]; c3 += v[j + 3]; }
              checksum += (c0 + c1) + (c2 + c3);
            }
          }
          ++uses[dbuf];
          if (becomes_a) abuf ^= 1u;
          continue;
        }
#else
        for (int c = 0; c < nchunks; ++c) {
          wait_for_chunk(c);
          const int col0 = c * 64 + part * kTsCols;
          uint32_t r[kTsCols];
          float v[kTsCols];
          tmem_load_cols(lane_taddr + d_col + col0, r);
          NODE_TRACE(28, g * 16 + c);
          float u[kTsCols];
          if (side_rows != nullptr) {   // warp-uniform
#if B200MD_NODE_TS_INPUT_WARPS
            input_collect(u);
            NODE_TRACE(29, g * 16 + c);
#else
            rows_collect(side_tile, u);
            NODE_TRACE(29, g * 16 + c);
            if (c + 1 < nchunks) rows_request(side_rows, H, col0 + 64, side_tile);
            NODE_TRACE(32, g * 16 + c);
#endif
          }
          if (G.kind == kNodeChain) {
#pragma unroll
            for (int j = 0; j < kTsCols / 4; ++j) {
              const float4 b = lds128(bias + (uint32_t)(col0 + 4 * j) * 4u);
              silu2(fma2(pack2(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])), unscale2, pack2(b.x, b.y)), v[4 * j + 0], v[4 * j + 1]);
              silu2(fma2(pack2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), unscale2, pack2(b.z, b.w)), v[4 * j + 2], v[4 * j + 3]);
            }
          } else if (G.kind == kNodeChainAddU) {
#pragma unroll
            for (int j = 0; j < kTsCols / 2; ++j)
              silu2(fma2(pack2(u[2 * j], u[2 * j + 1]), pack2(kF16ActScale, kF16ActScale),
                         mul2(pack2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), unscale2)), v[2 * j], v[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < kTsCols / 4; ++j) {
              const float4 b = lds128(bias + (uint32_t)(col0 + 4 * j) * 4u);
              v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), unscale_true, b.x);
              v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), unscale_true, b.y);
              v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), unscale_true, b.z);
              v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), unscale_true, b.w);
            }
            if (G.kind == kNodeChainLast) {
              if (a.residual) {
#pragma unroll
                for (int j = 0; j < kTsCols; ++j) v[j] = u[j] + v[j];
              }
              NODE_TRACE(33, (g * 16 + c) | ((__float_as_uint(v[0] + v[31]) & 1u) << 20));
              rows_store(a.h, H, col0, v, 0);
              NODE_TRACE(34, g * 16 + c);
              if (a.do_readout) {   // logits = W_c h + b_c (egnn.py:379): partial dot products while h' is in registers
#pragma unroll
                for (int cl = 0; cl < kNodeMaxClasses; ++cl)
                  if (cl < a.C) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < kTsCols / 4; ++j) {
                      const float4 w = *reinterpret_cast<const float4*>(s_readout + cl * kMaxN + col0 + 4 * j);
                      acc = fmaf(v[4 * j + 0], w.x, acc); acc = fmaf(v[4 * j + 1], w.y, acc);
                      acc = fmaf(v[4 * j + 2], w.z, acc); acc = fmaf(v[4 * j + 3], w.w, acc);
                    }
                    logit[cl] += acc;
                  }
              }
            } else {
              NODE_TRACE(33, (g * 16 + c) | ((__float_as_uint(v[0] + v[31]) & 1u) << 20));
              rows_store(G.out, G.N, col0, v, G.tm);
              NODE_TRACE(34, g * 16 + c);
            }
            if (valid) {
              float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;   // four chains (one chain of 32 adds was 3.7 % of the stall samples)
#pragma unroll
              for (int j = 0; j < kTsCols; j += 4) { c0 += v[j]; c1 += v[j + 1]; c2 += v[j + 2]; c3 += v[j + 3]; }
              checksum += (c0 + c1) + (c2 + c3);
            }
          }
          NODE_TRACE(31, g * 16 + c);
          if (becomes_a) publish(d_col + (uint32_t)col0, v, G.kind == kNodeChainLast ? kF16ActScale : 1.0f);
          NODE_TRACE(30, g * 16 + c);
          if (G.signal_d_free) {   // every column group that ends inside this chunk is drained now
#pragma unroll
            for (int h = 0; h < kTsNSplit; ++h)
              if ((h + 1) * nh > c * 64 && (h + 1) * nh <= (c + 1) * 64) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars.d_free(h));
              }
          }
        }
        ++uses[dbuf];
        if (becomes_a) abuf ^= 1u;
#endif
      }

      if (a.do_readout) {
#pragma unroll
        for (int cl = 0; cl < kNodeMaxClasses; ++cl)
          if (cl < a.C) s_logits[(part * kTileM + row) * kNodeMaxClasses + cl] = logit[cl];
        ts_epilogue_barrier();
        if (part == 0 && valid) {
          // MASK logit -> -inf (score_network.py:183-185)
          for (int c = 0; c < a.C; ++c) {
            float v = Wp[a.readout_b + c];
#pragma unroll
            for (int pp = 0; pp < kTsParts; ++pp) v += s_logits[(pp * kTileM + row) * kNodeMaxClasses + c];
            if (a.mask_logit && c == a.C - 1) v = -CUDART_INF_F;
            a.logits_out[node * a.C + c] = v;
          }
          // S^alpha = z . Gamma^alpha . zhat   egnn_score_network.py:276-281
          for (int al = 0; al < 3; ++al) {
            float s = 0.f;
            for (int i = 0; i < D; ++i) {
              float inner = 0.f;
              for (int j = 0; j < D; ++j) inner += Wp[a.proj + (al * D + i) * D + j] * a.xcur[node * D + j];
              s += a.z[node * D + i] * inner;
            }
            a.score_out[node * 3 + al] = s;
          }
        }
        ts_epilogue_barrier();
      }
    }
#if B200MD_NODE_TS_PIPELINED
    if (part == 0) sfu_acquire(tok_mine);   // the peer's last release
#endif
#if B200MD_NODE_TS_TMA_STORE
    if (lane == 0) bulk_wait_all();   // the staging tiles have been read and the rows written before the CTA goes away
#endif
    if (!(fabsf(checksum) <= CUDART_MAX_NORMAL_F)) local_status |= B200MD_STATUS_TC_RANGE;
    if (local_status && a.status) atomicOr(a.status, (int)local_status);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base_all, kTmemCols);
  }
}
