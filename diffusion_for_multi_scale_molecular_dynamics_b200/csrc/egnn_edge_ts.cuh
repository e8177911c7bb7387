// EGNN edge kernel, 3xFP16 with the A operand in TENSOR MEMORY (tcgen05.mma ".ts" form).  Included by egnn_tc.cu
// (inside its anonymous namespace; shares EdgeTcArgs and the tc_common.cuh wrappers).
//
// The epilogue converts an accumulator IN PLACE into the next layer's A operand.  Accumulator columns [64c + 16j, +16)
// of a row (fp32) become, for K chunk c and K step j, 8 columns of packed fp16 pairs `hi` followed by 8 columns `lo`
// -- exactly the (lane = row, column = k / 2) layout tcgen05.mma reads for a K = 16 step -- so no activation ever
// touches shared memory: no swizzled stores, no generic->async proxy fence, no A ring (its 64 KB go to a deeper weight
// ring), and the tensor pipe reads only B from shared memory.
//
// Because A(g+1) stays in TMEM for the whole GEMM, the MMA order is column-group-major: the accumulator is computed
// in kTsEdgeNSplit groups of N / kTsEdgeNSplit columns, the first group over all K chunks as the epilogue delivers
// them, the others back to back, and the epilogue of layer g + 1 starts on a group as soon as it is complete.  (One
// group is the default now; two and four were measured, see kTsEdgeNSplit below.)  At the end of a tile the columns of
// the last accumulator are reused the same way for the NEXT tile's first activation, so the tensor pipe never drains
// between tiles.
//   GEMM index n (running over tiles): D = buffer n & 1, A = buffer (n & 1) ^ 1 (2 x 256 columns = all of TMEM).
//
// Warps (10): 0 = weight producer, 1 = MMA issuer (converged, elected lane), 2-9 = epilogue: two per TMEM lane quarter,
// each converting 32 of the 64 columns of a chunk (few, fat epilogue threads: 168 registers each, 32 independent SiLU
// chains per thread for the SFU pipe).  (ts_common.cuh: B200MD_TS_WARPGROUPS for the 12-warp setmaxnreg layout.)
//
// The two CTAs of a cluster are ONE tcgen05 CTA pair by default (B200MD_TS_CTA_PAIR below): M = 256 MMAs issued by the
// leader, half of every weight part in each CTA's shared memory; warp 1 of the peer forwards barrier completions.
#pragma once

#include "ts_common.cuh"

// CTAs per thread-block cluster.  With 2, the CTAs of a pair walk the same weight stream.  Without B200MD_TS_CTA_PAIR each
// fetches half of every weight stage from L2 and multicasts it into the shared memory of both (the round-1 scheme); with it
// each keeps its half only.  Either way the L2 -> SM weight traffic is halved (117 -> 58 GB per launch at 4096 structures).
// Both CTAs run the same number of tiles (the last one may be empty).
#ifndef B200MD_TS_CLUSTER
#define B200MD_TS_CLUSTER 2
#endif
constexpr int kTsCluster = B200MD_TS_CLUSTER;
// SFU token (ts_common.cuh: sfu_acquire / sfu_release): the two epilogue warps of a scheduler take turns in the SiLU phase
// (32 ex2 + 32 rcp per thread, SFU-bound) so that one converts / stores while the other occupies the SFU.  Cycle trace:
// SiLU phase 1050 -> 750 cycles, layer period 8390 -> 8230; 2 % on the power-capped step.  The token is handed on after
// B200MD_TS_TOKEN_RELEASE + 1 of the kTsCols / 4 column groups (earlier hand-over measured the same).
#ifndef B200MD_TS_SFU_TOKEN
#define B200MD_TS_SFU_TOKEN 1
#endif
#ifndef B200MD_TS_TOKEN_RELEASE
#define B200MD_TS_TOKEN_RELEASE 7
#endif
constexpr int kTsTokenReleaseAfter = B200MD_TS_TOKEN_RELEASE;

// Reciprocals of the SiLU (ts_common.cuh: silu4): 0 = one per element, 1 = one per pair, 2 = one per four elements
#ifndef B200MD_TS_EDGE_SILU
#define B200MD_TS_EDGE_SILU 1
#endif
constexpr int kTsEdgeSilu = B200MD_TS_EDGE_SILU;
// Column groups and weight ring of the edge kernel (the node kernel keeps ts_common.cuh's two groups and 12 stages).
// One group = whole-width MMAs (N = 256): 139 cycles each against 2 x 75 for two halves, half as many instructions
// and weight stages; it lost to two groups while the epilogue was the bound (20.4 vs 20.15 ms) and wins now that the
// epilogue is shorter (A/B on one box: 19.03 -> 18.42 ms per launch).  The ring leaves room for the staging tiles of
// the segmented sum (segment_add below): 32 rows x kTsCols columns of fp32 per epilogue warp.
#ifndef B200MD_TS_SEG_SMEM
#define B200MD_TS_SEG_SMEM 1
#endif
#ifndef B200MD_TS_EDGE_NSPLIT
#define B200MD_TS_EDGE_NSPLIT 1
#endif
constexpr int kTsEdgeNSplit = B200MD_TS_EDGE_NSPLIT;
// B200MD_TS_CTA_PAIR = 1: the two CTAs of a cluster work as ONE tcgen05 CTA pair (cta_group::2): the leader (rank 0)
// issues M = 256 MMAs over both CTAs' rows, every CTA keeps only ITS half of the output columns of a weight part in shared
// memory (16 KB per stage instead of 32; no multicast), and the tensor core of each SM reads half of B from the peer.
// The peer's otherwise idle MMA warp forwards its "stage full" / "A chunk ready" barrier completions to the leader.
#ifndef B200MD_TS_CTA_PAIR
#define B200MD_TS_CTA_PAIR 1
#endif
constexpr bool kTsPair = B200MD_TS_CTA_PAIR != 0;
static_assert(!kTsPair || (B200MD_TS_CLUSTER == 2 && B200MD_TS_EDGE_NSPLIT == 1), "a CTA pair is a cluster of 2 computing whole-width accumulators");
constexpr int kTsEdgeBStageBytes = (kMaxN / kTsEdgeNSplit) * kRowBytes / (kTsPair ? 2 : 1);
// B200MD_TS_PQ_STAGE = 1: the P[src] / Q[dst] rows of the NEXT tile's first activation travel L2 -> shared memory with
// cp.async (LDGSTS: no registers, no L1 lines held while in flight) into a private 2 x 128-byte slot per epilogue thread,
// one chunk ahead of their use: chunk 0 is requested when the tile's last layer starts, chunk c + 1 when chunk c has been
// read.  A thread reads only what it copied itself (cp.async.wait_all, no barrier).  The Q slots alias the staging tiles
// of the segmented sum (idle between the message layer and the next tile's message layer); the P slots are paid for by
// one weight stage (ring of 4).  Replaces the exposed L2 round trips of the gather at the tile boundary (cycle trace:
// 19.6 k cycles for the last layer against 8.9 k for a plain one).
#ifndef B200MD_TS_PQ_STAGE
#define B200MD_TS_PQ_STAGE 0
#endif
// B200MD_TS_PQ_GATHER = 1: two extra GATHER WARPS (no registers of the epilogue, no L1 lines) bring the P[src] / Q[dst] rows of
// the NEXT tile's first activation L2 -> shared memory with warp-coalesced cp.async (16 lanes x 16 bytes = one 256-byte row
// piece, two pieces per instruction) into a ring of two chunk stages ([row][64 columns], rows padded to 272 bytes so that
// the row-per-lane reads of the epilogue are conflict-free), completion on an mbarrier (cp.async.mbarrier.arrive.noinc).
// Chunks 0 and 1 are requested as soon as the tile's message layer is over (several layers before their use), chunk c + 2
// when the eight epilogue warps have read chunk c.  Stage 1 of the Q rows aliases the staging tiles of the segmented sum
// (idle from the message layer of a tile to that of the next one; the "message layer done" barrier hands them over).
// Paid for by the weight ring (4 stages of 16 KB; its depth was measured not to matter).
#ifndef B200MD_TS_PQ_GATHER
#define B200MD_TS_PQ_GATHER 0
#endif
#if !defined(B200MD_TS_EDGE_STAGES) && B200MD_TS_CTA_PAIR
#define B200MD_TS_EDGE_STAGES (B200MD_TS_PQ_GATHER ? 4 : (B200MD_TS_PQ_STAGE ? 8 : 10))
#endif
#ifndef B200MD_TS_NEXT_EDGE_AHEAD
#define B200MD_TS_NEXT_EDGE_AHEAD 0
#endif
#ifndef B200MD_TS_EDGE_STAGES
#define B200MD_TS_EDGE_STAGES (((B200MD_TS_SEG_SMEM ? 11 : 12) * B200MD_TS_EDGE_NSPLIT / 2) - (B200MD_TS_PQ_STAGE ? B200MD_TS_EDGE_NSPLIT : 0))
#endif
constexpr int kTsEdgeBStages = B200MD_TS_EDGE_STAGES;
constexpr int kTsSegTileBytes = 32 * kTsCols * 4;
static_assert(!B200MD_TS_SEG_SMEM || kTsCols == 32, "segment_add maps one lane to one column");
static_assert(!B200MD_TS_PQ_STAGE || (B200MD_TS_SEG_SMEM && kTsCols == 32), "the Q slots alias the segmented-sum tiles (32 x 128 B per warp)");
constexpr bool kTsGather = B200MD_TS_PQ_GATHER != 0;
static_assert(!kTsGather || (B200MD_TS_SEG_SMEM && !B200MD_TS_PQ_STAGE && B200MD_TS_CTA_PAIR), "gather warps: CTA-pair ring, Q stage 1 aliases the segmented-sum tiles");
// the gather warps: the two idle warps of the control warpgroup (B200MD_TS_WARPGROUPS) or two warps behind the epilogue warps
// (a 10-warp block is handed the registers of 12 anyway)
constexpr int kTsGatherWarp0 = B200MD_TS_WARPGROUPS ? 2 : kTsFirstEpiWarp + kTsEpiWarps;
constexpr int kTsEdgeThreads = kTsThreads + ((kTsGather && !B200MD_TS_WARPGROUPS) ? 64 : 0);
constexpr int kTsPqRowBytes = 64 * 4 + 16;                 // one chunk of a row + 16 bytes: row stride = 17 x 16 bytes
constexpr int kTsPqStageBytes = kTileM * kTsPqRowBytes;    // 34 KB
struct TsSmem {
  uint32_t b, seg, qx, q0, pst, pslot, bars, tmem_slot, bias, scale, wr, wlast, dot, total;
};
__host__ __device__ constexpr TsSmem ts_smem(int n_gemms) {
  TsSmem L{};
  uint32_t o = 0;
  L.b = o; o += kTsEdgeBStages * kTsEdgeBStageBytes;
  L.seg = o; o += B200MD_TS_SEG_SMEM ? kTsEpiWarps * kTsSegTileBytes : 0;
  // gather stages: Q stage 1 = [seg tiles | qx] (contiguous), Q stage 0, P stages 0 and 1
  L.qx = o; o += kTsGather ? kTsPqStageBytes - kTsEpiWarps * kTsSegTileBytes : 0;
  L.q0 = o; o += kTsGather ? kTsPqStageBytes : 0;
  L.pst = o; o += kTsGather ? 2 * kTsPqStageBytes : 0;
  L.pslot = o; o += B200MD_TS_PQ_STAGE ? kTsEpiWarps * kTsSegTileBytes : 0;
  L.bars = o; o += (2 * kTsEdgeBStages + kTsSlots * kTsParts + 2 * kTsEdgeNSplit + 5) * 8;
  o = (o + 15u) & ~15u;   // what follows is read with 128-bit loads
  L.tmem_slot = o; o += 16;
  L.bias = o; o += (uint32_t)n_gemms * kMaxN * 4;
  L.scale = o; o += 16 * 4;   // 2^-s of each GEMM's fp16 weight image (tc_common.cuh)
  L.wr = o; o += kMaxN * 4;
  L.wlast = o; o += kMaxN * 4;
  L.dot = o; o += 2 * kTsParts * kTileM * 4;   // [tile parity][part][row]
  L.total = o;
  return L;
}
static_assert(ts_smem(kMaxGemms).total + 1024 <= 227 * 1024, "edge kernel: shared memory (weight ring + segmented-sum tiles + gather ring) beyond 227 KB");


// Optional cycle trace of CTA 0 (build with -DB200MD_TS_TRACE; read back with b200md_debug_ts_trace): one record
// (warp, event, clock) per TS_TRACE call of lane 0 of the MMA warp and of the two epilogue warps of lane quarter 0.
#ifdef B200MD_TS_TRACE
constexpr int kTsTraceMax = 3 * 2048;     // 3 traced warps x 2048 records, each warp fills its own third (no atomics)
__device__ uint4 g_ts_trace[kTsTraceMax];
__device__ unsigned g_ts_trace_count;
#define TS_TRACE(ev, arg)                                                                      \
  do {                                                                                         \
    if (trace_on && trace_idx < 2048u) {                                                       \
      g_ts_trace[trace_slot * 2048u + trace_idx] = make_uint4((unsigned)warp, (unsigned)(ev), (unsigned)(arg), (unsigned)clock()); \
      ++trace_idx;                                                                             \
    }                                                                                          \
  } while (0)
#else
#define TS_TRACE(ev, arg) do { } while (0)
#endif

struct TsPipe {
  uint32_t s_base, bar0, tmem_base;
  __device__ __forceinline__ uint32_t b_full(uint32_t s) const { return bar0 + 8u * s; }
  __device__ __forceinline__ uint32_t b_empty(uint32_t s) const { return bar0 + 8u * (kTsEdgeBStages + s); }
  // "columns [part * kTsCols, +kTsCols) of chunk slot `slot` are in TMEM": the 4 lane-quarter warps of that part arrive
  __device__ __forceinline__ uint32_t a_full(uint32_t slot, uint32_t part) const {
    return bar0 + 8u * (2 * kTsEdgeBStages + slot * kTsParts + part);
  }
  // "column group `group` of accumulator buffer `buf` is complete" (groups complete in order)
  __device__ __forceinline__ uint32_t acc_full(uint32_t buf, uint32_t group) const {
    return bar0 + 8u * (2 * kTsEdgeBStages + kTsSlots * kTsParts + buf * kTsEdgeNSplit + group);
  }
  // gather ring (B200MD_TS_PQ_GATHER): "stage s holds a chunk" (the 64 gather lanes' copies have landed), "stage s has
  // been read" (8 epilogue warps), "the message layer of this tile is over" (8 epilogue warps: the seg tiles are free)
  __device__ __forceinline__ uint32_t pq_full(uint32_t s) const {
    return bar0 + 8u * (2 * kTsEdgeBStages + kTsSlots * kTsParts + 2 * kTsEdgeNSplit + s);
  }
  __device__ __forceinline__ uint32_t pq_empty(uint32_t s) const { return pq_full(s) + 16u; }
  __device__ __forceinline__ uint32_t msg_done() const { return pq_full(0) + 32u; }
};

// B200MD_TS_PQ_AHEAD = 1 requests the P / Q rows of the next tile one chunk ahead of their use (needs the registers of
// B200MD_TS_WARPGROUPS = 1); measured slower, see ts_common.cuh.
#ifndef B200MD_TS_PQ_AHEAD
#define B200MD_TS_PQ_AHEAD 0
#endif
__global__ void __launch_bounds__(kTsEdgeThreads, 1) egnn_edge_ts_kernel(const EdgeTcArgs a) {
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

  for (int t = threadIdx.x; t < p.n_gemms * kMaxN; t += kTsEdgeThreads) {
    const int g = t / kMaxN, n = t - g * kMaxN;
    s_bias[t] = (n < p.gN[g] && p.gB[g] >= 0) ? p.packed[p.gB[g] + n] * kF16ActScale : 0.f;   // x kF16ActScale: ts_common.cuh
  }
  for (int t = threadIdx.x; t < kMaxN; t += kTsEdgeThreads) {
    s_wr[t] = (t < M) ? p.packed[a.msg_wr + t] : 0.f;
    s_wlast[t] = (t < p.gN[p.n_gemms - 1]) ? p.packed[a.coord_last_w + t] : 0.f;
  }
  if (threadIdx.x < p.n_gemms)
    reinterpret_cast<float*>(smem + L.scale)[threadIdx.x] = p.wtc[p.gW[threadIdx.x] + weight_unscale_offset(p.gK[threadIdx.x], p.gN[threadIdx.x])];
  TsPipe pipe;
  pipe.s_base = smem_u32(smem);
  pipe.bar0 = pipe.s_base + L.bars;
  const int cta_rank = kTsCluster > 1 ? (int)cluster_ctarank() : 0;
  const bool pair_leader = kTsPair && cta_rank == 0;
  if (threadIdx.x == 0) {
    // CTA pair: the leader's "full" barriers also count the peer's forwarded completion; a stage is released by ONE
    // commit (the leader's, multicast) instead of one per CTA
    for (int s = 0; s < kTsEdgeBStages; ++s) { mbar_init(pipe.b_full(s), pair_leader ? 2 : 1); mbar_init(pipe.b_empty(s), kTsPair ? 1 : kTsCluster); }
    for (int s = 0; s < kTsSlots; ++s)
      for (int j = 0; j < kTsParts; ++j) mbar_init(pipe.a_full(s, j), pair_leader ? 5 : 4);
    for (int s = 0; s < 2 * kTsEdgeNSplit; ++s) mbar_init(pipe.acc_full(s / kTsEdgeNSplit, s % kTsEdgeNSplit), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(pipe.pq_full(s), 64); mbar_init(pipe.pq_empty(s), kTsEpiWarps); }
    mbar_init(pipe.msg_done(), kTsEpiWarps);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kTsPair) tmem_alloc_pair(pipe.s_base + L.tmem_slot, kTmemCols);
    else tmem_alloc(pipe.s_base + L.tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (kTsCluster > 1) cluster_sync();   // every CTA's barriers exist before a peer's multicast can reach them
  tc_fence_after();
  pipe.tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);

  // tile schedule: cluster k of n_clusters takes tiles (i n_clusters + k) kTsCluster + rank, i = 0, 1, ...; every CTA of
  // a cluster runs the same number of iterations (a tile index >= n_tiles is an empty tile)
  const int n_clusters = (int)gridDim.x / kTsCluster, cluster_id = (int)blockIdx.x / kTsCluster;
  const int n_groups = (n_tiles + kTsCluster - 1) / kTsCluster;
  const int n_iterations = n_groups > cluster_id ? (n_groups - cluster_id + n_clusters - 1) / n_clusters : 0;
  const int first_tile = cluster_id * kTsCluster + cta_rank, tile_stride = n_clusters * kTsCluster;
#ifdef B200MD_TS_TRACE
  bool trace_on = false;
  unsigned trace_idx = 0;
  const unsigned trace_slot = warp == 1 ? 0u : (warp == 4 ? 1u : 2u);
  if (blockIdx.x == 0 && threadIdx.x == 0) g_ts_trace_count = kTsTraceMax;
#endif

  // setmaxnreg sits at the top of each warpgroup's branch: ptxas budgets the code a setmaxnreg dominates (after a merge
  // of the two branches it would compile everything for the smaller value)
  if (warp < kTsFirstEpiWarp || warp >= kTsFirstEpiWarp + kTsEpiWarps) {
#if B200MD_TS_WARPGROUPS
  ts_regs_release<kTsRegsControl>();
#endif
  if (kTsGather && warp >= kTsGatherWarp0 && warp < kTsGatherWarp0 + 2) {
    // ===== gather warps: rows [64 gw, 64 gw + 64) of the NEXT tile's P[src] / Q[dst] chunks -> shared memory =====
    const int gw = warp - kTsGatherWarp0, nk0 = a.M / 64;
    const uint32_t p_st = pipe.s_base + L.pst, q_st0 = pipe.s_base + L.q0, q_st1 = pipe.s_base + L.seg;
    const uint32_t piece = (uint32_t)(lane & 15);
    uint32_t n = 0;   // running chunk index: stage n & 1
    if (a.n_msg > 0)
      for (int it = 0; it + 1 < n_iterations; ++it) {
        const int e0 = (first_tile + (it + 1) * tile_stride) * kTileM + 64 * gw + lane;
        const int s_lo = e0 < E ? a.edge_src[e0] : -1, s_hi = e0 + 32 < E ? a.edge_src[e0 + 32] : -1;
        const int d_lo = e0 < E ? a.edge_dst[e0] : 0, d_hi = e0 + 32 < E ? a.edge_dst[e0 + 32] : 0;
        mbar_wait(pipe.msg_done(), (uint32_t)it & 1u);   // the seg tiles (Q stage 1) are free
        for (int c = 0; c < nk0; ++c, ++n) {
          const uint32_t s = n & 1u;
          mbar_wait(pipe.pq_empty(s), ((n >> 1) & 1u) ^ 1u);
          const uint32_t p_row0 = p_st + s * kTsPqStageBytes + (uint32_t)(64 * gw) * kTsPqRowBytes + piece * 16u;
          const uint32_t q_row0 = (s ? q_st1 : q_st0) + (uint32_t)(64 * gw) * kTsPqRowBytes + piece * 16u;
          const float* pc = a.P + c * 64 + piece * 4;
          const float* qc = a.Q + c * 64 + piece * 4;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {   // two rows per instruction: lanes 0-15 row 2 i, lanes 16-31 row 2 i + 1
            const int r = 2 * i + (lane >> 4);
            const int src = __shfl_sync(0xffffffffu, i < 16 ? s_lo : s_hi, r & 31);
            const int dst = __shfl_sync(0xffffffffu, i < 16 ? d_lo : d_hi, r & 31);
            if (src >= 0) {
              cp_async16(p_row0 + (uint32_t)r * kTsPqRowBytes, pc + (int64_t)src * a.M);
              cp_async16(q_row0 + (uint32_t)r * kTsPqRowBytes, qc + (int64_t)dst * a.M);
            }
          }
          cp_async_mbar_arrive_noinc(pipe.pq_full(s));
        }
      }
  } else if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t cb = 0;
      for (int it = 0; it < n_iterations; ++it)
        for (int g = 0; g < p.n_gemms; ++g) {
          const int nk = p.gK[g] / 64, nh = p.gN[g] / kTsEdgeNSplit;
          const uint32_t bytes = (uint32_t)nh * kRowBytes;
          const int64_t part_floats = (int64_t)p.gN[g] * (kRowBytes / 4);
          const float* src = p.wtc + p.gW[g] + f16_image_offset(p.gK[g], p.gN[g]);
          for (int h = 0; h < kTsEdgeNSplit; ++h)
            for (int part = 0; part < 2 * nk; ++part, ++cb) {   // chunk 0 hi, chunk 0 lo, chunk 1 hi, ...
              const uint32_t s = cb % kTsEdgeBStages;
              mbar_wait(pipe.b_empty(s), ((cb / kTsEdgeBStages) & 1u) ^ 1u);
              mbar_arrive_expect_tx(pipe.b_full(s), kTsPair ? bytes / 2 : bytes);
              const float* stage_src = src + (int64_t)part * part_floats + (int64_t)h * nh * (kRowBytes / 4);
              if (kTsPair) {   // this CTA's half of the output columns, into its own shared memory only
                const uint32_t half = bytes / 2;
                bulk_g2s(pipe.s_base + L.b + s * kTsEdgeBStageBytes, stage_src + cta_rank * (half / 4), half, pipe.b_full(s));
              } else if (kTsCluster > 1) {   // this CTA's slice of the stage, delivered to every CTA of the cluster
                const uint32_t slice = bytes / kTsCluster;
                bulk_g2s_multicast(pipe.s_base + L.b + s * kTsEdgeBStageBytes + cta_rank * slice, stage_src + cta_rank * (slice / 4),
                                   slice, pipe.b_full(s), (uint16_t)((1u << kTsCluster) - 1u));
              } else {
                bulk_g2s(pipe.s_base + L.b + s * kTsEdgeBStageBytes, stage_src, bytes, pipe.b_full(s));
              }
            }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The WHOLE warp walks the loop (converged), so every address and descriptor lives in uniform registers; one
    // elected lane issues.  (With only lane 0 inside the loop the compiler wraps each tcgen05.mma in an
    // ELECT / R2UR / branch "waterfall": ~40 instructions per MMA, and the issuing thread -- not the tensor pipe --
    // paces the kernel.)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, pipe.tmem_base, 0);
    const uint32_t b_ring = pipe.s_base + L.b;
    const bool leader = elect_one();
    uint32_t cb = 0, mca = 0, mgc = 0;
    if (kTsPair && cta_rank != 0) {
      // ===== peer of a CTA pair: no MMAs; hand every local "full" completion on to the leader, in the leader's order =====
      for (int it = 0; it < n_iterations; ++it)
        for (int g = 0; g < p.n_gemms; ++g) {
          const int nk = p.gK[g] / 64;
          for (int c = 0; c < nk; ++c) {
            const uint32_t s1 = cb % kTsEdgeBStages, s2 = (cb + 1) % kTsEdgeBStages;
            mbar_wait(pipe.b_full(s1), (cb / kTsEdgeBStages) & 1u);
            if (leader) mbar_arrive_remote(pipe.b_full(s1), 0u);
            mbar_wait(pipe.b_full(s2), ((cb + 1) / kTsEdgeBStages) & 1u);
            if (leader) mbar_arrive_remote(pipe.b_full(s2), 0u);
            const uint32_t chunk = mca + (uint32_t)c;
#pragma unroll
            for (int j = 0; j < kTsParts; ++j) {
              mbar_wait(pipe.a_full(chunk % kTsSlots, j), (chunk / kTsSlots) & 1u);
              tc_fence_after();
              tc_fence_before();
              if (leader) mbar_arrive_remote(pipe.a_full(chunk % kTsSlots, j), 0u);
            }
            __syncwarp();
            cb += 2;
          }
          mca += nk;
        }
    } else
    for (int it = 0; it < n_iterations; ++it)
      for (int g = 0; g < p.n_gemms; ++g, ++mgc) {
#ifdef B200MD_TS_TRACE
        trace_on = (blockIdx.x == 0 && lane == 0 && it >= 3 && it < 5);
#endif
        const int nk = p.gK[g] / 64, nh = p.gN[g] / kTsEdgeNSplit;
        const uint32_t idesc = kTsPair ? make_idesc_pair(nh, 0u) : make_idesc(nh, 0u);
        const uint32_t d_col = (mgc & 1u) * kMaxN, a_col = d_col ^ kMaxN;
        for (int h = 0; h < kTsEdgeNSplit; ++h) {
          const uint32_t d_tmem = tmem_base + d_col + (uint32_t)(h * nh);
          for (int c = 0; c < nk; ++c) {
            const uint32_t s1 = cb % kTsEdgeBStages, s2 = (cb + 1) % kTsEdgeBStages;
            mbar_wait(pipe.b_full(s1), (cb / kTsEdgeBStages) & 1u);
            mbar_wait(pipe.b_full(s2), ((cb + 1) / kTsEdgeBStages) & 1u);
            const uint32_t bd1 = desc_lo_sw128(b_ring + s1 * kTsEdgeBStageBytes), bd2 = desc_lo_sw128(b_ring + s2 * kTsEdgeBStageBytes);
            const uint32_t chunk = mca + (uint32_t)c;
            const uint32_t a_chunk = tmem_base + a_col + (uint32_t)(c * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // K steps of 16
              if (h == 0 && (j % (4 / kTsParts)) == 0) {
                TS_TRACE(100 + j, g * 16 + c);
                mbar_wait(pipe.a_full(chunk % kTsSlots, j / (4 / kTsParts)), (chunk / kTsSlots) & 1u);
                TS_TRACE(110 + j, g * 16 + c);
              }
              tc_fence_after();
              if (leader) {
                umma_f16_ts<kTsPair>(d_tmem, a_chunk + j * 16, bd1 + 2 * j, idesc, (c | j) != 0);   // A_1 B_1
                umma_f16_ts<kTsPair>(d_tmem, a_chunk + j * 16 + 8, bd1 + 2 * j, idesc | kIdescA2, 1u);   // A_2 B_1 (A_2 is stored negated)
                umma_f16_ts<kTsPair>(d_tmem, a_chunk + j * 16, bd2 + 2 * j, idesc, 1u);             // A_1 B_2
              }
            }
            if (leader) {   // a stage is free when EVERY CTA of the cluster has consumed it: the peers write into it too
              if (kTsPair) {
                umma_commit_pair(pipe.b_empty(s1), 3u);
                umma_commit_pair(pipe.b_empty(s2), 3u);
              } else if (kTsCluster > 1) {
                umma_commit_multicast(pipe.b_empty(s1), (uint16_t)((1u << kTsCluster) - 1u));
                umma_commit_multicast(pipe.b_empty(s2), (uint16_t)((1u << kTsCluster) - 1u));
              } else {
                umma_commit(pipe.b_empty(s1));
                umma_commit(pipe.b_empty(s2));
              }
            }
            __syncwarp();
            cb += 2;
          }
          if (leader) {
            if (kTsPair) umma_commit_pair(pipe.acc_full(mgc & 1u, h), 3u);   // both CTAs' epilogues
            else umma_commit(pipe.acc_full(mgc & 1u, h));
          }
          __syncwarp();
          TS_TRACE(120 + h, g);
        }
        mca += nk;
      }
  }
  } else {
#if B200MD_TS_WARPGROUPS
    ts_regs_acquire<kTsRegsEpilogue>();
#endif
    // ===== epilogue warps =====
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int part = (warp - kTsFirstEpiWarp) >> 2;   // which kTsCols columns of a chunk this warp converts
    const int row = q * 32 + lane;
    const uint32_t lane_taddr = pipe.tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sa_bias = pipe.s_base + L.bias, sa_wr = pipe.s_base + L.wr, sa_wlast = pipe.s_base + L.wlast;
    uint32_t ca = 0, gc = 0;
    uint32_t dep_first = 0;   // sink of silu4's token dependency where no token is passed
    bool poisoned = false;
#if B200MD_TS_SFU_TOKEN
    const uint32_t tok_mine = 2u + 2u * (uint32_t)q + (uint32_t)part, tok_peer = 2u + 2u * (uint32_t)q + (uint32_t)(part ^ 1);
    if (part == 1) sfu_arrive(tok_peer);   // part 0 goes first
#endif
    // store the A-operand image of kTsCols columns at `col` of this lane and publish them
    auto publish = [&](uint32_t col, const uint32_t (&w)[kTsCols]) {
#pragma unroll
      for (int i = 0; i < kTsCols / 16; ++i) tmem_st16(lane_taddr + col + 16 * i, w + 16 * i);
      TS_TRACE(10, col);
      tmem_wait_st();
      TS_TRACE(11, col);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pipe.a_full(ca % kTsSlots, part));
      ++ca;
    };

    // per-tile state of this thread's edge (row of the tile)
    struct Edge {
      int src, dst, seg_start;
      unsigned tails;   // bit r: row r of this warp ends a segment (and is a real edge)
      bool valid, is_tail;
      float radial;
    };
    auto load_edge = [&](int it) {
      Edge ed;
      const int e = (first_tile + it * tile_stride) * kTileM + row;
      ed.valid = e < E;
      ed.src = ed.valid ? a.edge_src[e] : -1;
      ed.dst = ed.valid ? a.edge_dst[e] : -1;
      // segments of equal source index inside this warp's 32 rows (edges are sorted by source)
      const int src_prev = __shfl_up_sync(0xffffffffu, ed.src, 1);
      const int src_next = __shfl_down_sync(0xffffffffu, ed.src, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || ed.src != src_prev);
      ed.seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
      ed.is_tail = (lane == 31) || (ed.src != src_next);
      ed.tails = __ballot_sync(0xffffffffu, ed.is_tail && ed.valid);
      ed.radial = 0.f;   // coord2radial (models/egnn.py:230-238)
      if (ed.valid)
        for (int d = 0; d < a.D; ++d) {
          const float df = a.xcur[(int64_t)ed.src * a.D + d] - a.xcur[(int64_t)ed.dst * a.D + d];
          ed.radial += df * df;
        }
      return ed;
    };
    // target[src * ld + col0 + (0 .. kTsCols)] += sum of v over each run of equal src inside this warp's 32 rows.
#if B200MD_TS_SEG_SMEM
    // Transposed through shared memory: each thread stages its row (XOR-swizzled 16-byte groups, conflict-free both
    // ways), then lane j walks column j down the 32 rows and issues one coalesced red.global.add per segment tail
    // (warp-uniform branch on the tails mask).  8 st.shared.v4 + 32 ld.shared per thread instead of 160 shuffles and
    // 160 selects / adds (the shuffles use the same crossbar as shared memory): the message layer's period drops from
    // 21.3 k to 13.0 k cycles (other layers: 8.9 k) and the launch from 19.44 to 18.98 ms (A/B on one box).
    const uint32_t seg_tile = pipe.s_base + L.seg + (uint32_t)(warp - kTsFirstEpiWarp) * kTsSegTileBytes;
    const uint32_t seg_read = seg_tile + (uint32_t)(lane & 3) * 4u;
    auto segment_add = [&](const Edge& ed, const float (&v)[kTsCols], float* target, int ld, int col0) {
#pragma unroll
      for (int g = 0; g < kTsCols / 4; ++g) {
        const uint32_t at = seg_tile + (uint32_t)lane * (kTsCols * 4) + (uint32_t)((g ^ (lane & 7)) << 4);
        if (ed.valid) sts128(at, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        else sts128(at, 0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      // pass 1, branch-free: running sum of column `lane` down the rows, restarted behind every tail, written back in
      // place.  (A branch per row costs a BSSY / BSYNC pair: 1800 cycles per chunk, measured.  Masked per-segment sums
      // held in registers instead of the write-back spill and are slower.  What bounds this now is shared-memory
      // bandwidth: 100 wavefronts per warp and chunk next to the tensor pipe's weight reads, 1300 cycles per chunk.)
      float t[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) t[r] = lds32(seg_read + (uint32_t)r * (kTsCols * 4) + (uint32_t)(((lane >> 2) ^ (r & 7)) << 4));
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        acc = (r > 0 && (ed.tails & (1u << (r > 0 ? r - 1 : 0)))) ? t[r] : acc + t[r];
        sts32(seg_read + (uint32_t)r * (kTsCols * 4) + (uint32_t)(((lane >> 2) ^ (r & 7)) << 4), acc);
      }
      // pass 2: one coalesced reduction per tail (two or three per warp at 25 edges per atom)
      float* const column = target + col0 + lane;
      for (unsigned m = ed.tails; m != 0; m &= m - 1) {
        const int r = __ffs(m) - 1;
        const float sum = lds32(seg_read + (uint32_t)r * (kTsCols * 4) + (uint32_t)(((lane >> 2) ^ (r & 7)) << 4));
        const int s = __shfl_sync(0xffffffffu, ed.src, r);
        red_add_f32(column + (int64_t)s * ld, sum * (1.0f / kF16ActScale));   // v is kF16ActScale x the message
      }
      __syncwarp();
    };
#else
    // inclusive segmented scans over the warp's rows (interleaved, so the shuffle latencies overlap), then one vector
    // reduction per segment tail
    auto segment_add = [&](const Edge& ed, const float (&v)[kTsCols], float* target, int ld, int col0) {
      float* target_row = target + (int64_t)ed.src * ld + col0;
      float sum[kTsCols];
#pragma unroll
      for (int j = 0; j < kTsCols; ++j) sum[j] = ed.valid ? v[j] : 0.f;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const bool take = lane - off >= ed.seg_start;
#pragma unroll
        for (int j = 0; j < kTsCols; ++j) {
          const float t = __shfl_up_sync(0xffffffffu, sum[j], off);
          sum[j] += take ? t : 0.f;
        }
      }
      if (ed.is_tail && ed.valid) {
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j)
          red_add_v4(target_row + 4 * j, sum[4 * j] * (1.0f / kF16ActScale), sum[4 * j + 1] * (1.0f / kF16ActScale),
                     sum[4 * j + 2] * (1.0f / kF16ActScale), sum[4 * j + 3] * (1.0f / kF16ActScale));
      }
    };
#endif
    // chunk c of the first activation of a tile: silu(P[src] + Q[dst] + radial * w_r) -> A operand at a_col
    auto prefetch_first = [&](const Edge& ed, int c) {
#ifdef B200MD_TS_PREFETCH   // off: with 227 KB of shared memory in use there is no L1 left to prefetch into (A/B: -1 %)
      if (ed.valid) {
        const float* pp = a.P + (int64_t)ed.src * M + c * 64 + part * kTsCols;
        const float* qq = a.Q + (int64_t)ed.dst * M + c * 64 + part * kTsCols;
        prefetch_l1(pp);
        prefetch_l1(qq);
      }
#endif
    };
    // The gathered rows P[src], Q[dst] of one chunk (kTsCols columns each).  At the tile boundary they are requested one
    // chunk ahead of their use (load_pq), so the L2 round trip of the gather runs under the SiLU / dot-product work
    // of the current tile's last layer instead of in front of the next tile's first MMA (the cycle trace showed 19.6 k
    // cycles for that layer against 8.9 k for a plain one).
    struct PQ {
      float4 p[kTsCols / 4], q[kTsCols / 4];
    };
    const uint64_t act2 = pack2(kF16ActScale, kF16ActScale);   // the first activation is scaled like every other A operand
    auto load_pq = [&](const Edge& ed, int c, PQ& t) {
      if (ed.valid) {
        const int col0 = c * 64 + part * kTsCols;
        const float4* p4 = reinterpret_cast<const float4*>(a.P + (int64_t)ed.src * M + col0);
        const float4* q4 = reinterpret_cast<const float4*>(a.Q + (int64_t)ed.dst * M + col0);
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) { t.p[j] = __ldg(p4 + j); t.q[j] = __ldg(q4 + j); }
      }
    };
    auto first_activation_from = [&](const Edge& ed, int c, uint32_t a_col, const PQ& t) {
      const int col0 = c * 64 + part * kTsCols;
      float v[kTsCols];
      if (ed.valid) {
        const uint64_t rad2 = pack2(ed.radial, ed.radial);
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) {
          const float4 pv = t.p[j], qv = t.q[j], wv = lds128(sa_wr + (uint32_t)(col0 + 4 * j) * 4u);
          silu4<kTsEdgeSilu>(mul2(fma2(rad2, pack2(wv.x, wv.y), add2(pack2(pv.x, pv.y), pack2(qv.x, qv.y))), act2), mul2(fma2(rad2, pack2(wv.z, wv.w), add2(pack2(pv.z, pv.w), pack2(qv.z, qv.w))), act2),
                             v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dep_first);
        }
      } else {
#pragma unroll
        for (int j = 0; j < kTsCols; ++j) v[j] = 0.f;
      }
      if (a.n_msg == 0) segment_add(ed, v, a.agg, M, col0);  // no hidden layer: this IS the message
      uint32_t w[kTsCols];
      split_hi_lo(v, w);
      publish(a_col + (uint32_t)col0, w);
    };
    auto first_activation_chunk = [&](const Edge& ed, int c, uint32_t a_col) {
      const int col0 = c * 64 + part * kTsCols;
      float v[kTsCols];
      TS_TRACE(20, c);
      if (ed.valid) {
        const uint64_t rad2 = pack2(ed.radial, ed.radial);
        const float4* p4 = reinterpret_cast<const float4*>(a.P + (int64_t)ed.src * M + col0);
        const float4* q4 = reinterpret_cast<const float4*>(a.Q + (int64_t)ed.dst * M + col0);
#ifdef B200MD_TS_TRACE
        float4 pvs[kTsCols / 4], qvs[kTsCols / 4];
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) { pvs[j] = __ldg(p4 + j); qvs[j] = __ldg(q4 + j); }
        float probe = 0.f;
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) probe += pvs[j].x + qvs[j].w;
        TS_TRACE(24, __float_as_uint(probe) & 1u);
#endif
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) {
#ifdef B200MD_TS_TRACE
          const float4 pv = pvs[j], qv = qvs[j], wv = lds128(sa_wr + (uint32_t)(col0 + 4 * j) * 4u);
#else
          const float4 pv = __ldg(p4 + j), qv = __ldg(q4 + j), wv = lds128(sa_wr + (uint32_t)(col0 + 4 * j) * 4u);
#endif
          // silu((P + Q) + radial * w_r), two columns per instruction
          silu4<kTsEdgeSilu>(mul2(fma2(rad2, pack2(wv.x, wv.y), add2(pack2(pv.x, pv.y), pack2(qv.x, qv.y))), act2), mul2(fma2(rad2, pack2(wv.z, wv.w), add2(pack2(pv.z, pv.w), pack2(qv.z, qv.w))), act2),
                             v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dep_first);
        }
      } else {
#pragma unroll
        for (int j = 0; j < kTsCols; ++j) v[j] = 0.f;
      }
      if (a.n_msg == 0) segment_add(ed, v, a.agg, M, col0);  // no hidden layer: this IS the message
      TS_TRACE(21, __float_as_uint(v[0] + v[kTsCols - 1]) & 1u);
      uint32_t w[kTsCols];
      split_hi_lo(v, w);
      TS_TRACE(22, (w[0] ^ w[kTsCols - 1]) & 1u);
      publish(a_col + (uint32_t)col0, w);
    };

#if B200MD_TS_PQ_STAGE
    const int nk0_pq = M / 64;
    // this thread's private slots: 8 x 16 bytes each, XOR-swizzled so that a warp-wide 128-bit access is conflict-free
    const uint32_t q_slot = pipe.s_base + L.seg + (uint32_t)(warp - kTsFirstEpiWarp) * kTsSegTileBytes + (uint32_t)lane * 128u;
    const uint32_t p_slot = pipe.s_base + L.pslot + (uint32_t)(warp - kTsFirstEpiWarp) * kTsSegTileBytes + (uint32_t)lane * 128u;
    const bool stage_pq = a.n_msg > 0;   // without a hidden message layer segment_add runs inside the first activation
    auto stage_request = [&](const Edge& ed, int c) {
      if (ed.valid) {
        const int col0 = c * 64 + part * kTsCols;
        const float* pp = a.P + (int64_t)ed.src * M + col0;
        const float* qq = a.Q + (int64_t)ed.dst * M + col0;
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) {
          const uint32_t piece = (uint32_t)((j ^ (lane & 7)) << 4);
          cp_async16(p_slot + piece, pp + 4 * j);
          cp_async16(q_slot + piece, qq + 4 * j);
        }
      }
      cp_async_commit();
    };
    // chunk c of the next tile's first activation from the slots; requests chunk c + 1 once the slots have been read
    auto first_activation_staged = [&](const Edge& ed, int c, uint32_t a_col) {
      const int col0 = c * 64 + part * kTsCols;
      float v[kTsCols];
      TS_TRACE(20, c);
      cp_async_wait_all();
      TS_TRACE(25, c);
      if (ed.valid) {
        const uint64_t rad2 = pack2(ed.radial, ed.radial);
        uint64_t s01[kTsCols / 4], s23[kTsCols / 4];
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) {
          const uint32_t piece = (uint32_t)((j ^ (lane & 7)) << 4);
          const float4 pv = lds128_ordered(p_slot + piece), qv = lds128_ordered(q_slot + piece);
          s01[j] = add2(pack2(pv.x, pv.y), pack2(qv.x, qv.y));
          s23[j] = add2(pack2(pv.z, pv.w), pack2(qv.z, qv.w));
        }
        TS_TRACE(24, (uint32_t)(s01[0] ^ s23[kTsCols / 4 - 1]) & 1u);
        if (c + 1 < nk0_pq) stage_request(ed, c + 1);
        TS_TRACE(26, c);
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) {
          const float4 wv = lds128(sa_wr + (uint32_t)(col0 + 4 * j) * 4u);
          silu4<kTsEdgeSilu>(mul2(fma2(rad2, pack2(wv.x, wv.y), s01[j]), act2), mul2(fma2(rad2, pack2(wv.z, wv.w), s23[j]), act2),
                             v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dep_first);
        }
      } else {
#pragma unroll
        for (int j = 0; j < kTsCols; ++j) v[j] = 0.f;
      }
      TS_TRACE(21, __float_as_uint(v[0] + v[kTsCols - 1]) & 1u);
      uint32_t w[kTsCols];
      split_hi_lo(v, w);
      publish(a_col + (uint32_t)col0, w);
    };
#endif

#if B200MD_TS_PQ_GATHER
    // chunk c of the NEXT tile's first activation from the gather ring (its rows were requested by the gather warps long
    // before): wait for the stage, read this row's two 128-byte pieces, hand the stage back
    const bool gathered = a.n_msg > 0;   // without a hidden message layer segment_add runs inside the first activation
    uint32_t pqn = 0;
    const uint32_t pq_row = (uint32_t)row * kTsPqRowBytes + (uint32_t)part * (kTsCols * 4);
    auto first_activation_gathered = [&](const Edge& ed, int c, uint32_t a_col) {
      const int col0 = c * 64 + part * kTsCols;
      const uint32_t s = pqn & 1u;
      float v[kTsCols];
      TS_TRACE(20, c);
      mbar_wait(pipe.pq_full(s), (pqn >> 1) & 1u);
      TS_TRACE(25, c);
      if (ed.valid) {
        const uint64_t rad2 = pack2(ed.radial, ed.radial);
        const uint32_t p_at = pipe.s_base + L.pst + s * kTsPqStageBytes + pq_row;
        const uint32_t q_at = pipe.s_base + (s ? L.seg : L.q0) + pq_row;
#pragma unroll
        for (int j = 0; j < kTsCols / 4; ++j) {
          const float4 pv = lds128_ordered(p_at + 16u * j), qv = lds128_ordered(q_at + 16u * j), wv = lds128(sa_wr + (uint32_t)(col0 + 4 * j) * 4u);
          silu4<kTsEdgeSilu>(mul2(fma2(rad2, pack2(wv.x, wv.y), add2(pack2(pv.x, pv.y), pack2(qv.x, qv.y))), act2), mul2(fma2(rad2, pack2(wv.z, wv.w), add2(pack2(pv.z, pv.w), pack2(qv.z, qv.w))), act2),
                             v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dep_first);
        }
      } else {
#pragma unroll
        for (int j = 0; j < kTsCols; ++j) v[j] = 0.f;
      }
      TS_TRACE(21, __float_as_uint(v[0] + v[kTsCols - 1]) & 1u);
      uint32_t w[kTsCols];
      split_hi_lo(v, w);   // every v depends on the stage's reads: they have completed by the time the stage is handed back
      __syncwarp();
      if (lane == 0) mbar_arrive(pipe.pq_empty(s));
      ++pqn;
      publish(a_col + (uint32_t)col0, w);
    };
#endif

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
    const int g_next_edge = p.n_gemms - 1 - B200MD_TS_NEXT_EDGE_AHEAD > 0 ? p.n_gemms - 1 - B200MD_TS_NEXT_EDGE_AHEAD : 0;
    for (int it = 0; it < n_iterations; ++it) {
#ifdef B200MD_TS_TRACE
      trace_on = (blockIdx.x == 0 && lane == 0 && q == 0 && it >= 3 && it < 5);
#endif
      // ---- the chain of Linear layers ----
      float dot = 0.f;
      Edge next{};
      const bool has_next = it + 1 < n_iterations;
      for (int g = 0; g < p.n_gemms; ++g, ++gc) {
        const bool is_message = (g == a.n_msg - 1);
        const bool is_last = (g == p.n_gemms - 1);
        const uint32_t d_col = (gc & 1u) * kMaxN;
        const int nchunks = p.gN[g] / 64, nh = p.gN[g] / kTsEdgeNSplit;
        const uint32_t bias = sa_bias + (uint32_t)(g * kMaxN) * 4u;
        // accumulator = 2^s (weights) x kF16ActScale (A operand) x product; kF16ActScale x pre-activation = acc 2^-s + bias'
        const float unscale = lds32(pipe.s_base + L.scale + (uint32_t)g * 4u);
        const uint64_t unscale2 = pack2(unscale, unscale);
        // the next tile's edge (indices, radial: two dependent round trips) is requested B200MD_TS_NEXT_EDGE_AHEAD layers
        // before the last one, so that it is in registers when the staged copies of its P / Q rows are issued
        if (g == g_next_edge && has_next) {
          next = load_edge(it + 1);
          prefetch_first(next, 0);
        }
#if B200MD_TS_PQ_STAGE
        if (is_last && has_next && stage_pq) stage_request(next, 0);
#endif
        int waited = -1;
        auto wait_for_chunk = [&](int c) {
          const int need = ts_last_group_of_chunk<kTsEdgeNSplit>(c, nh);
          if (need > waited) {
            TS_TRACE(1, g * 16 + c);
            mbar_wait(pipe.acc_full(gc & 1u, need), (gc >> 1) & 1u);
            tc_fence_after();
            waited = need;
            TS_TRACE(2, g * 16 + c);
          }
        };
        // v = kF16ActScale x silu(pre-activation), with bias' = kF16ActScale x bias in shared memory
        auto activate = [&](const uint32_t (&r)[kTsCols], int col0, float (&v)[kTsCols]) {
          TS_TRACE(4, col0);
#if B200MD_TS_SFU_TOKEN
          uint32_t dep = 0;
#pragma unroll
          for (int j = 0; j < kTsCols / 4; ++j) {
            const float4 b = lds128(bias + (uint32_t)(col0 + 4 * j) * 4u);
            silu4<kTsEdgeSilu>(fma2(pack2(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])), unscale2, pack2(b.x, b.y)),
                               fma2(pack2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), unscale2, pack2(b.z, b.w)),
                               v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dep);
            if (j == kTsTokenReleaseAfter) sfu_release(tok_peer, dep);
          }
#else
#pragma unroll
          for (int j = 0; j < kTsCols / 4; ++j) {
            const float4 b = lds128(bias + (uint32_t)(col0 + 4 * j) * 4u);
            silu4<kTsEdgeSilu>(fma2(pack2(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])), unscale2, pack2(b.x, b.y)),
                               fma2(pack2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), unscale2, pack2(b.z, b.w)),
                               v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dep_first);
          }
#endif
          TS_TRACE(5, col0);
        };
        // accumulator columns -> registers; with the SFU token, the turn is taken before the registers become usable
        auto load_acc = [&](uint32_t taddr, uint32_t (&r)[kTsCols]) {
#pragma unroll
          for (int i = 0; i < kTsCols / 16; ++i) tmem_ld16_issue(taddr + 16 * i, r + 16 * i);
#if B200MD_TS_SFU_TOKEN
          sfu_acquire(tok_mine);
#endif
#pragma unroll
          for (int i = 0; i < kTsCols / 16; ++i) tmem_ld_wait16(r + 16 * i);
        };
        if (!is_last) {
          float v[kTsCols];
          uint32_t r[kTsCols];
          wait_for_chunk(0);
          load_acc(lane_taddr + d_col + part * kTsCols, r);
          activate(r, part * kTsCols, v);
          for (int c = 0; c < nchunks; ++c) {
            const int col0 = c * 64 + part * kTsCols;
            const bool more = c + 1 < nchunks;
            if (more) {   // the next chunk's accumulator columns travel while this chunk is split and published
              wait_for_chunk(c + 1);
              TS_TRACE(14, g * 16 + c);
#pragma unroll
              for (int i = 0; i < kTsCols / 16; ++i) tmem_ld16_issue(lane_taddr + d_col + col0 + 64 + 16 * i, r + 16 * i);
              TS_TRACE(15, g * 16 + c);
            }
            uint32_t w[kTsCols];
            split_hi_lo(v, w);
            TS_TRACE(9, g * 16 + c);
            publish(d_col + (uint32_t)col0, w);
            TS_TRACE(6, g * 16 + c);
            if (is_message) {
              TS_TRACE(16, g * 16 + c);
              segment_add(cur, v, a.agg, M, col0);  // agg[src] += m (models/egnn.py:178-181)
              TS_TRACE(17, g * 16 + c);
#if B200MD_TS_PQ_GATHER
              if (!more && lane == 0) mbar_arrive(pipe.msg_done());   // (segment_add ends with a __syncwarp) the seg tiles are free
#endif
            }
            if (more) {
              TS_TRACE(12, g * 16 + c);
#if B200MD_TS_SFU_TOKEN
              sfu_acquire(tok_mine);
#endif
              TS_TRACE(13, g * 16 + c);
#pragma unroll
              for (int i = 0; i < kTsCols / 16; ++i) tmem_ld_wait16(r + 16 * i);
              activate(r, col0 + 64, v);
            }
          }
        } else {
#if B200MD_TS_PQ_AHEAD
          PQ ahead;
          if (has_next) load_pq(next, 0, ahead);
#endif
          for (int c = 0; c < nchunks; ++c) {
            wait_for_chunk(c);
            const int col0 = c * 64 + part * kTsCols;
            float v[kTsCols];
            uint32_t r[kTsCols];
            load_acc(lane_taddr + d_col + col0, r);
            activate(r, col0, v);
            // last coordinate layer: Linear(Cx -> 1, no bias)
#pragma unroll
            for (int j = 0; j < kTsCols / 4; ++j) {
              const float4 w = lds128(sa_wlast + (uint32_t)(col0 + 4 * j) * 4u);
              dot = fmaf(v[4 * j + 0], w.x, dot);
              dot = fmaf(v[4 * j + 1], w.y, dot);
              dot = fmaf(v[4 * j + 2], w.z, dot);
              dot = fmaf(v[4 * j + 3], w.w, dot);
            }
            TS_TRACE(23, __float_as_uint(dot) & 1u);
            // these accumulator columns are free now: the next tile's first activation goes there IN PLACE, so its
            // first GEMM starts while this tile is being finished
            if (has_next && c < nk0) {
#if B200MD_TS_PQ_AHEAD
              first_activation_from(next, c, d_col, ahead);
              if (c + 1 < nk0) load_pq(next, c + 1, ahead);   // used after the next chunk's SiLU and dot product
#elif B200MD_TS_PQ_STAGE
              if (stage_pq) first_activation_staged(next, c, d_col);
              else first_activation_chunk(next, c, d_col);
#elif B200MD_TS_PQ_GATHER
              if (gathered) first_activation_gathered(next, c, d_col);
              else first_activation_chunk(next, c, d_col);
#else
              if (c + 1 < nk0) prefetch_first(next, c + 1);
              first_activation_chunk(next, c, d_col);
#endif
            }
          }
          if (has_next)
            for (int c = n_last; c < nk0; ++c) {   // message width > coordinate width: columns no accumulator used
#if B200MD_TS_PQ_AHEAD
              first_activation_from(next, c, d_col, ahead);
              if (c + 1 < nk0) load_pq(next, c + 1, ahead);
#elif B200MD_TS_PQ_STAGE
              if (stage_pq) first_activation_staged(next, c, d_col);
              else first_activation_chunk(next, c, d_col);
#elif B200MD_TS_PQ_GATHER
              if (gathered) first_activation_gathered(next, c, d_col);
              else first_activation_chunk(next, c, d_col);
#else
              if (c + 1 < nk0) prefetch_first(next, c + 1);
              first_activation_chunk(next, c, d_col);
#endif
            }
        }
      }
      // an fp16 overflow anywhere in the chain has turned into inf / NaN by now (see pack_f16x2_inf)
      if (cur.valid && !(fabsf(dot) <= CUDART_MAX_NORMAL_F)) poisoned = true;

      // ---- coordinate update: xagg[src] += (x_src - x_dst) * phi_x(m)   (models/egnn.py:189-215) ----
      float* dot_slot = s_dot + (it & 1) * kTsParts * kTileM;   // [parity][part][row]: one barrier per tile is enough
      dot_slot[part * kTileM + row] = dot;
      TS_TRACE(7, it);
      ts_epilogue_barrier();
      TS_TRACE(8, it);
      {
        float phi = 0.f;
#pragma unroll
        for (int j = 0; j < kTsParts; ++j) phi += dot_slot[j * kTileM + row];
        phi *= 1.0f / kF16ActScale;   // the last hidden activation was kF16ActScale x the true one (exact)
        if (a.tanh_out) phi = tanhf(phi);
        float norm = 1.0f;
        if (a.normalize) norm = tanhf(cur.radial) / sqrtf(cur.radial + 1e-16f);  // models/egnn.py:240-264, epsilon = 1e-8
        for (int d = part; d < a.D; d += kTsParts) {   // the coordinate columns are dealt over the warps of a row
          float t = 0.f;
          if (cur.valid) t = (a.xcur[(int64_t)cur.src * a.D + d] - a.xcur[(int64_t)cur.dst * a.D + d]) * norm * phi;
          const float sum = seg_scan(t, lane, cur.seg_start);
          if (cur.is_tail && cur.valid) atomicAdd(a.xagg + (int64_t)cur.src * a.D + d, sum);
        }
      }
      cur = next;
    }
    if (poisoned && a.status) atomicOr(a.status, (int)B200MD_STATUS_TC_RANGE);
#if B200MD_TS_SFU_TOKEN
    if (part == 0) sfu_acquire(tok_mine);   // the peer's last release
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (kTsCluster > 1) cluster_sync();   // no CTA leaves while a peer may still multicast into it
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    if (kTsPair) tmem_dealloc_pair(pipe.tmem_base, kTmemCols);
    else tmem_dealloc(pipe.tmem_base, kTmemCols);
  }
}
