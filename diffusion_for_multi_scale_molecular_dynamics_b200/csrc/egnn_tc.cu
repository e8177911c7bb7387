// EGNN edge kernel on the 5th-generation tensor cores (tcgen05 / TMEM), fp32-accurate through a hi/lo split.
//
// Reference arithmetic: models/egnn.py:136-160 (message MLP), :189-215 (coordinate MLP and update), :217-238
// (coord2radial), models/egnn_utils.py:11-70 (segment sum / mean).  Same fusion as the fp32 CUDA-core kernel of
// egnn.cu -- gather, distance featurisation, message MLP, coordinate MLP and the segmented reductions in one
// launch, no per-edge tensor ever reaches HBM -- but every Linear(K -> N) of the chain runs on the tensor cores
// (see tc_common.cuh for the pipeline and the two precision modes).
//
// Two kernels:
//   egnn_edge_ts_kernel (egnn_edge_ts.cuh)   3xFP16, activations live in tensor memory (tcgen05.mma .ts): the production path;
//   egnn_edge_tc_kernel<PREC> (below)         3xTF32 / 3xFP16 with the A operand staged in shared memory: the first
//                                             tensor-core version, kept for widths that are only multiples of 32, for the
//                                             fp32 exponent range, and for A/B runs.
// egnn_edge_tc_kernel: one persistent CTA per SM, tile = 128 consecutive edges of the (src-sorted) edge list.  Warps:
//   warp 0      bulk-copy producer of the weight chunks (L2-resident image -> 4-stage shared-memory ring);
//   warp 1      TMEM allocator and MMA issuer (converged warp, elected lane);
//   warps 2-17  "epilogue" warps (4 per TMEM lane quarter): build the first activation from P[src] + Q[dst] + r w_r, and
//               between two Linear layers read the accumulator (tcgen05.ld), add the bias, apply SiLU, split into hi/lo
//               and write the next layer's A operand chunk by chunk, so that the MMA of layer g+1 on chunk c overlaps
//               the epilogue of layer g on chunk c+1.  They also do the segmented warp reductions (messages -> agg,
//               coordinate updates -> xagg) with one atomic per (segment, column).
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
  uint32_t bias, scale, wr, wlast, dot, total;
};
__host__ __device__ inline EdgeSmem edge_smem(int n_gemms) {
  EdgeSmem L;
  uint32_t o = pipe_smem().end;
  L.bias = o; o += (uint32_t)n_gemms * kMaxN * 4;
  L.scale = o; o += 16 * 4;   // 2^-s of each GEMM's weight image (1 in the TF32 mode)
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
  float* s_scale = reinterpret_cast<float*>(smem + L.scale);
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
  if (threadIdx.x < p.n_gemms)
    s_scale[threadIdx.x] = (PREC == kF16x3) ? p.wtc[p.gW[threadIdx.x] + weight_unscale_offset(p.gK[threadIdx.x], p.gN[threadIdx.x])] * (1.0f / kF16ActScale) : 1.0f;   // A operands are stored x kF16ActScale
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
        const float unscale = s_scale[g];
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = c * kChunk + hh * kVals;
          float v[kVals];
          tmem_ld<kVals>(lane_taddr + acc_col + col0, v);
#pragma unroll
          for (int j = 0; j < kVals; ++j) v[j] = silu_fast(fmaf(v[j], unscale, bias[col0 + j]));
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

#include "egnn_edge_ts.cuh"

// ---- weight image (tc_common.cuh): TF32 parts [w_1 | w_2] per 32-K chunk, then fp16 parts [w_1 | w_2] of 2^s w per
// 64-K chunk, then the trailer [max |w|, 2^-s]
__global__ void tc_absmax_kernel(const float* __restrict__ wt, int total, float* __restrict__ slot) {
  float m = 0.f;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const float w = fabsf(wt[t]);
    if (w <= CUDART_MAX_NORMAL_F) m = fmaxf(m, w);   // NaN / inf weights do not set the scale
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned*>(slot), __float_as_uint(m));   // m >= 0: bit order = value order
}
// power of two that brings the largest weight magnitude into [2^14, 2^15) (1 for an all-zero matrix)
__device__ __forceinline__ int f16_weight_exponent(float absmax) {
  if (!(absmax > 0.f)) return 0;
  int e;
  frexpf(absmax, &e);   // absmax = m 2^e, m in [0.5, 1)
  const int s = 15 - e;
  return s < -100 ? -100 : (s > 100 ? 100 : s);
}
__global__ void tc_pack_kernel(const float* __restrict__ wt /* [K][N] */, int K, int N, float* __restrict__ dst) {
  const int total = K * N;
  __half* dst16 = reinterpret_cast<__half*>(dst + f16_image_offset(K, N));
  const int64_t part = (int64_t)N * 32;  // floats per part (N rows of 128 bytes)
  const int s = f16_weight_exponent(dst[weight_absmax_offset(K, N)]);
  const float scale = ldexpf(1.0f, s);
  if (blockIdx.x == 0 && threadIdx.x == 0) dst[weight_unscale_offset(K, N)] = ldexpf(1.0f, -s);
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
    if ((K & 63) == 0) {  // fp16 image of 2^s w: 2-byte elements, 64 per row, 8 per 16-byte unit
      const float ws = w * scale;   // exact (power of two, no overflow: |ws| < 2^15)
      const __half h1 = __float2half_rn(ws);
      const __half h2 = __float2half_rn(ws - __half2float(h1));
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
  // the trailer slot is zero here: b200md_egnn_pack_weights clears the whole image first
  { ProfileScope ps__(B200MD_KERNEL_OTHER, st); tc_absmax_kernel<<<64, 256, 0, st>>>(wt, K * N, dst + weight_absmax_offset(K, N)); }
  B200MD_CHECK_LAUNCH();
  { ProfileScope ps__(B200MD_KERNEL_OTHER, st); tc_pack_kernel<<<128, 256, 0, st>>>(wt, K, N, dst); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}
}  // namespace tc

#ifdef B200MD_TS_TRACE
}  // namespace b200md
// debug only: copy the cycle trace of the last egnn_edge_ts_kernel launches to the host and reset it
extern "C" int b200md_debug_ts_trace(unsigned* out_records_u4, int max_records) {
  unsigned n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, b200md::g_ts_trace_count, sizeof(n));
  if (n > (unsigned)b200md::kTsTraceMax) n = b200md::kTsTraceMax;
  if ((int)n > max_records) n = max_records;
  cudaMemcpyFromSymbol(out_records_u4, b200md::g_ts_trace, sizeof(uint4) * n);
  unsigned zero = 0;
  cudaMemcpyToSymbol(b200md::g_ts_trace_count, &zero, sizeof(zero));
  return (int)n;
}
namespace b200md {
#endif

int tc_sm_count() {
  static int n_sm[PerDeviceOnce::kMaxDevices] = {};   // per device ordinal
  const int dev = current_device();
  if (dev < 0 || dev >= PerDeviceOnce::kMaxDevices) return 148;
  if (n_sm[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    n_sm[dev] = n > 0 ? n : 148;
  }
  return n_sm[dev];
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
  const int M = cfg.message_hidden, Cx = cfg.coord_hidden;
  return cfg.message_n_hidden * weight_image_floats(M, M) + weight_image_floats(M, Cx) + cfg.coord_n_hidden * weight_image_floats(Cx, Cx);
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
  static PerDeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e0 = cudaFuncSetAttribute(egnn_edge_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e0 != cudaSuccess) return set_cuda_error(e0);
    cudaError_t e1 = cudaFuncSetAttribute(egnn_edge_tc_kernel<kTf32x3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaError_t e2 = cudaFuncSetAttribute(egnn_edge_tc_kernel<kF16x3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e1 != cudaSuccess) return set_cuda_error(e1);
    if (e2 != cudaSuccess) return set_cuda_error(e2);
    attr_set.done();
  }
  if (ts) {
    ProfileScope ps__(B200MD_KERNEL_EDGE, st);
    if (kTsCluster > 1) {
      cudaLaunchConfig_t cfg_l{};
      cfg_l.gridDim = dim3((unsigned)(tc_sm_count() / kTsCluster) * kTsCluster);
      cfg_l.blockDim = dim3(kTsEdgeThreads);
      cfg_l.dynamicSmemBytes = smem;
      cfg_l.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kTsCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg_l.attrs = attr;
      cfg_l.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg_l, egnn_edge_ts_kernel, a);
      if (e != cudaSuccess) return set_cuda_error(e);
    } else {
      egnn_edge_ts_kernel<<<(unsigned)tc_sm_count(), kTsEdgeThreads, smem, st>>>(a);
    }
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
