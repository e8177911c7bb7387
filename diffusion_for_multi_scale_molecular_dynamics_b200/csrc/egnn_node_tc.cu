// EGNN node kernel on the tensor cores: embedding, node MLP, residual, coordinate update, the node-level halves
// P = h W_a^T + b, Q = h W_b^T of the next layer's first message Linear, and the read-out (logits, score projection).
//
// Reference arithmetic: models/egnn.py:162-187 (node_model), :212-215 (coord += agg), :376-379 (embedding, read-out),
// models/score_networks/egnn_score_network.py:165-224,276-281 (node attributes, uplift, projection),
// models/score_networks/score_network.py:183-185 (MASK logit).  Same pipeline as the edge kernel (tc_common.cuh);
// a tile is 128 consecutive nodes and the program of a call is
//     [ Linear(H+M -> H) SiLU, n x (Linear(H -> H) SiLU), Linear(H -> H) + residual ]      when a layer is updated
//     [ Linear(H -> M) -> P,  Linear(H -> M) -> Q ]                                          when another layer follows.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda at link time)
#include <math_constants.h>

#include "egnn.cuh"
#include "tc_common.cuh"

namespace b200md {

using namespace tc;

namespace {

struct NodeTcArgs {
  Program prog;
  int n_update;          // GEMMs of the node MLP in this call (0 when nothing is updated)
  int do_embed, do_update, do_pq, do_readout;
  const int64_t* atom_types; const float* x_rel; const float* sigma; float sigma_scalar; int sigma_stride;
  float *z, *xcur, *xagg, *h, *P, *Q, *agg;
  const int32_t* deg;
  float* logits_out; float* score_out; int32_t* status;
  int64_t n_nodes; int32_t n_atoms;
  int C, F0, D, H, M, nb;
  int residual, coords_mean, message_mean, mask_logit;
  int64_t bloch, proj, embed_w, embed_b, readout_w, readout_b;
};

struct NodeSmem {
  uint32_t bias, scale, total;
};
__host__ __device__ inline NodeSmem node_smem(int n_gemms) {
  NodeSmem L;
  uint32_t o = pipe_smem().end;
  L.bias = o; o += (uint32_t)(n_gemms > 0 ? n_gemms : 1) * kMaxN * 4;
  L.scale = o; o += 16 * 4;   // 2^-s of each GEMM's weight image (1 in the TF32 mode)
  L.total = o;
  return L;
}

template <int N>
__device__ __forceinline__ void load_row(const float* p, float (&v)[N]) {
#pragma unroll
  for (int j = 0; j < N / 4; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
template <int N>
__device__ __forceinline__ void store_row(float* p, const float (&v)[N]) {
#pragma unroll
  for (int j = 0; j < N / 4; ++j) *reinterpret_cast<float4*>(p + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

template <int PREC>
__global__ void __launch_bounds__(kThreads, 1) egnn_node_tc_kernel(const NodeTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Program& p = a.prog;
  const NodeSmem L = node_smem(p.n_gemms);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  float* s_scale = reinterpret_cast<float*>(smem + L.scale);
  const float* Wp = p.packed;
  constexpr int kChunk = Mode<PREC>::kChunk, kVals = Mode<PREC>::kVals;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (int)((a.n_nodes + kTileM - 1) / kTileM);
  const int H = a.H, M = a.M, D = a.D;

  for (int t = threadIdx.x; t < p.n_gemms * kMaxN; t += kThreads) {
    const int g = t / kMaxN, n = t - g * kMaxN;
    s_bias[t] = (n < p.gN[g] && p.gB[g] >= 0) ? Wp[p.gB[g] + n] : 0.f;
  }
  if (threadIdx.x < p.n_gemms)
    s_scale[threadIdx.x] = (PREC == kF16x3) ? p.wtc[p.gW[threadIdx.x] + weight_unscale_offset(p.gK[threadIdx.x], p.gN[threadIdx.x])] * (1.0f / kF16ActScale) : 1.0f;   // A operands are stored x kF16ActScale
  Pipe<PREC> pipe;
  pipe.setup(smem);
  const TileSchedule sched = make_schedule(n_tiles);

  if (warp == 0) {
    if (lane == 0 && p.n_gemms > 0) pipe.run_producer(p, sched.n_iterations);
  } else if (warp == 1) {
    if (p.n_gemms > 0) pipe.run_mma(p, sched.n_iterations);
  } else {
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_taddr = pipe.tmem_base + ((uint32_t)(q * 32) << 16);
    unsigned local_status = 0;
    float vmax = 0.f;
    for (int it = 0; it < sched.n_iterations; ++it) {
      const int tile = sched.tile(it);
      const int64_t node = (int64_t)tile * kTileM + row;
      const bool valid = node < a.n_nodes;
      const float degf = valid ? fmaxf((float)a.deg[node], 1.0f) : 1.0f;
      float sg = 0.f;
      int atc = 0;
      if (a.do_embed && valid) {
        const int64_t at = a.atom_types[node];
        sg = a.sigma ? a.sigma[(node / a.n_atoms) * a.sigma_stride] : a.sigma_scalar;
        if (hh == 0) {
          if (at < 0 || at >= a.C) local_status |= B200MD_STATUS_A_OUT_OF_RANGE;
          if (!(sg > 0.f)) local_status |= B200MD_STATUS_SIGMA_OUT_OF_RANGE;
        }
        atc = (int)min(max(at, (int64_t)0), (int64_t)(a.C - 1));
      }
      if (hh == 0 && valid) {
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
        if (a.do_update) {  // coord += agg   models/egnn.py:212-215
          for (int d = 0; d < D; ++d) {
            float v = a.xagg[node * D + d];
            if (a.coords_mean) v = v / degf;
            a.xcur[node * D + d] += v;
            a.xagg[node * D + d] = 0.f;
          }
        }
      }

      if (a.do_embed) {
        // h = W_in [sigma, onehot(a)] + b   egnn_score_network.py:165-194, egnn.py:376
        for (int c = 0; c < H / kChunk; ++c) {
          const int col0 = c * kChunk + hh * kVals;
          float v[kVals];
#pragma unroll
          for (int j = 0; j < kVals; ++j) {
            const int o = col0 + j;
            v[j] = valid ? Wp[a.embed_b + o] + (Wp[a.embed_w + o * a.F0] * sg + Wp[a.embed_w + o * a.F0 + 1 + atc]) : 0.f;
          }
          if (valid) store_row<kVals>(a.h + node * H + col0, v);
          if (a.do_pq) vmax = fmaxf(vmax, pipe.push_chunk(row, hh, lane, v));
        }
        if (valid) {
          float zero[kVals];
#pragma unroll
          for (int j = 0; j < kVals; ++j) zero[j] = 0.f;
          for (int c = 0; c < M / kChunk; ++c) store_row<kVals>(a.agg + node * M + c * kChunk + hh * kVals, zero);
        }
      }

      if (a.do_update) {
        // node MLP input [h | aggregated messages]   models/egnn.py:178-183
        for (int c = 0; c < (H + M) / kChunk; ++c) {
          const int col0 = c * kChunk + hh * kVals;
          float v[kVals];
          if (!valid) {
#pragma unroll
            for (int j = 0; j < kVals; ++j) v[j] = 0.f;
          } else if (col0 < H) {
            load_row<kVals>(a.h + node * H + col0, v);
          } else {
            float* ap = a.agg + node * M + (col0 - H);
            load_row<kVals>(ap, v);
            if (a.message_mean) {
#pragma unroll
              for (int j = 0; j < kVals; ++j) v[j] = v[j] / degf;
            }
            float zero[kVals];
#pragma unroll
            for (int j = 0; j < kVals; ++j) zero[j] = 0.f;
            store_row<kVals>(ap, zero);  // reset the accumulator for the next layer
          }
          vmax = fmaxf(vmax, pipe.push_chunk(row, hh, lane, v));
        }
        for (int g = 0; g < a.n_update; ++g) {
          const bool last = (g == a.n_update - 1);
          const uint32_t acc_col = pipe.wait_accumulator();
          const float* bias = s_bias + g * kMaxN;
          const float unscale = s_scale[g];
          for (int c = 0; c < p.gN[g] / kChunk; ++c) {
            const int col0 = c * kChunk + hh * kVals;
            float v[kVals];
            tmem_ld<kVals>(lane_taddr + acc_col + col0, v);
            if (!last) {
#pragma unroll
              for (int j = 0; j < kVals; ++j) v[j] = silu_fast(fmaf(v[j], unscale, bias[col0 + j]));
              vmax = fmaxf(vmax, pipe.push_chunk(row, hh, lane, v));
            } else {
#pragma unroll
              for (int j = 0; j < kVals; ++j) v[j] = fmaf(v[j], unscale, bias[col0 + j]);
              if (valid) {
                if (a.residual) {
                  float old[kVals];
                  load_row<kVals>(a.h + node * H + col0, old);
#pragma unroll
                  for (int j = 0; j < kVals; ++j) v[j] = old[j] + v[j];
                }
                store_row<kVals>(a.h + node * H + col0, v);
              }
              if (a.do_pq) vmax = fmaxf(vmax, pipe.push_chunk(row, hh, lane, v));
            }
          }
          tc_fence_before();
        }
      }

      if (a.do_pq) {
        // second copy of h as the A operand of the Q GEMM (re-read of this thread's own columns)
        for (int c = 0; c < H / kChunk; ++c) {
          const int col0 = c * kChunk + hh * kVals;
          float v[kVals];
          if (valid) {
            load_row<kVals>(a.h + node * H + col0, v);
          } else {
#pragma unroll
            for (int j = 0; j < kVals; ++j) v[j] = 0.f;
          }
          vmax = fmaxf(vmax, pipe.push_chunk(row, hh, lane, v));
        }
        for (int which = 0; which < 2; ++which) {
          const int g = a.n_update + which;
          const uint32_t acc_col = pipe.wait_accumulator();
          const float* bias = s_bias + g * kMaxN;
          const float unscale = s_scale[g];
          float* out = which == 0 ? a.P : a.Q;
          for (int c = 0; c < p.gN[g] / kChunk; ++c) {
            const int col0 = c * kChunk + hh * kVals;
            float v[kVals];
            tmem_ld<kVals>(lane_taddr + acc_col + col0, v);
#pragma unroll
            for (int j = 0; j < kVals; ++j) v[j] = fmaf(v[j], unscale, bias[col0 + j]);
            if (valid) store_row<kVals>(out + node * M + col0, v);
          }
          tc_fence_before();
        }
      }

      if (a.do_readout) {
        __threadfence_block();
        epilogue_barrier();  // both column halves of every h row are now in memory
        if (hh == 0 && valid) {
          // logits = W_c h + b_c (egnn.py:379); MASK logit -> -inf (score_network.py:183-185)
          for (int c = 0; c < a.C; ++c) {
            float s = 0.f;
            for (int k = 0; k < H; ++k) s = fmaf(Wp[a.readout_w + c * H + k], a.h[node * H + k], s);
            float v = s + Wp[a.readout_b + c];
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
        epilogue_barrier();
      }
    }
    if (PREC == kF16x3 && vmax > kF16Safe) local_status |= B200MD_STATUS_TC_RANGE;
    if (local_status && a.status) atomicOr(a.status, (int)local_status);
  }
  pipe.teardown();
}

#include "egnn_node_ts.cuh"

}  // namespace

#ifdef B200MD_TS_TRACE
}  // namespace b200md
// debug only: copy the cycle trace of the traced egnn_node_ts_kernel launch to the host and reset it
extern "C" int b200md_debug_node_trace(unsigned* out_records_u4, int max_records) {
  unsigned n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, b200md::g_node_trace_count, sizeof(n));
  if ((int)n > max_records) n = max_records;
  cudaMemcpyFromSymbol(out_records_u4, b200md::g_node_trace, sizeof(uint4) * n);
  unsigned zero = 0;
  cudaMemcpyToSymbol(b200md::g_node_trace_count, &zero, sizeof(zero));
  return (int)n;
}
namespace b200md {
#endif

bool tc_node_supported(const b200md_egnn_config& cfg, int precision) {
  const int H = cfg.node_hidden, M = cfg.message_hidden;
  const int chunk = precision == kTf32x3 ? 32 : 64;
  if (H % chunk || M % chunk || H > kMaxN || M > kMaxN || H < chunk || M < chunk) return false;
  if (cfg.node_n_hidden + 2 + 3 > kMaxGemms) return false;
  if (precision == kF16x3Ts && cfg.num_atom_types + 1 > kNodeMaxClasses) return false;   // logits kept in registers
  return true;
}

// per layer: [W_a (H -> M)] [W_b (H -> M)] [node_w0 (H+M -> H)] [node hidden ...] [node last (H -> H)]
int64_t tc_node_packed_floats(const b200md_egnn_config& cfg) {
  const int H = cfg.node_hidden, M = cfg.message_hidden;
  return 2 * weight_image_floats(H, M) + weight_image_floats(H + M, H) + (cfg.node_n_hidden + 1) * weight_image_floats(H, H);
}

struct NodeImage {
  int64_t wa, wb, w0, hidden, last;  // offsets (floats) inside a layer's node image
};
static NodeImage node_image(const b200md_egnn_config& cfg) {
  const int64_t H = cfg.node_hidden, M = cfg.message_hidden;
  NodeImage im;
  int64_t o = 0;
  im.wa = o; o += weight_image_floats(H, M);
  im.wb = o; o += weight_image_floats(H, M);
  im.w0 = o; o += weight_image_floats(H + M, H);
  im.hidden = o; o += cfg.node_n_hidden * weight_image_floats(H, H);
  im.last = o;
  return im;
}

int tc_node_pack_layer(const b200md_egnn_config& cfg, const EgnnLayout& lay, int layer, float* packed, cudaStream_t st) {
  const int H = cfg.node_hidden, M = cfg.message_hidden;
  const EgnnLayerOffsets& Lo = lay.layer[layer];
  const NodeImage im = node_image(cfg);
  float* base = packed + lay.tc_node_base + (int64_t)layer * lay.tc_node_stride;
  int rc;
  if ((rc = pack_weight_image(packed + Lo.msg_wa_t, H, M, base + im.wa, st))) return rc;
  if ((rc = pack_weight_image(packed + Lo.msg_wb_t, H, M, base + im.wb, st))) return rc;
  if ((rc = pack_weight_image(packed + Lo.node_w0_t, H + M, H, base + im.w0, st))) return rc;
  for (int i = 0; i < cfg.node_n_hidden; ++i)
    if ((rc = pack_weight_image(packed + Lo.node_hidden_wt + (int64_t)i * H * H, H, H, base + im.hidden + i * weight_image_floats(H, H), st))) return rc;
  if ((rc = pack_weight_image(packed + Lo.node_last_wt, H, H, base + im.last, st))) return rc;
  return B200MD_OK;
}

#if B200MD_NODE_TS_TMA_STORE
// Tensor map of a row-major fp32 array [n_rows, ld] for the node kernel's row stores: box = 32 rows x 32 columns (128 bytes),
// SWIZZLE_128B (the staging tile's XOR pattern), rows beyond n_rows clipped.
typedef CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int node_row_tensor_map(CUtensorMap* map, float* rows, int64_t n_rows, int ld) {
  static TensorMapEncodeTiled encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult status;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status) != cudaSuccess ||
        status != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return (TensorMapEncodeTiled)fn;
  }();
  if (encode == nullptr) return set_cuda_error(cudaErrorNotSupported);   // cuTensorMapEncodeTiled is not available from this driver
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)(n_rows > 0 ? n_rows : 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {32u, 32u};
  const cuuint32_t elem[2] = {1u, 1u};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, rows, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_cuda_error(cudaErrorInvalidValue);   // the encoder refused the array (alignment, extent)
  return B200MD_OK;
}
#endif

int tc_launch_node(const b200md_egnn_config& cfg, const EgnnLayout& lay, const float* packed, const NodeCall& call,
                   int precision, cudaStream_t st) {
  const int H = cfg.node_hidden, M = cfg.message_hidden;
  const NodeImage im = node_image(cfg);
  NodeTcArgs a{};
  Program& p = a.prog;
  p.packed = packed;
  p.wtc = packed + lay.tc_node_base;
  int g = 0;
  if (call.update_layer >= 0) {
    const EgnnLayerOffsets& Lo = lay.layer[call.update_layer];
    const int64_t base = (int64_t)call.update_layer * lay.tc_node_stride;
    p.gK[g] = H + M; p.gN[g] = H; p.gW[g] = base + im.w0; p.gB[g] = Lo.node_b0; ++g;
    for (int i = 0; i < cfg.node_n_hidden; ++i) {
      p.gK[g] = H; p.gN[g] = H; p.gW[g] = base + im.hidden + i * weight_image_floats(H, H);
      p.gB[g] = Lo.node_hidden_b + (int64_t)i * H; ++g;
    }
    p.gK[g] = H; p.gN[g] = H; p.gW[g] = base + im.last; p.gB[g] = Lo.node_last_b; ++g;
  }
  a.n_update = g;
  if (call.pq_layer >= 0) {
    const EgnnLayerOffsets& Lo = lay.layer[call.pq_layer];
    const int64_t base = (int64_t)call.pq_layer * lay.tc_node_stride;
    p.gK[g] = H; p.gN[g] = M; p.gW[g] = base + im.wa; p.gB[g] = Lo.msg_b0; ++g;
    p.gK[g] = H; p.gN[g] = M; p.gW[g] = base + im.wb; p.gB[g] = -1; ++g;
  }
  p.n_gemms = g;
  a.do_embed = call.embed; a.do_update = call.update_layer >= 0; a.do_pq = call.pq_layer >= 0; a.do_readout = call.readout;
  a.atom_types = call.atom_types; a.x_rel = call.x_rel; a.sigma = call.sigma; a.sigma_scalar = call.sigma_scalar; a.sigma_stride = call.sigma_stride;
  a.z = call.ws->z; a.xcur = call.ws->xcur; a.xagg = call.ws->xagg; a.h = call.ws->h; a.P = call.ws->P; a.Q = call.ws->Q;
  a.agg = call.ws->agg; a.deg = call.ws->deg;
  a.logits_out = call.logits_out; a.score_out = call.score_out; a.status = call.status;
  a.n_nodes = call.n_nodes; a.n_atoms = call.n_atoms;
  a.C = lay.C; a.F0 = lay.F0; a.D = lay.D; a.H = H; a.M = M; a.nb = lay.nb;
  a.residual = (cfg.flags & B200MD_EGNN_RESIDUAL) ? 1 : 0;
  a.coords_mean = (cfg.flags & B200MD_EGNN_COORDS_MEAN) ? 1 : 0;
  a.message_mean = (cfg.flags & B200MD_EGNN_MESSAGE_MEAN) ? 1 : 0;
  a.mask_logit = call.mask_logit;
  a.bloch = lay.bloch; a.proj = lay.proj; a.embed_w = lay.embed_w; a.embed_b = lay.embed_b;
  a.readout_w = lay.readout_w; a.readout_b = lay.readout_b;

  if (precision == kF16x3Ts) {
    // the same call in tensor-memory form: the first node Linear split into agg W0b^T (here) + U = h W0a^T + b0 (made
    // by the previous call, next to P and Q)
    NodeTsArgs t{};
    t.wtc = p.wtc; t.packed = packed;
    int n = 0;
    auto f16 = [&](int64_t image, int K, int N) { return image + f16_image_offset(K, N); };
    auto uns = [&](int64_t image, int K, int N) { return image + weight_unscale_offset(K, N); };
    if (call.update_layer >= 0) {
      const EgnnLayerOffsets& Lo = lay.layer[call.update_layer];
      const int64_t base = (int64_t)call.update_layer * lay.tc_node_stride;
      const int64_t w0 = f16(base + im.w0, H + M, H);
      t.g[n] = NodeTsGemm{M, H, kNodeChainAddU, 1, 0, 0, w0 + (int64_t)(H / 64) * 2 * H * 32, uns(base + im.w0, H + M, H), -1, nullptr, -1}; ++n;
      for (int i = 0; i < cfg.node_n_hidden; ++i) {
        const int64_t image = base + im.hidden + i * weight_image_floats(H, H);
        t.g[n] = NodeTsGemm{H, H, kNodeChain, 1, 0, 0, f16(image, H, H), uns(image, H, H), Lo.node_hidden_b + (int64_t)i * H, nullptr, -1}; ++n;
      }
      t.g[n] = NodeTsGemm{H, H, kNodeChainLast, 1, 0, 0, f16(base + im.last, H, H), uns(base + im.last, H, H), Lo.node_last_b, nullptr, -1}; ++n;
    }
    if (call.pq_layer >= 0) {
      const EgnnLayerOffsets& Lo = lay.layer[call.pq_layer];
      const int64_t base = (int64_t)call.pq_layer * lay.tc_node_stride;
      t.g[n] = NodeTsGemm{H, M, kNodeOut, 1, 0, 1, f16(base + im.wa, H, M), uns(base + im.wa, H, M), Lo.msg_b0, call.ws->P, 1}; ++n;
      t.g[n] = NodeTsGemm{H, M, kNodeOut, 0, 1, 1, f16(base + im.wb, H, M), uns(base + im.wb, H, M), -1, call.ws->Q, 2}; ++n;
      t.g[n] = NodeTsGemm{H, H, kNodeOut, 0, 1, 0, f16(base + im.w0, H + M, H), uns(base + im.w0, H + M, H), Lo.node_b0, call.ws->U, 3}; ++n;
    }
    t.n_gemms = n;
    t.do_embed = call.embed; t.do_update = call.update_layer >= 0; t.do_readout = call.readout;
    t.atom_types = call.atom_types; t.x_rel = call.x_rel; t.sigma = call.sigma; t.sigma_scalar = call.sigma_scalar; t.sigma_stride = call.sigma_stride;
    t.z = call.ws->z; t.xcur = call.ws->xcur; t.xagg = call.ws->xagg; t.h = call.ws->h; t.agg = call.ws->agg;
    t.U = call.ws->U; t.deg = call.ws->deg;
    t.logits_out = call.logits_out; t.score_out = call.score_out; t.status = call.status;
    t.n_nodes = call.n_nodes; t.n_atoms = call.n_atoms;
    t.C = lay.C; t.F0 = lay.F0; t.D = lay.D; t.H = H; t.M = M; t.nb = lay.nb;
    t.residual = a.residual; t.coords_mean = a.coords_mean; t.message_mean = a.message_mean; t.mask_logit = call.mask_logit;
    t.bloch = lay.bloch; t.proj = lay.proj; t.embed_w = lay.embed_w; t.embed_b = lay.embed_b;
    t.readout_w = lay.readout_w; t.readout_b = lay.readout_b;
    t.trace = (call.update_layer == 1 && call.pq_layer >= 0) ? 1 : 0;
#if B200MD_NODE_TS_TMA_STORE
    {
      const int rc_h = node_row_tensor_map(&t.tm[0], call.ws->h, call.n_nodes, H);
      if (rc_h != B200MD_OK) return rc_h;
      if (call.pq_layer >= 0) {
        int rc_o;
        if ((rc_o = node_row_tensor_map(&t.tm[1], call.ws->P, call.n_nodes, M)) != B200MD_OK) return rc_o;
        if ((rc_o = node_row_tensor_map(&t.tm[2], call.ws->Q, call.n_nodes, M)) != B200MD_OK) return rc_o;
        if ((rc_o = node_row_tensor_map(&t.tm[3], call.ws->U, call.n_nodes, H)) != B200MD_OK) return rc_o;
      }
    }
#endif
    const size_t ts_bytes = node_ts_smem(n).total + 1024;
    static PerDeviceOnce ts_attr_set;
    if (ts_attr_set.needed()) {
      cudaError_t e0 = cudaFuncSetAttribute(egnn_node_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e0 != cudaSuccess) return set_cuda_error(e0);
      ts_attr_set.done();
    }
    const int64_t tiles = ceil_div64(call.n_nodes, kTileM);
    const unsigned grid = (unsigned)(tiles < 1 ? 1 : (tiles > tc_sm_count() ? tc_sm_count() : tiles));
    { ProfileScope ps__(B200MD_KERNEL_NODE, st); egnn_node_ts_kernel<<<grid, kTsThreads, ts_bytes, st>>>(t); }
    B200MD_CHECK_LAUNCH();
    return B200MD_OK;
  }
  const size_t smem = node_smem(p.n_gemms).total + 1024;
  static PerDeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e1 = cudaFuncSetAttribute(egnn_node_tc_kernel<kTf32x3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaError_t e2 = cudaFuncSetAttribute(egnn_node_tc_kernel<kF16x3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e1 != cudaSuccess) return set_cuda_error(e1);
    if (e2 != cudaSuccess) return set_cuda_error(e2);
    attr_set.done();
  }
  const int64_t n_tiles = ceil_div64(call.n_nodes, kTileM);
  {
    ProfileScope ps__(B200MD_KERNEL_NODE, st);
    cudaError_t e = launch_clustered(precision == kF16x3 ? egnn_node_tc_kernel<kF16x3> : egnn_node_tc_kernel<kTf32x3>,
                                     (unsigned)(n_tiles < 1 ? 1 : (n_tiles > 1000000 ? 1000000 : n_tiles)),
                                     (unsigned)tc_sm_count(), smem, st, a);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

}  // namespace b200md
