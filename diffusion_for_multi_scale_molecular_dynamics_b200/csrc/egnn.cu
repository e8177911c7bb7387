// EGNN score network forward: weight packing, node kernels, fp32 CUDA-core edge kernel, launch logic.
// Reference: models/score_networks/egnn_score_network.py:131-303, models/egnn.py:21-385,
//            models/egnn_utils.py:11-70, models/score_networks/score_network.py:183-185.
//
// Per forward (all launches on one stream, no host synchronisation):
//   graph (3 kernels, graph.cu) -> node kernel [embed + uplift + first-layer partial products]
//   -> per layer: edge kernel (gather + distance featurisation + message MLP + coordinate MLP + segmented
//      reduction) and node kernel (node MLP + residual + coordinate update + next layer's partial products,
//      or the read-out: classification logits + score projection on the last layer).
//
// The first message Linear(2H+1 -> M) acts on the concatenation [h_i, h_j, r_ij]; it is evaluated as
// P[i] + Q[j] + r_ij * w_r with P = h W_a^T + b, Q = h W_b^T computed once per NODE (not per edge).
#include <math_constants.h>

#include "egnn.cuh"
#include "graph.cuh"

namespace b200md {

// ---------------------------------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------------------------------
static inline int64_t align64(int64_t v) { return (v + 63) & ~int64_t(63); }  // 256-byte alignment in floats

int build_layout(const b200md_egnn_config& cfg, EgnnLayout& lay) {
  if (cfg.spatial_dimension != 3) return B200MD_ERR_UNSUPPORTED;
  if (cfg.n_layers < 1 || cfg.n_layers > 16 || cfg.num_atom_types < 1 || cfg.n_bloch < 1) return B200MD_ERR_INVALID_ARGUMENT;
  lay.C = cfg.num_atom_types + 1;
  lay.F0 = cfg.num_atom_types + 2;
  lay.nb = cfg.n_bloch;
  lay.D = 2 * cfg.n_bloch;
  lay.H = cfg.node_hidden; lay.M = cfg.message_hidden; lay.Cx = cfg.coord_hidden;
  if (lay.D > kMaxUplift || lay.C > 32) return B200MD_ERR_UNSUPPORTED;
  const int widths[3] = {lay.H, lay.M, lay.Cx};
  for (int w : widths)
    if (w < 4 || w > kMaxWidth || (w % 4) != 0) return B200MD_ERR_UNSUPPORTED;
  if (cfg.node_n_hidden < 0 || cfg.message_n_hidden < 0 || cfg.coord_n_hidden < 0) return B200MD_ERR_INVALID_ARGUMENT;
  int64_t o = 0;
  auto take = [&](int64_t n) { int64_t r = o; o = align64(o + n); return r; };
  const int H = lay.H, M = lay.M, Cx = lay.Cx;
  lay.bloch = take((int64_t)lay.nb * 3);
  lay.proj = take((int64_t)3 * lay.D * lay.D);
  lay.embed_w = take((int64_t)H * lay.F0);
  lay.embed_b = take(H);
  for (int l = 0; l < cfg.n_layers; ++l) {
    EgnnLayerOffsets& L = lay.layer[l];
    L.msg_wa_t = take((int64_t)H * M); L.msg_wb_t = take((int64_t)H * M); L.msg_wr = take(M); L.msg_b0 = take(M);
    L.msg_hidden_wt = take((int64_t)cfg.message_n_hidden * M * M); L.msg_hidden_b = take((int64_t)cfg.message_n_hidden * M);
    L.att_w = take(M); L.att_b = take(1);
    L.coord_w0_t = take((int64_t)M * Cx); L.coord_b0 = take(Cx);
    L.coord_hidden_wt = take((int64_t)cfg.coord_n_hidden * Cx * Cx); L.coord_hidden_b = take((int64_t)cfg.coord_n_hidden * Cx);
    L.coord_last_w = take(Cx);
    L.node_w0_t = take((int64_t)(H + M) * H); L.node_b0 = take(H);
    L.node_hidden_wt = take((int64_t)cfg.node_n_hidden * H * H); L.node_hidden_b = take((int64_t)cfg.node_n_hidden * H);
    L.node_last_wt = take((int64_t)H * H); L.node_last_b = take(H);
  }
  lay.readout_w = take((int64_t)lay.C * H);
  lay.readout_b = take(lay.C);
  lay.tc_edge_base = o;
  lay.tc_edge_stride = 0;
  if (tc_edge_supported(cfg, 0)) {
    lay.tc_edge_stride = align64(tc_edge_packed_floats(cfg));
    o = align64(o + lay.tc_edge_stride * cfg.n_layers);
  }
  lay.tc_node_base = o;
  lay.tc_node_stride = 0;
  if (tc_node_supported(cfg, 0)) {
    lay.tc_node_stride = align64(tc_node_packed_floats(cfg));
    o = align64(o + lay.tc_node_stride * cfg.n_layers);
  }
  lay.total = o;
  return B200MD_OK;
}

int64_t carve_workspace(const b200md_egnn_config& cfg, int64_t B, int32_t N, int64_t cap, void* base, EgnnWorkspace& ws) {
  const int D = 2 * cfg.n_bloch, H = cfg.node_hidden, M = cfg.message_hidden;
  const int64_t BN = B * N;
  const int W = (N + 31) / 32;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o = (o + bytes + 255) & ~int64_t(255); return r; };
  char* p = reinterpret_cast<char*>(base);
  auto at = [&](int64_t off) { return base ? (void*)(p + off) : (void*)nullptr; };
  ws.z = (float*)at(take(BN * D * 4)); ws.xcur = (float*)at(take(BN * D * 4)); ws.xagg = (float*)at(take(BN * D * 4));
  ws.h = (float*)at(take(BN * H * 4)); ws.P = (float*)at(take(BN * M * 4)); ws.Q = (float*)at(take(BN * M * 4));
  ws.agg = (float*)at(take(BN * M * 4));
  ws.U = (float*)at(take(BN * H * 4));
  ws.deg = (int32_t*)at(take(BN * 4)); ws.row_ptr = (int32_t*)at(take((BN + 1) * 4));
  ws.bitmask = (uint32_t*)at(take(BN * W * 4));
  ws.struct_total = (int32_t*)at(take(B * 4)); ws.struct_off = (int32_t*)at(take((B + 1) * 4));
  ws.n_edges = (int32_t*)at(take(256));
  ws.edge_src = (int32_t*)at(take(cap * 4)); ws.edge_dst = (int32_t*)at(take(cap * 4));
  ws.bytes = o;
  return o;
}

// ---------------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------------
// dst[k * n_out + n] = src[n * ld + k0 + k]   (transpose of a column block of a [n_out][ld] matrix)
__global__ void pack_transpose_kernel(const float* __restrict__ src, int n_out, int ld, int k0, int K,
                                      float* __restrict__ dst) {
  const int total = n_out * K;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int k = t / n_out, n = t - k * n_out;
    dst[t] = src[(int64_t)n * ld + k0 + k];
  }
}
__global__ void pack_copy_kernel(const float* __restrict__ src, int n, int stride, float* __restrict__ dst) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = src[(int64_t)t * stride];
}

static int pack_t(const float* src, int n_out, int ld, int k0, int K, float* dst, cudaStream_t st) {
  { ProfileScope ps__(B200MD_KERNEL_OTHER, st); pack_transpose_kernel<<<64, 256, 0, st>>>(src, n_out, ld, k0, K, dst); }
  B200MD_CHECK_LAUNCH();
  return 0;
}
static int pack_c(const float* src, int n, int stride, float* dst, cudaStream_t st) {
  { ProfileScope ps__(B200MD_KERNEL_OTHER, st); pack_copy_kernel<<<8, 256, 0, st>>>(src, n, stride, dst); }
  B200MD_CHECK_LAUNCH();
  return 0;
}

static int param_count(const b200md_egnn_config& cfg) {
  const bool att = cfg.flags & B200MD_EGNN_ATTENTION;
  const int per_layer = 2 * (1 + cfg.message_n_hidden) + 2 * (2 + cfg.node_n_hidden) + 2 * (1 + cfg.coord_n_hidden) + 1 + (att ? 2 : 0);
  return 2 + 2 + cfg.n_layers * per_layer + 2;
}

// ---------------------------------------------------------------------------------------------------
// block-level fp32 tile GEMM:  out[TM][Nout] = act(in[TM][K] @ Wt[K][Nout] + bias)
//   256 threads, warp w owns rows [w*RW, (w+1)*RW), lane owns columns lane + 32 j.
//   Weights stream global(L2) -> smem in K-chunks of 16 rows with cp.async double buffering.
// ---------------------------------------------------------------------------------------------------
constexpr int kThreads = 256;
constexpr int kKC = 16;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int TM>
__device__ __forceinline__ void tile_linear(const float* in_s, int ld_in, float* out_s, int ld_out,
                                            const float* __restrict__ Wt, const float* __restrict__ bias,
                                            int K, int Nout, bool act, float* wst) {
  constexpr int RW = TM / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[RW][8];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
  const int nchunks = (K + kKC - 1) / kKC;
  auto load_chunk = [&](int c, float* dst) {
    const int rows = min(kKC, K - c * kKC);
    const int n4 = rows * Nout / 4;
    const float4* src = reinterpret_cast<const float4*>(Wt + (int64_t)c * kKC * Nout);
    for (int t = threadIdx.x; t < n4; t += kThreads) cp_async16(reinterpret_cast<float4*>(dst) + t, src + t);
  };
  load_chunk(0, wst);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    float* cur = wst + (c & 1) * (kKC * kMaxWidth);
    if (c + 1 < nchunks) {
      load_chunk(c + 1, wst + ((c + 1) & 1) * (kKC * kMaxWidth));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int rows = min(kKC, K - c * kKC);
    for (int kk = 0; kk < rows; kk += 4) {
      float4 a[RW];
#pragma unroll
      for (int r = 0; r < RW; ++r)
        a[r] = *reinterpret_cast<const float4*>(in_s + (warp * RW + r) * ld_in + c * kKC + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = lane + 32 * j;
          w[j] = (col < Nout) ? cur[(kk + q) * Nout + col] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const float av = q == 0 ? a[r].x : (q == 1 ? a[r].y : (q == 2 ? a[r].z : a[r].w));
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(av, w[j], acc[r][j]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = lane + 32 * j;
    if (col < Nout) {
      const float bv = bias ? bias[col] : 0.f;
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        float v = acc[r][j] + bv;
        if (act) v = silu_f(v);
        out_s[(warp * RW + r) * ld_out + col] = v;
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------
// node kernel
// ---------------------------------------------------------------------------------------------------
struct NodeArgs {
  const int64_t* atom_types; const float* x_rel; const float* sigma; float sigma_scalar; int sigma_stride;   // 1: sigma[sample]; 0: sigma[0] for every sample (one device-resident scalar)
  const float* packed;
  float *z, *xcur, *xagg, *h, *P, *Q, *agg;
  const int32_t* deg;
  float* logits_out; float* score_out; int32_t* status;
  int64_t n_nodes; int32_t N;  // atoms per structure
  int C, F0, D, H, M, nb;
  int node_n_hidden;
  int do_embed, do_update, do_pq, do_readout;
  int residual, coords_mean, message_mean, mask_logit;
  // weight offsets
  int64_t bloch, proj, embed_w, embed_b, readout_w, readout_b;
  int64_t node_w0_t, node_b0, node_hidden_wt, node_hidden_b, node_last_wt, node_last_b;  // layer being updated
  int64_t msg_wa_t, msg_wb_t, msg_b0;                                                    // next layer (PQ)
};

constexpr int kTMN = 32;  // nodes per CTA

__global__ void __launch_bounds__(kThreads, 1) egnn_node_kernel(NodeArgs a) {
  extern __shared__ __align__(16) float smem_f[];
  const int H = a.H, M = a.M, D = a.D;
  const int ld_in = H + M;
  float* buf_in = smem_f;                         // [kTMN][H+M]   [h | aggregated messages]
  float* buf_a = buf_in + kTMN * (kMaxWidth * 2); // [kTMN][kMaxWidth]
  float* buf_b = buf_a + kTMN * kMaxWidth;        // [kTMN][kMaxWidth]
  float* wst = buf_b + kTMN * kMaxWidth;          // [2][kKC][kMaxWidth]
  const float* Wp = a.packed;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t node0 = (int64_t)blockIdx.x * kTMN;
  unsigned local_status = 0;

  // ---- stage h (and on the first call: embedding + uplift) ----
  if (a.do_embed) {
    // uplift  z = [cos(2 pi k.x), sin(2 pi k.x)] per Bloch vector   egnn_score_network.py:196-224
    for (int t = threadIdx.x; t < kTMN * a.nb; t += kThreads) {
      const int r = t / a.nb, bidx = t - r * a.nb;
      const int64_t node = node0 + r;
      if (node < a.n_nodes) {
        const float two_pi = 6.283185307179586f;  // float32(2.0 * torch.pi)
        float kr = 0.f;
        for (int s = 0; s < 3; ++s) {
          const float xv = a.x_rel[node * 3 + s];
          if (bidx == 0 && !(xv >= 0.0f && xv < 1.0f)) local_status |= B200MD_STATUS_X_OUT_OF_RANGE;
          kr += Wp[a.bloch + bidx * 3 + s] * (two_pi * xv);
        }
        float sn, cs;
        sincosf(kr, &sn, &cs);
        a.z[node * D + 2 * bidx] = cs; a.z[node * D + 2 * bidx + 1] = sn;
        a.xcur[node * D + 2 * bidx] = cs; a.xcur[node * D + 2 * bidx + 1] = sn;
        a.xagg[node * D + 2 * bidx] = 0.f; a.xagg[node * D + 2 * bidx + 1] = 0.f;
      }
    }
    // h = W_in [sigma, onehot(a)] + b   egnn_score_network.py:165-194, egnn.py:376
    for (int t = threadIdx.x; t < kTMN * H; t += kThreads) {
      const int r = t / H, o = t - r * H;
      const int64_t node = node0 + r;
      float v = 0.f;
      if (node < a.n_nodes) {
        const int64_t at = a.atom_types[node];
        const float sg = a.sigma ? a.sigma[(node / a.N) * a.sigma_stride] : a.sigma_scalar;
        if (o == 0) {
          if (at < 0 || at >= a.C) local_status |= B200MD_STATUS_A_OUT_OF_RANGE;
          if (!(sg > 0.f)) local_status |= B200MD_STATUS_SIGMA_OUT_OF_RANGE;
        }
        const int atc = (int)min(max(at, (int64_t)0), (int64_t)(a.C - 1));
        v = Wp[a.embed_b + o] + (Wp[a.embed_w + o * a.F0] * sg + Wp[a.embed_w + o * a.F0 + 1 + atc]);
      }
      buf_in[r * ld_in + o] = v;
      if (node < a.n_nodes) {
        a.h[node * H + o] = v;
      }
    }
    for (int t = threadIdx.x; t < kTMN * M; t += kThreads) {
      const int r = t / M, o = t - r * M;
      const int64_t node = node0 + r;
      if (node < a.n_nodes) a.agg[node * M + o] = 0.f;
    }
  } else {
    for (int t = threadIdx.x; t < kTMN * H; t += kThreads) {
      const int r = t / H, o = t - r * H;
      const int64_t node = node0 + r;
      buf_in[r * ld_in + o] = (node < a.n_nodes) ? a.h[node * H + o] : 0.f;
    }
  }

  if (a.do_update) {
    // aggregated messages (mean or sum), egnn_utils.py:11-70; then reset the accumulators for the next layer
    for (int t = threadIdx.x; t < kTMN * M; t += kThreads) {
      const int r = t / M, o = t - r * M;
      const int64_t node = node0 + r;
      float v = 0.f;
      if (node < a.n_nodes) {
        v = a.agg[node * M + o];
        if (a.message_mean) v = v / fmaxf((float)a.deg[node], 1.0f);
        a.agg[node * M + o] = 0.f;
      }
      buf_in[r * ld_in + H + o] = v;
    }
    // coordinate update  coord += agg   egnn.py:212-215
    for (int t = threadIdx.x; t < kTMN * D; t += kThreads) {
      const int r = t / D, d = t - r * D;
      const int64_t node = node0 + r;
      if (node < a.n_nodes) {
        float v = a.xagg[node * D + d];
        if (a.coords_mean) v = v / fmaxf((float)a.deg[node], 1.0f);
        a.xcur[node * D + d] += v;
        a.xagg[node * D + d] = 0.f;
      }
    }
    __syncthreads();
    // node MLP  egnn.py:96-105,183-187
    tile_linear<kTMN>(buf_in, ld_in, buf_a, kMaxWidth, Wp + a.node_w0_t, Wp + a.node_b0, H + M, H, true, wst);
    float* cur = buf_a; float* nxt = buf_b;
    for (int i = 0; i < a.node_n_hidden; ++i) {
      tile_linear<kTMN>(cur, kMaxWidth, nxt, kMaxWidth, Wp + a.node_hidden_wt + (int64_t)i * H * H,
                        Wp + a.node_hidden_b + (int64_t)i * H, H, H, true, wst);
      float* t = cur; cur = nxt; nxt = t;
    }
    tile_linear<kTMN>(cur, kMaxWidth, nxt, kMaxWidth, Wp + a.node_last_wt, Wp + a.node_last_b, H, H, false, wst);
    for (int t = threadIdx.x; t < kTMN * H; t += kThreads) {
      const int r = t / H, o = t - r * H;
      const int64_t node = node0 + r;
      float v = nxt[r * kMaxWidth + o];
      if (a.residual) v = buf_in[r * ld_in + o] + v;
      buf_in[r * ld_in + o] = v;
      if (node < a.n_nodes) a.h[node * H + o] = v;
    }
  }
  __syncthreads();

  if (a.do_pq) {
    // P = h W_a^T + b0, Q = h W_b^T : the node-level halves of the next layer's first message Linear
    tile_linear<kTMN>(buf_in, ld_in, buf_a, kMaxWidth, Wp + a.msg_wa_t, Wp + a.msg_b0, H, M, false, wst);
    tile_linear<kTMN>(buf_in, ld_in, buf_b, kMaxWidth, Wp + a.msg_wb_t, nullptr, H, M, false, wst);
    for (int t = threadIdx.x; t < kTMN * M; t += kThreads) {
      const int r = t / M, o = t - r * M;
      const int64_t node = node0 + r;
      if (node < a.n_nodes) { a.P[node * M + o] = buf_a[r * kMaxWidth + o]; a.Q[node * M + o] = buf_b[r * kMaxWidth + o]; }
    }
  }

  if (a.do_readout) {
    // logits = W_c h + b_c (egnn.py:379); MASK logit -> -inf (score_network.py:183-185)
    for (int t = warp; t < kTMN * a.C; t += kThreads / 32) {
      const int r = t / a.C, c = t - r * a.C;
      const int64_t node = node0 + r;
      float s = 0.f;
      for (int k = lane; k < H; k += 32) s += Wp[a.readout_w + c * H + k] * buf_in[r * ld_in + k];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0 && node < a.n_nodes) {
        float v = s + Wp[a.readout_b + c];
        if (a.mask_logit && c == a.C - 1) v = -CUDART_INF_F;
        a.logits_out[node * a.C + c] = v;
      }
    }
    // S^alpha = z . Gamma^alpha . zhat   egnn_score_network.py:276-281
    for (int t = threadIdx.x; t < kTMN * 3; t += kThreads) {
      const int r = t / 3, al = t - r * 3;
      const int64_t node = node0 + r;
      if (node < a.n_nodes) {
        float s = 0.f;
        for (int i = 0; i < D; ++i) {
          const float zi = a.z[node * D + i];
          float inner = 0.f;
          for (int j = 0; j < D; ++j) inner += Wp[a.proj + (al * D + i) * D + j] * a.xcur[node * D + j];
          s += zi * inner;
        }
        a.score_out[node * 3 + al] = s;
      }
    }
  }
  if (local_status && a.status) atomicOr(a.status, (int)local_status);
}

// ---------------------------------------------------------------------------------------------------
// edge kernel (fp32 CUDA cores)
// ---------------------------------------------------------------------------------------------------
struct EdgeArgs {
  const float* packed;
  const float *xcur, *P, *Q;
  float *agg, *xagg;
  const int32_t *edge_src, *edge_dst, *n_edges;
  int D, H, M, Cx;
  int message_n_hidden, coord_n_hidden;
  int attention, normalize, tanh_out;
  EgnnLayerOffsets off;
};

constexpr int kTE = 64;  // edges per tile

__global__ void __launch_bounds__(kThreads, 1) egnn_edge_kernel(EdgeArgs a) {
  extern __shared__ __align__(16) float smem_f[];
  float* buf_a = smem_f;                      // [kTE][kMaxWidth]
  float* buf_b = buf_a + kTE * kMaxWidth;     // [kTE][kMaxWidth]
  float* wst = buf_b + kTE * kMaxWidth;       // [2][kKC][kMaxWidth]
  float* s_diff = wst + 2 * kKC * kMaxWidth;  // [kTE][kMaxUplift]
  float* s_radial = s_diff + kTE * kMaxUplift;  // [kTE]
  float* s_scalar = s_radial + kTE;             // [kTE]
  int* s_row = reinterpret_cast<int*>(s_scalar + kTE);  // [kTE]
  int* s_col = s_row + kTE;                             // [kTE]
  const float* Wp = a.packed;
  const int D = a.D, M = a.M, Cx = a.Cx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = *a.n_edges;
  const int ntiles = (E + kTE - 1) / kTE;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int e0 = tile * kTE;
    const int ne = min(kTE, E - e0);
    if (threadIdx.x < kTE) {
      const int e = e0 + threadIdx.x;
      s_row[threadIdx.x] = (threadIdx.x < ne) ? a.edge_src[e] : -1;
      s_col[threadIdx.x] = (threadIdx.x < ne) ? a.edge_dst[e] : -1;
    }
    __syncthreads();
    // coord2radial  egnn.py:230-238
    for (int t = threadIdx.x; t < kTE * D; t += kThreads) {
      const int e = t / D, d = t - e * D;
      float v = 0.f;
      if (e < ne) v = a.xcur[(int64_t)s_row[e] * D + d] - a.xcur[(int64_t)s_col[e] * D + d];
      s_diff[e * kMaxUplift + d] = v;
    }
    __syncthreads();
    if (threadIdx.x < kTE) {
      float r = 0.f;
      for (int d = 0; d < D; ++d) { const float v = s_diff[threadIdx.x * kMaxUplift + d]; r += v * v; }
      s_radial[threadIdx.x] = r;
    }
    __syncthreads();
    if (a.normalize) {  // egnn.py:240-264, epsilon = 1e-8
      for (int t = threadIdx.x; t < kTE * D; t += kThreads) {
        const int e = t / D, d = t - e * D;
        const float r = s_radial[e];
        s_diff[e * kMaxUplift + d] *= tanhf(r) / sqrtf(r + 1e-16f);
      }
    }
    // first message layer on the concatenation [h_i, h_j, radial], as P[i] + Q[j] + radial * w_r
    for (int e = warp * (kTE / 8); e < (warp + 1) * (kTE / 8); ++e) {
      const bool valid = e < ne;
      const float r = s_radial[e];
      const float* Pr = valid ? a.P + (int64_t)s_row[e] * M : nullptr;
      const float* Qr = valid ? a.Q + (int64_t)s_col[e] * M : nullptr;
      for (int f = lane; f < M; f += 32) {
        float v = 0.f;
        if (valid) v = silu_f(Pr[f] + Qr[f] + r * Wp[a.off.msg_wr + f]);
        buf_a[e * kMaxWidth + f] = v;
      }
    }
    __syncthreads();
    float* cur = buf_a; float* nxt = buf_b;
    for (int i = 0; i < a.message_n_hidden; ++i) {
      tile_linear<kTE>(cur, kMaxWidth, nxt, kMaxWidth, Wp + a.off.msg_hidden_wt + (int64_t)i * M * M,
                       Wp + a.off.msg_hidden_b + (int64_t)i * M, M, M, true, wst);
      float* t = cur; cur = nxt; nxt = t;
    }
    if (a.attention) {  // egnn.py:131-134,157-159
      for (int e = warp * (kTE / 8); e < (warp + 1) * (kTE / 8); ++e) {
        float s = 0.f;
        for (int f = lane; f < M; f += 32) s += cur[e * kMaxWidth + f] * Wp[a.off.att_w + f];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float att = 1.0f / (1.0f + expf(-(s + Wp[a.off.att_b])));
        for (int f = lane; f < M; f += 32) cur[e * kMaxWidth + f] *= att;
      }
      __syncthreads();
    }
    // segmented reduction of the messages over the (sorted) source index: agg[row] += m   egnn.py:178-181
    for (int f = threadIdx.x; f < M; f += kThreads) {
      float run = 0.f;
      int cur_row = s_row[0];
      for (int e = 0; e < ne; ++e) {
        const int r = s_row[e];
        if (r != cur_row) { atomicAdd(a.agg + (int64_t)cur_row * M + f, run); run = 0.f; cur_row = r; }
        run += cur[e * kMaxWidth + f];
      }
      if (ne > 0) atomicAdd(a.agg + (int64_t)cur_row * M + f, run);
    }
    __syncthreads();
    // coordinate MLP  egnn.py:110-129,211
    tile_linear<kTE>(cur, kMaxWidth, nxt, kMaxWidth, Wp + a.off.coord_w0_t, Wp + a.off.coord_b0, M, Cx, true, wst);
    { float* t = cur; cur = nxt; nxt = t; }
    for (int i = 0; i < a.coord_n_hidden; ++i) {
      tile_linear<kTE>(cur, kMaxWidth, nxt, kMaxWidth, Wp + a.off.coord_hidden_wt + (int64_t)i * Cx * Cx,
                       Wp + a.off.coord_hidden_b + (int64_t)i * Cx, Cx, Cx, true, wst);
      float* t = cur; cur = nxt; nxt = t;
    }
    for (int e = warp * (kTE / 8); e < (warp + 1) * (kTE / 8); ++e) {
      float s = 0.f;
      for (int f = lane; f < Cx; f += 32) s += cur[e * kMaxWidth + f] * Wp[a.off.coord_last_w + f];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (a.tanh_out) s = tanhf(s);
      if (lane == 0) s_scalar[e] = s;
    }
    __syncthreads();
    // trans = coord_diff * phi_x(m); xagg[row] += trans
    for (int d = threadIdx.x; d < D; d += kThreads) {
      float run = 0.f;
      int cur_row = s_row[0];
      for (int e = 0; e < ne; ++e) {
        const int r = s_row[e];
        if (r != cur_row) { atomicAdd(a.xagg + (int64_t)cur_row * D + d, run); run = 0.f; cur_row = r; }
        run += s_diff[e * kMaxUplift + d] * s_scalar[e];
      }
      if (ne > 0) atomicAdd(a.xagg + (int64_t)cur_row * D + d, run);
    }
    __syncthreads();
  }
}

static size_t node_smem_bytes() {
  return sizeof(float) * (size_t)(kTMN * kMaxWidth * 2 + 2 * kTMN * kMaxWidth + 2 * kKC * kMaxWidth);
}
static size_t edge_smem_bytes() {
  return sizeof(float) * (size_t)(2 * kTE * kMaxWidth + 2 * kKC * kMaxWidth + kTE * kMaxUplift + 2 * kTE) + sizeof(int) * 2 * kTE;
}

int egnn_forward_impl(const b200md_egnn_config& cfg, const EgnnLayout& lay, const float* packed,
                      const int64_t* atom_types, const float* x, const float* lattice, const float* sigma,
                      float sigma_scalar, int sigma_stride, int64_t B, int32_t N, int mask_logit, float* logits_out,
                      float* score_out, const EgnnWorkspace& ws, int64_t edge_capacity, int32_t* status,
                      int impl, cudaStream_t st) {
  static PerDeviceOnce attrs;
  if (attrs.needed()) {
    cudaError_t e1 = cudaFuncSetAttribute(egnn_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)node_smem_bytes());
    cudaError_t e2 = cudaFuncSetAttribute(egnn_edge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)edge_smem_bytes());
    if (e1 != cudaSuccess) return set_cuda_error(e1);
    if (e2 != cudaSuccess) return set_cuda_error(e2);
    attrs.done();
  }
  // B200MD_IMPL_TCGEN05: 3xTF32 tensor-core kernels (any widths that are multiples of 32); B200MD_IMPL_TCGEN05_F16X3:
  // 3xFP16 (multiples of 64), A operand through shared memory; B200MD_IMPL_TCGEN05_F16X3_TS: 3xFP16 with the edge
  // kernel's A operand in tensor memory; AUTO prefers F16X3_TS, then TF32, then the CUDA-core kernels.
  bool use_tc = false;
  int precision = 0;        // of the node kernel
  int edge_precision = 0;   // of the edge kernel
  if (impl == B200MD_IMPL_TCGEN05 || impl == B200MD_IMPL_TCGEN05_F16X3 || impl == B200MD_IMPL_TCGEN05_F16X3_TS) {
    precision = (impl == B200MD_IMPL_TCGEN05) ? 0 : 1;
    edge_precision = (impl == B200MD_IMPL_TCGEN05) ? 0 : (impl == B200MD_IMPL_TCGEN05_F16X3 ? 1 : 2);
    if (!tc_edge_supported(cfg, edge_precision)) return B200MD_ERR_UNSUPPORTED;
    use_tc = true;
  } else if (impl == B200MD_IMPL_AUTO) {
    if (tc_edge_supported(cfg, 2) && tc_node_supported(cfg, 1)) { use_tc = true; precision = 1; edge_precision = 2; }
    else if (tc_edge_supported(cfg, 0)) { use_tc = true; precision = 0; edge_precision = 0; }
  } else if (impl != B200MD_IMPL_FFMA) {
    return B200MD_ERR_INVALID_ARGUMENT;
  }
  if (use_tc && edge_precision == 2 && tc_node_supported(cfg, 2)) precision = 2;   // node kernel in tensor memory too
  const bool use_tc_node = use_tc && tc_node_supported(cfg, precision);
  const int64_t BN = B * N;
  const bool radial = cfg.flags & B200MD_EGNN_RADIAL_CUTOFF;
  int rc = launch_egnn_graph(x, lattice, cfg.lattice_clip_min, cfg.radial_cutoff, B, N, radial ? 0 : 1,
                             edge_capacity, ws.bitmask, ws.deg, ws.struct_total, ws.struct_off, ws.n_edges,
                             ws.row_ptr, ws.edge_src, ws.edge_dst, status, st);
  if (rc != B200MD_OK) return rc;

  NodeArgs na{};
  na.atom_types = atom_types; na.x_rel = x; na.sigma = sigma; na.sigma_scalar = sigma_scalar; na.sigma_stride = sigma_stride; na.packed = packed;
  na.z = ws.z; na.xcur = ws.xcur; na.xagg = ws.xagg; na.h = ws.h; na.P = ws.P; na.Q = ws.Q; na.agg = ws.agg;
  na.deg = ws.deg; na.logits_out = logits_out; na.score_out = score_out; na.status = status;
  na.n_nodes = BN; na.N = N;
  na.C = lay.C; na.F0 = lay.F0; na.D = lay.D; na.H = lay.H; na.M = lay.M; na.nb = lay.nb;
  na.node_n_hidden = cfg.node_n_hidden;
  na.residual = (cfg.flags & B200MD_EGNN_RESIDUAL) ? 1 : 0;
  na.coords_mean = (cfg.flags & B200MD_EGNN_COORDS_MEAN) ? 1 : 0;
  na.message_mean = (cfg.flags & B200MD_EGNN_MESSAGE_MEAN) ? 1 : 0;
  na.mask_logit = mask_logit;
  na.bloch = lay.bloch; na.proj = lay.proj; na.embed_w = lay.embed_w; na.embed_b = lay.embed_b;
  na.readout_w = lay.readout_w; na.readout_b = lay.readout_b;
  const unsigned node_grid = (unsigned)ceil_div64(BN, kTMN);

  NodeCall nc{};
  nc.mask_logit = mask_logit; nc.atom_types = atom_types; nc.x_rel = x; nc.sigma = sigma; nc.sigma_scalar = sigma_scalar; nc.sigma_stride = sigma_stride;
  nc.ws = &ws; nc.logits_out = logits_out; nc.score_out = score_out; nc.status = status; nc.n_nodes = BN; nc.n_atoms = N;

  // call 0: embedding + uplift + partial products of layer 0
  if (use_tc_node) {
    nc.embed = 1; nc.update_layer = -1; nc.pq_layer = 0; nc.readout = 0;
    rc = tc_launch_node(cfg, lay, packed, nc, precision, st);
    if (rc != B200MD_OK) return rc;
  } else {
    na.do_embed = 1; na.do_update = 0; na.do_pq = 1; na.do_readout = 0;
    na.msg_wa_t = lay.layer[0].msg_wa_t; na.msg_wb_t = lay.layer[0].msg_wb_t; na.msg_b0 = lay.layer[0].msg_b0;
    { ProfileScope ps__(B200MD_KERNEL_NODE, st); egnn_node_kernel<<<node_grid, kThreads, node_smem_bytes(), st>>>(na); }
    B200MD_CHECK_LAUNCH();
  }

  for (int l = 0; l < cfg.n_layers; ++l) {
    if (use_tc) {
      rc = tc_launch_edge_layer(cfg, lay, l, packed, ws, status, edge_precision, st);
      if (rc != B200MD_OK) return rc;
    } else {
      EdgeArgs ea{};
      ea.packed = packed; ea.xcur = ws.xcur; ea.P = ws.P; ea.Q = ws.Q; ea.agg = ws.agg; ea.xagg = ws.xagg;
      ea.edge_src = ws.edge_src; ea.edge_dst = ws.edge_dst; ea.n_edges = ws.n_edges;
      ea.D = lay.D; ea.H = lay.H; ea.M = lay.M; ea.Cx = lay.Cx;
      ea.message_n_hidden = cfg.message_n_hidden; ea.coord_n_hidden = cfg.coord_n_hidden;
      ea.attention = (cfg.flags & B200MD_EGNN_ATTENTION) ? 1 : 0;
      ea.normalize = (cfg.flags & B200MD_EGNN_NORMALIZE) ? 1 : 0;
      ea.tanh_out = (cfg.flags & B200MD_EGNN_TANH) ? 1 : 0;
      ea.off = lay.layer[l];
      int64_t max_tiles = ceil_div64(edge_capacity, kTE);
      unsigned grid = (unsigned)(max_tiles < 148 * 2 ? (max_tiles < 1 ? 1 : max_tiles) : 148 * 2);
      { ProfileScope ps__(B200MD_KERNEL_EDGE, st); egnn_edge_kernel<<<grid, kThreads, edge_smem_bytes(), st>>>(ea); }
      B200MD_CHECK_LAUNCH();
    }
    const bool last = (l == cfg.n_layers - 1);
    if (use_tc_node) {
      nc.embed = 0; nc.update_layer = l; nc.pq_layer = last ? -1 : l + 1; nc.readout = last ? 1 : 0;
      rc = tc_launch_node(cfg, lay, packed, nc, precision, st);
      if (rc != B200MD_OK) return rc;
      continue;
    }
    na.do_embed = 0; na.do_update = 1; na.do_pq = last ? 0 : 1; na.do_readout = last ? 1 : 0;
    const EgnnLayerOffsets& L = lay.layer[l];
    na.node_w0_t = L.node_w0_t; na.node_b0 = L.node_b0; na.node_hidden_wt = L.node_hidden_wt;
    na.node_hidden_b = L.node_hidden_b; na.node_last_wt = L.node_last_wt; na.node_last_b = L.node_last_b;
    if (!last) {
      na.msg_wa_t = lay.layer[l + 1].msg_wa_t; na.msg_wb_t = lay.layer[l + 1].msg_wb_t; na.msg_b0 = lay.layer[l + 1].msg_b0;
    }
    { ProfileScope ps__(B200MD_KERNEL_NODE, st); egnn_node_kernel<<<node_grid, kThreads, node_smem_bytes(), st>>>(na); }
    B200MD_CHECK_LAUNCH();
  }
  return B200MD_OK;
}

}  // namespace b200md

using namespace b200md;

extern "C" {

int32_t b200md_egnn_param_count(const b200md_egnn_config* cfg) { return cfg ? param_count(*cfg) : -1; }

int64_t b200md_egnn_packed_floats(const b200md_egnn_config* cfg) {
  if (!cfg) return B200MD_ERR_INVALID_ARGUMENT;
  EgnnLayout lay;
  int rc = build_layout(*cfg, lay);
  return rc == B200MD_OK ? lay.total : rc;
}

int b200md_egnn_pack_weights(const b200md_egnn_config* cfg, const float* const* p, int32_t n_params,
                             float* packed, b200md_stream_t stream) {
  if (!cfg || !p || !packed) return B200MD_ERR_INVALID_ARGUMENT;
  EgnnLayout lay;
  int rc = build_layout(*cfg, lay);
  if (rc != B200MD_OK) return rc;
  if (n_params != param_count(*cfg)) return B200MD_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < n_params; ++i)
    if (!p[i]) return B200MD_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(packed, 0, sizeof(float) * lay.total, st);
  const int H = lay.H, M = lay.M, Cx = lay.Cx;
  const bool att = cfg->flags & B200MD_EGNN_ATTENTION;
  int i = 0;
  if (pack_c(p[i++], lay.nb * 3, 1, packed + lay.bloch, st)) return B200MD_ERR_CUDA;
  if (pack_c(p[i++], 3 * lay.D * lay.D, 1, packed + lay.proj, st)) return B200MD_ERR_CUDA;
  if (pack_c(p[i++], H * lay.F0, 1, packed + lay.embed_w, st)) return B200MD_ERR_CUDA;
  if (pack_c(p[i++], H, 1, packed + lay.embed_b, st)) return B200MD_ERR_CUDA;
  for (int l = 0; l < cfg->n_layers; ++l) {
    const EgnnLayerOffsets& L = lay.layer[l];
    const float* w0 = p[i++]; const float* b0 = p[i++];  // message_mlp.0  [M][2H+1]
    if (pack_t(w0, M, 2 * H + 1, 0, H, packed + L.msg_wa_t, st)) return B200MD_ERR_CUDA;
    if (pack_t(w0, M, 2 * H + 1, H, H, packed + L.msg_wb_t, st)) return B200MD_ERR_CUDA;
    if (pack_c(w0 + 2 * H, M, 2 * H + 1, packed + L.msg_wr, st)) return B200MD_ERR_CUDA;
    if (pack_c(b0, M, 1, packed + L.msg_b0, st)) return B200MD_ERR_CUDA;
    for (int k = 0; k < cfg->message_n_hidden; ++k) {
      if (pack_t(p[i++], M, M, 0, M, packed + L.msg_hidden_wt + (int64_t)k * M * M, st)) return B200MD_ERR_CUDA;
      if (pack_c(p[i++], M, 1, packed + L.msg_hidden_b + (int64_t)k * M, st)) return B200MD_ERR_CUDA;
    }
    // node_mlp
    if (pack_t(p[i++], H, H + M, 0, H + M, packed + L.node_w0_t, st)) return B200MD_ERR_CUDA;
    if (pack_c(p[i++], H, 1, packed + L.node_b0, st)) return B200MD_ERR_CUDA;
    for (int k = 0; k < cfg->node_n_hidden; ++k) {
      if (pack_t(p[i++], H, H, 0, H, packed + L.node_hidden_wt + (int64_t)k * H * H, st)) return B200MD_ERR_CUDA;
      if (pack_c(p[i++], H, 1, packed + L.node_hidden_b + (int64_t)k * H, st)) return B200MD_ERR_CUDA;
    }
    if (pack_t(p[i++], H, H, 0, H, packed + L.node_last_wt, st)) return B200MD_ERR_CUDA;
    if (pack_c(p[i++], H, 1, packed + L.node_last_b, st)) return B200MD_ERR_CUDA;
    // coord_mlp
    if (pack_t(p[i++], Cx, M, 0, M, packed + L.coord_w0_t, st)) return B200MD_ERR_CUDA;
    if (pack_c(p[i++], Cx, 1, packed + L.coord_b0, st)) return B200MD_ERR_CUDA;
    for (int k = 0; k < cfg->coord_n_hidden; ++k) {
      if (pack_t(p[i++], Cx, Cx, 0, Cx, packed + L.coord_hidden_wt + (int64_t)k * Cx * Cx, st)) return B200MD_ERR_CUDA;
      if (pack_c(p[i++], Cx, 1, packed + L.coord_hidden_b + (int64_t)k * Cx, st)) return B200MD_ERR_CUDA;
    }
    if (pack_c(p[i++], Cx, 1, packed + L.coord_last_w, st)) return B200MD_ERR_CUDA;
    if (att) {
      if (pack_c(p[i++], M, 1, packed + L.att_w, st)) return B200MD_ERR_CUDA;
      if (pack_c(p[i++], 1, 1, packed + L.att_b, st)) return B200MD_ERR_CUDA;
    }
  }
  if (pack_c(p[i++], lay.C * H, 1, packed + lay.readout_w, st)) return B200MD_ERR_CUDA;
  if (pack_c(p[i++], lay.C, 1, packed + lay.readout_b, st)) return B200MD_ERR_CUDA;
  for (int l = 0; l < cfg->n_layers; ++l) {
    if (lay.tc_edge_stride > 0 && (rc = tc_edge_pack_layer(*cfg, lay, l, packed, st)) != B200MD_OK) return rc;
    if (lay.tc_node_stride > 0 && (rc = tc_node_pack_layer(*cfg, lay, l, packed, st)) != B200MD_OK) return rc;
  }
  return B200MD_OK;
}

int64_t b200md_egnn_default_edge_capacity(const b200md_egnn_config* cfg, int64_t B, int32_t N,
                                          const float* cell_lengths_host) {
  if (!cfg || B < 0 || N <= 0) return B200MD_ERR_INVALID_ARGUMENT;
  int64_t per_node = N - 1;
  if ((cfg->flags & B200MD_EGNN_RADIAL_CUTOFF) && N > 128) {
    // expected degree for a uniform gas in the clipped box, times 2.5, plus slack
    double vol = 1.0;
    for (int d = 0; d < 3; ++d) {
      double len = cell_lengths_host ? (double)cell_lengths_host[d] : (double)cfg->lattice_clip_min;
      if (len < cfg->lattice_clip_min) len = cfg->lattice_clip_min;
      vol *= len;
    }
    const double rc = cfg->radial_cutoff;
    const double expected = (double)N * (4.0 / 3.0) * 3.141592653589793 * rc * rc * rc / vol;
    int64_t est = (int64_t)(2.5 * expected) + 32;
    if (est < per_node) per_node = est;
  }
  int64_t cap = B * N * per_node;
  if (cap < 64) cap = 64;
  if (cap > 0x7fffffffLL) cap = 0x7fffffffLL;
  return cap;
}

int64_t b200md_egnn_workspace_bytes(const b200md_egnn_config* cfg, int64_t B, int32_t N, int64_t cap) {
  if (!cfg || B < 0 || N <= 0 || cap < 0) return B200MD_ERR_INVALID_ARGUMENT;
  EgnnWorkspace ws;
  return carve_workspace(*cfg, B, N, cap, nullptr, ws);
}

int b200md_egnn_resolve_impl(const b200md_egnn_config* cfg, int32_t impl) {
  if (!cfg) return B200MD_ERR_INVALID_ARGUMENT;
  EgnnLayout lay;
  int rc = build_layout(*cfg, lay);
  if (rc != B200MD_OK) return rc;
  if (impl == B200MD_IMPL_FFMA) return B200MD_IMPL_FFMA;
  if (impl == B200MD_IMPL_TCGEN05 || impl == B200MD_IMPL_TCGEN05_F16X3 || impl == B200MD_IMPL_TCGEN05_F16X3_TS)
    return tc_edge_supported(*cfg, impl == B200MD_IMPL_TCGEN05 ? 0 : (impl == B200MD_IMPL_TCGEN05_F16X3 ? 1 : 2))
               ? impl : B200MD_ERR_UNSUPPORTED;
  if (impl == B200MD_IMPL_AUTO) {
    if (tc_edge_supported(*cfg, 2) && tc_node_supported(*cfg, 1)) return B200MD_IMPL_TCGEN05_F16X3_TS;
    return tc_edge_supported(*cfg, 0) ? B200MD_IMPL_TCGEN05 : B200MD_IMPL_FFMA;
  }
  return B200MD_ERR_INVALID_ARGUMENT;
}

int b200md_egnn_forward(const b200md_egnn_config* cfg, const float* packed, const int64_t* atom_types,
                        const float* x, const float* lattice, const float* sigma, int64_t B, int32_t N,
                        int32_t mask_logit, float* logits_out, float* score_out, void* workspace,
                        int64_t workspace_bytes, int64_t edge_capacity, int64_t* n_edges_out,
                        int32_t* status, int32_t impl, b200md_stream_t stream) {
  if (!cfg || !packed || !atom_types || !x || !lattice || !sigma || !logits_out || !score_out || !workspace ||
      !status || B < 0 || N <= 0 || edge_capacity <= 0)
    return B200MD_ERR_INVALID_ARGUMENT;
  if (B == 0) return B200MD_OK;
  EgnnLayout lay;
  int rc = build_layout(*cfg, lay);
  if (rc != B200MD_OK) return rc;
  EgnnWorkspace ws;
  if (carve_workspace(*cfg, B, N, edge_capacity, workspace, ws) > workspace_bytes) return B200MD_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = as_stream(stream);
  rc = egnn_forward_impl(*cfg, lay, packed, atom_types, x, lattice, sigma, 0.f, 1, B, N, mask_logit, logits_out,
                         score_out, ws, edge_capacity, status, impl, st);
  if (rc != B200MD_OK) return rc;
  if (n_edges_out) {
    // widen the device-side int32 edge count into the caller's int64 slot
    cudaMemsetAsync(n_edges_out, 0, 8, st);
    cudaMemcpyAsync(n_edges_out, ws.n_edges, 4, cudaMemcpyDeviceToDevice, st);
  }
  return B200MD_OK;
}

}  // extern "C"
