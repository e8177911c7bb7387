// EGNN score network on sm_100a: packed-weight layout, workspace layout and launch entry points.
// Reference: models/score_networks/egnn_score_network.py:226-303, models/egnn.py:266-288,362-385.
#pragma once
#include "common.cuh"

namespace b200md {

constexpr int kMaxWidth = 256;   // hidden widths supported by the tile kernels
constexpr int kMaxUplift = 32;   // 2 * n_bloch

// Offsets (in floats) into the packed weight image.  All GEMM matrices are stored transposed,
// Wt[K][Nout] (Nout contiguous), so a K-chunk is one contiguous block.
struct EgnnLayerOffsets {
  int64_t msg_wa_t, msg_wb_t, msg_wr, msg_b0;  // first message Linear(2H+1 -> M), split by input block
  int64_t msg_hidden_wt, msg_hidden_b;         // message_n_hidden x ([M][M], [M])
  int64_t att_w, att_b;                        // optional Linear(M -> 1)
  int64_t coord_w0_t, coord_b0;                // Linear(M -> Cx)
  int64_t coord_hidden_wt, coord_hidden_b;     // coord_n_hidden x ([Cx][Cx], [Cx])
  int64_t coord_last_w;                        // Linear(Cx -> 1, no bias)
  int64_t node_w0_t, node_b0;                  // Linear(H+M -> H)
  int64_t node_hidden_wt, node_hidden_b;       // node_n_hidden x ([H][H], [H])
  int64_t node_last_wt, node_last_b;           // Linear(H -> H)
};

struct EgnnLayout {
  int C, F0, D, H, M, Cx, nb;
  int64_t bloch, proj, embed_w, embed_b, readout_w, readout_b;
  EgnnLayerOffsets layer[16];
  int64_t total;
  // tcgen05 weight images (appended after the fp32 section; see tc_common.cuh): one block per layer for the edge
  // MLPs (egnn_tc.cu) and one for the node MLPs (egnn_node_tc.cu); stride 0 = that kernel is not available
  int64_t tc_edge_base, tc_edge_stride;
  int64_t tc_node_base, tc_node_stride;
};

int build_layout(const b200md_egnn_config& cfg, EgnnLayout& lay);

struct EgnnWorkspace {
  float *z, *xcur, *xagg, *h, *P, *Q, *agg;
  float* U;   // h W0a^T + b0 of the next node update (egnn_node_ts.cuh): [B N, H]
  int32_t *deg, *row_ptr, *struct_total, *struct_off, *n_edges, *edge_src, *edge_dst;
  uint32_t* bitmask;
  int64_t bytes;
};

int64_t carve_workspace(const b200md_egnn_config& cfg, int64_t B, int32_t N, int64_t edge_capacity,
                        void* base, EgnnWorkspace& ws);

int egnn_forward_impl(const b200md_egnn_config& cfg, const EgnnLayout& lay, const float* packed,
                      const int64_t* atom_types, const float* x, const float* lattice, const float* sigma,
                      float sigma_scalar, int sigma_stride, int64_t B, int32_t N, int mask_logit, float* logits_out,
                      float* score_out, const EgnnWorkspace& ws, int64_t edge_capacity, int32_t* status,
                      int impl, cudaStream_t st);

// tcgen05 kernels.  `precision` is tc::kTf32x3 (0) or tc::kF16x3 (1), see tc_common.cuh.
int tc_sm_count();
bool tc_edge_supported(const b200md_egnn_config& cfg, int precision);
int64_t tc_edge_packed_floats(const b200md_egnn_config& cfg);
int tc_edge_pack_layer(const b200md_egnn_config& cfg, const EgnnLayout& lay, int layer, float* packed, cudaStream_t st);
int tc_launch_edge_layer(const b200md_egnn_config& cfg, const EgnnLayout& lay, int layer, const float* packed,
                         const EgnnWorkspace& ws, int32_t* status, int precision, cudaStream_t st);

// One launch of the node kernel: optionally embed, update layer `update_layer` (-1: none), produce P/Q for layer
// `pq_layer` (-1: none), read out.
struct NodeCall {
  int embed, update_layer, pq_layer, readout, mask_logit;
  const int64_t* atom_types; const float* x_rel; const float* sigma; float sigma_scalar; int sigma_stride;
  const EgnnWorkspace* ws;
  float* logits_out; float* score_out; int32_t* status;
  int64_t n_nodes; int32_t n_atoms;
};
bool tc_node_supported(const b200md_egnn_config& cfg, int precision);
int64_t tc_node_packed_floats(const b200md_egnn_config& cfg);
int tc_node_pack_layer(const b200md_egnn_config& cfg, const EgnnLayout& lay, int layer, float* packed, cudaStream_t st);
int tc_launch_node(const b200md_egnn_config& cfg, const EgnnLayout& lay, const float* packed, const NodeCall& call,
                   int precision, cudaStream_t st);

}  // namespace b200md
