// Periodic radius graph: replaces get_periodic_adjacency_information (utils/neighbors.py:36-224),
// get_adj_matrix (models/graph_utils.py:10-50) and get_edges_with_radial_cutoff
// (models/egnn_utils.py:107-144).
//
// One CTA per structure.  The structure's Cartesian positions are staged once in shared memory
// (12 B per atom); each warp owns source atoms and sweeps destination atoms 32 at a time, testing the 27
// periodic images with the reference's dense fp32 rounding chain, then compacts the hits with
// __ballot_sync / __popc so that edges come out already sorted by (src, dst) -- the order torch.unique
// would produce -- without any sort.  At the sizes of the sampling path (L'/rc between 2.2 and 2.9) a
// cell list has at most two bins per axis and prunes nothing, so the sweep is the degenerate (one-bin)
// cell list; see DESIGN.md.
#include "common.cuh"
#include "graph.cuh"

namespace b200md {

// ---- API kernels (int64 outputs, arbitrary cell) ------------------------------------------------------

__global__ void radius_graph_count_kernel(const float* __restrict__ positions,
                                          const float* __restrict__ cell, int relative, float rc,
                                          float rc2, int32_t N, int32_t mode,
                                          int32_t* __restrict__ counts,
                                          int64_t* __restrict__ edges_per_structure,
                                          int32_t* __restrict__ status, uint32_t* __restrict__ hit_mask) {
  extern __shared__ float smem[];
  float* px = smem; float* py = px + N; float* pz = py + N;
  float* lat = pz + N;  // [27][3]
  __shared__ float s_cell[9];
  __shared__ unsigned long long s_total;
  __shared__ ImageCull s_cull;
  const int64_t b = blockIdx.x;
  stage_structure(positions + b * N * 3, cell + b * 9, relative, N, px, py, pz, lat, s_cell);
  if (threadIdx.x == 0) {
    s_total = 0ull;
    if (!(shortest_cell_crossing_distance(s_cell) > rc)) atomicOr(status, (int)B200MD_STATUS_CUTOFF_TOO_LARGE);
  }
  __syncthreads();
  image_cull_setup(&s_cull, px, py, pz, N, s_cell);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  unsigned long long local = 0;
  if (mode == B200MD_GRAPH_UNIQUE_PAIRS) {
    for (int i = warp; i < N; i += nwarps) {
      const float xi = px[i], yi = py[i], zi = pz[i];
      const unsigned candidates = image_candidates(&s_cull, xi, yi, zi, rc);
      int deg = 0;
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j < N) hit = first_image_hit_among(xi, yi, zi, px[j], py[j], pz[j], lat, rc2, candidates) >= 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        deg += __popc(ballot);
        // the hit bit row of (b, i): the fill expands it without sweeping again (b200md_radius_graph_fill_from_mask)
        if (hit_mask != nullptr && lane == 0) hit_mask[(b * N + i) * (int64_t)((N + 31) >> 5) + (j0 >> 5)] = ballot;
      }
      if (lane == 0) { counts[b * N + i] = deg; local += deg; }
    }
  } else {
    for (int r = warp; r < 27 * N; r += nwarps) {
      const int img = r / N, i = r - img * N;
      const float xi = px[i], yi = py[i], zi = pz[i];
      const float lx = lat[img * 3], ly = lat[img * 3 + 1], lz = lat[img * 3 + 2];
      int deg = 0;
      if (!image_can_hit(&s_cull, xi, yi, zi, img, rc)) {   // no atom of this image can be inside the cutoff
        if (lane == 0) counts[(b * 27 + img) * N + i] = 0;
        continue;
      }
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        float d2;
        if (j < N) hit = image_hit(xi, yi, zi, px[j], py[j], pz[j], lx, ly, lz, rc2, d2);
        deg += __popc(__ballot_sync(0xffffffffu, hit));
      }
      if (lane == 0) { counts[(b * 27 + img) * N + i] = deg; local += deg; }
    }
  }
  if (lane == 0 && local) atomicAdd(&s_total, local);
  __syncthreads();
  if (threadIdx.x == 0) edges_per_structure[b] = (int64_t)s_total;
}

__global__ void radius_graph_fill_kernel(const float* __restrict__ positions, const float* __restrict__ cell,
                                         int relative, float rc, float rc2, int32_t N, int32_t mode,
                                         const int64_t* __restrict__ offsets, int64_t node_offset,
                                         int64_t* __restrict__ src_out, int64_t* __restrict__ dst_out,
                                         float* __restrict__ shifts_out, int64_t* __restrict__ edge_batch) {
  extern __shared__ float smem[];
  float* px = smem; float* py = px + N; float* pz = py + N;
  float* lat = pz + N;  // [27][3]
  float* scratch = lat + 81;  // per-warp [N] d2 + [N] j  (PER_IMAGE mode only)
  __shared__ float s_cell[9];
  __shared__ ImageCull s_cull;
  const int64_t b = blockIdx.x;
  stage_structure(positions + b * N * 3, cell + b * 9, relative, N, px, py, pz, lat, s_cell);
  __syncthreads();
  image_cull_setup(&s_cull, px, py, pz, N, s_cell);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t base_node = b * node_offset;
  if (mode == B200MD_GRAPH_UNIQUE_PAIRS) {
    for (int i = warp; i < N; i += nwarps) {
      const float xi = px[i], yi = py[i], zi = pz[i];
      int64_t pos = offsets[b * N + i];
      const unsigned candidates = image_candidates(&s_cull, xi, yi, zi, rc);
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        int img = -1;
        if (j < N) img = first_image_hit_among(xi, yi, zi, px[j], py[j], pz[j], lat, rc2, candidates);
        const unsigned ballot = __ballot_sync(0xffffffffu, img >= 0);
        if (img >= 0) {
          const int64_t e = pos + __popc(ballot & lt_mask);
          src_out[e] = base_node + i;
          dst_out[e] = base_node + j;
          if (shifts_out) { shifts_out[e * 3] = lat[img * 3]; shifts_out[e * 3 + 1] = lat[img * 3 + 1]; shifts_out[e * 3 + 2] = lat[img * 3 + 2]; }
          if (edge_batch) edge_batch[e] = b;
        }
        pos += __popc(ballot);
      }
    }
  } else {
    float* wd2 = scratch + (size_t)warp * 2 * N;
    int* wj = reinterpret_cast<int*>(wd2 + N);
    for (int r = warp; r < 27 * N; r += nwarps) {
      const int img = r / N, i = r - img * N;
      const float xi = px[i], yi = py[i], zi = pz[i];
      const float lx = lat[img * 3], ly = lat[img * 3 + 1], lz = lat[img * 3 + 2];
      int cnt = 0;
      if (!image_can_hit(&s_cull, xi, yi, zi, img, rc)) continue;   // an empty row (its count is 0)
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        float d2 = 0.f;
        if (j < N) hit = image_hit(xi, yi, zi, px[j], py[j], pz[j], lx, ly, lz, rc2, d2);
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (hit) { const int k = cnt + __popc(ballot & lt_mask); wd2[k] = d2; wj[k] = j; }
        cnt += __popc(ballot);
      }
      __syncwarp();
      const int64_t pos = offsets[(b * 27 + img) * N + i];
      // rank by (d2, j): ascending distance like Kmin_argKmin (neighbors.py:184-186)
      for (int k = lane; k < cnt; k += 32) {
        const float dk = wd2[k]; const int jk = wj[k];
        int rank = 0;
        for (int m = 0; m < cnt; ++m) {
          const float dm = wd2[m];
          rank += (dm < dk) || (dm == dk && wj[m] < jk);
        }
        const int64_t e = pos + rank;
        src_out[e] = base_node + i;
        dst_out[e] = base_node + jk;
        if (shifts_out) { shifts_out[e * 3] = lx; shifts_out[e * 3 + 1] = ly; shifts_out[e * 3 + 2] = lz; }
        if (edge_batch) edge_batch[e] = b;
      }
      __syncwarp();
    }
  }
}

// Unique-pair edges from the hit bit rows the count pass left behind: one warp per source atom, no positions, no second
// N^2 sweep.  Ascending destination inside a row, rows in (structure, source) order: the torch.unique(dim=1) order.
__global__ void radius_graph_fill_from_mask_kernel(const uint32_t* __restrict__ hit_mask, int64_t n_rows, int32_t N,
                                                   const int64_t* __restrict__ offsets, int64_t node_offset,
                                                   int64_t* __restrict__ src_out, int64_t* __restrict__ dst_out,
                                                   int64_t* __restrict__ edge_batch) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int words = (N + 31) >> 5;
  const int64_t b = row / N;
  const int i = (int)(row - b * N);
  const int64_t base_node = b * node_offset;
  const unsigned lt_mask = (1u << lane) - 1u;
  int64_t pos = offsets[row];
  for (int w = 0; w < words; ++w) {
    const unsigned bits = hit_mask[row * words + w];   // warp-uniform
    if (bits >> lane & 1u) {
      const int64_t e = pos + __popc(bits & lt_mask);
      src_out[e] = base_node + i;
      dst_out[e] = base_node + (w * 32 + lane);
      if (edge_batch) edge_batch[e] = b;
    }
    pos += __popc(bits);
  }
}

// ---- internal kernels for the EGNN forward (int32 CSR, cell = diag(clip(L))) ---------------------------

__global__ void egnn_graph_count_kernel(const float* __restrict__ x_rel, const float* __restrict__ lattice,
                                        float clip_min, float rc, float rc2, int32_t N, int32_t W,
                                        int fully_connected, uint32_t* __restrict__ bitmask,
                                        int32_t* __restrict__ deg, int32_t* __restrict__ struct_total,
                                        int32_t* __restrict__ status) {
  extern __shared__ float smem[];
  float* px = smem; float* py = px + N; float* pz = py + N;
  float* lat = pz + N;
  __shared__ float s_cell[9];
  __shared__ int s_total;
  __shared__ ImageCull s_cull;
  const int64_t b = blockIdx.x;
  if (threadIdx.x == 0) s_total = 0;
  if (!fully_connected) {
    if (threadIdx.x < 9) {
      // lattice_parameters.clip(min=2.2 rc); [:, d:] = 0; diag_embed   egnn_score_network.py:236-240
      const int r = threadIdx.x / 3, c = threadIdx.x % 3;
      s_cell[threadIdx.x] = (r == c) ? fmaxf(lattice[b * 6 + r], clip_min) : 0.0f;
    }
    __syncthreads();
    stage_structure_with_cell(x_rel + b * N * 3, 1, N, px, py, pz, lat, s_cell);
    if (threadIdx.x == 0 && !(shortest_cell_crossing_distance(s_cell) > rc))
      atomicOr(status, (int)B200MD_STATUS_CUTOFF_TOO_LARGE);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  int local = 0;
  // minimum-image shortcut (graph.cuh): valid when the cutoff is below half of every box length, which the clipped
  // EGNN box guarantees unless the cutoff assert above has fired
  const float Lx = s_cell[0], Ly = s_cell[4], Lz = s_cell[8];
  const bool min_image = !fully_connected && (2.0f * rc < fminf(Lx, fminf(Ly, Lz)));
  if (!fully_connected && !min_image) image_cull_setup(&s_cull, px, py, pz, N, s_cell);   // CTA-uniform branch
  for (int i = warp; i < N; i += nwarps) {
    float xi = 0.f, yi = 0.f, zi = 0.f;
    if (!fully_connected) { xi = px[i]; yi = py[i]; zi = pz[i]; }
    const unsigned candidates = (!fully_connected && !min_image) ? image_candidates(&s_cull, xi, yi, zi, rc) : 0u;
    int d = 0;
    for (int w = 0; w < W; ++w) {
      const int j = w * 32 + lane;
      bool hit = false;
      if (j < N)
        hit = fully_connected ? (j != i)
                              : (min_image ? min_image_hit(xi, yi, zi, px[j], py[j], pz[j], Lx, Ly, Lz, rc2)
                                           : first_image_hit_among(xi, yi, zi, px[j], py[j], pz[j], lat, rc2, candidates) >= 0);
      const unsigned ballot = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) bitmask[(b * N + i) * W + w] = ballot;
      d += __popc(ballot);
    }
    if (lane == 0) { deg[b * N + i] = d; local += d; }
  }
  if (lane == 0 && local) atomicAdd(&s_total, local);
  __syncthreads();
  if (threadIdx.x == 0) struct_total[b] = s_total;
}

// exclusive scan over B structure totals (single CTA), capacity check
__global__ void egnn_graph_scan_kernel(const int32_t* __restrict__ struct_total, int64_t B,
                                       int64_t capacity, int32_t* __restrict__ struct_off,
                                       int32_t* __restrict__ n_edges, int32_t* __restrict__ status) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t base = 0; base < B; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    long long v = (i < B) ? struct_total[i] : 0;
    long long incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      long long w = lane < nwarps ? s_warp[lane] : 0;
      long long wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      s_warp[lane] = wi - w;  // exclusive prefix of the warp sums
    }
    __syncthreads();
    const long long excl = s_carry + s_warp[warp] + incl - v;
    if (i < B) struct_off[i] = (int32_t)(excl > 0x7fffffffLL ? 0x7fffffffLL : excl);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    long long total = s_carry;
    if (total > capacity || total > 0x7fffffffLL) {
      atomicOr(status, (int)B200MD_STATUS_EDGE_OVERFLOW);
      total = 0;  // downstream kernels process no edges; the host raises after reading status
    }
    struct_off[B] = (int32_t)total;
    *n_edges = (int32_t)total;
  }
}

__global__ void egnn_graph_fill_kernel(const uint32_t* __restrict__ bitmask, const int32_t* __restrict__ deg,
                                       const int32_t* __restrict__ struct_off, const int32_t* __restrict__ n_edges,
                                       int32_t N, int32_t W, int64_t B, int32_t* __restrict__ row_ptr,
                                       int32_t* __restrict__ edge_src, int32_t* __restrict__ edge_dst) {
  extern __shared__ int s_rowoff[];  // [N+1] local exclusive scan of the degrees
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int64_t b = blockIdx.x;
  const bool overflow = (*n_edges == 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < N; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int v = (i < N) ? deg[b * N + i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = lane < nwarps ? s_warp[lane] : 0;
      int wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - v;
    if (i < N) s_rowoff[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
    __syncthreads();
  }
  const int base_edge = overflow ? 0 : struct_off[b];
  for (int i = threadIdx.x; i < N; i += blockDim.x) row_ptr[b * N + i] = overflow ? 0 : base_edge + s_rowoff[i];
  if (b == B - 1 && threadIdx.x == 0) row_ptr[B * N] = *n_edges;
  if (overflow) return;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int i = warp; i < N; i += nwarps) {
    int pos = base_edge + s_rowoff[i];
    const int node = (int)(b * N) + i;
    for (int w = 0; w < W; ++w) {
      const unsigned word = bitmask[(b * N + i) * W + w];
      if (word & (1u << lane)) {
        const int e = pos + __popc(word & lt_mask);
        edge_src[e] = node;
        edge_dst[e] = (int)(b * N) + w * 32 + lane;
      }
      pos += __popc(word);
    }
  }
}

int launch_egnn_graph(const float* x_rel, const float* lattice, float clip_min, float rc, int64_t B,
                      int32_t N, int fully_connected, int64_t capacity, uint32_t* bitmask, int32_t* deg,
                      int32_t* struct_total, int32_t* struct_off, int32_t* n_edges, int32_t* row_ptr,
                      int32_t* edge_src, int32_t* edge_dst, int32_t* status, cudaStream_t st) {
  if (N > 4096) return B200MD_ERR_UNSUPPORTED;
  const int W = (N + 31) / 32;
  const float rc2 = rc * rc;  // radial_cutoff**2 on a float32 tensor (neighbors.py:103,172)
  int threads = N >= 256 ? 256 : (N >= 64 ? 128 : 64);
  size_t smem = (size_t)(3 * N + 81) * sizeof(float);
  if (smem > 48 * 1024) {   // N > 4069: above the default dynamic shared-memory limit
    cudaError_t e = cudaFuncSetAttribute(egnn_graph_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, st); egnn_graph_count_kernel<<<(unsigned)B, threads, smem, st>>>(x_rel, lattice, clip_min, rc, rc2, N, W,
                                                              fully_connected, bitmask, deg, struct_total, status); }
  B200MD_CHECK_LAUNCH();
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, st); egnn_graph_scan_kernel<<<1, 1024, 0, st>>>(struct_total, B, capacity, struct_off, n_edges, status); }
  B200MD_CHECK_LAUNCH();
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, st); egnn_graph_fill_kernel<<<(unsigned)B, threads, (size_t)(N + 1) * sizeof(int), st>>>(
      bitmask, deg, struct_off, n_edges, N, W, B, row_ptr, edge_src, edge_dst); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

}  // namespace b200md

using namespace b200md;

extern "C" {

static int radius_graph_count_impl(const float* positions, const float* cell, int32_t positions_are_relative,
                                   float radial_cutoff, int64_t n_samples, int32_t n_atoms, int32_t mode,
                                   int32_t* counts, int64_t* edges_per_structure, int32_t* status,
                                   uint32_t* hit_mask, b200md_stream_t stream) {
  if (!positions || !cell || !counts || !edges_per_structure || !status || n_samples < 0 || n_atoms <= 0 ||
      !(radial_cutoff > 0.f) || (mode != B200MD_GRAPH_UNIQUE_PAIRS && mode != B200MD_GRAPH_PER_IMAGE))
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_atoms > 4096) return B200MD_ERR_UNSUPPORTED;
  if (n_samples == 0) return B200MD_OK;
  const int threads = n_atoms >= 256 ? 256 : (n_atoms >= 64 ? 128 : 64);
  const size_t smem = (size_t)(3 * n_atoms + 81) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(radius_graph_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, as_stream(stream)); radius_graph_count_kernel<<<(unsigned)n_samples, threads, smem, as_stream(stream)>>>(
      positions, cell, positions_are_relative, radial_cutoff, radial_cutoff * radial_cutoff, n_atoms, mode,
      counts, edges_per_structure, status, hit_mask); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_radius_graph_count(const float* positions, const float* cell, int32_t positions_are_relative,
                              float radial_cutoff, int64_t n_samples, int32_t n_atoms, int32_t mode,
                              int32_t* counts, int64_t* edges_per_structure, int32_t* status,
                              b200md_stream_t stream) {
  return radius_graph_count_impl(positions, cell, positions_are_relative, radial_cutoff, n_samples, n_atoms, mode, counts,
                                 edges_per_structure, status, nullptr, stream);
}

int b200md_radius_graph_count_masked(const float* positions, const float* cell, int32_t positions_are_relative,
                                     float radial_cutoff, int64_t n_samples, int32_t n_atoms, int32_t* counts,
                                     int64_t* edges_per_structure, int32_t* status, uint32_t* hit_mask,
                                     b200md_stream_t stream) {
  if (!hit_mask) return B200MD_ERR_INVALID_ARGUMENT;
  return radius_graph_count_impl(positions, cell, positions_are_relative, radial_cutoff, n_samples, n_atoms,
                                 B200MD_GRAPH_UNIQUE_PAIRS, counts, edges_per_structure, status, hit_mask, stream);
}

int b200md_radius_graph_fill_from_mask(const uint32_t* hit_mask, int64_t n_samples, int32_t n_atoms,
                                       const int64_t* offsets, int64_t node_index_offset, int64_t* src_out,
                                       int64_t* dst_out, int64_t* edge_batch, b200md_stream_t stream) {
  if (!hit_mask || !offsets || !src_out || !dst_out || n_samples < 0 || n_atoms <= 0) return B200MD_ERR_INVALID_ARGUMENT;
  if (n_samples == 0) return B200MD_OK;
  const int64_t n_rows = n_samples * n_atoms;
  const int warps_per_block = 8;
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, as_stream(stream));
    radius_graph_fill_from_mask_kernel<<<(unsigned)((n_rows + warps_per_block - 1) / warps_per_block), warps_per_block * 32, 0, as_stream(stream)>>>(
        hit_mask, n_rows, n_atoms, offsets, node_index_offset, src_out, dst_out, edge_batch); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_radius_graph_fill(const float* positions, const float* cell, int32_t positions_are_relative,
                             float radial_cutoff, int64_t n_samples, int32_t n_atoms, int32_t mode,
                             const int64_t* offsets, int64_t node_index_offset, int64_t* src_out,
                             int64_t* dst_out, float* shifts_out, int64_t* edge_batch,
                             b200md_stream_t stream) {
  if (!positions || !cell || !offsets || !src_out || !dst_out || n_samples < 0 || n_atoms <= 0 ||
      !(radial_cutoff > 0.f) || (mode != B200MD_GRAPH_UNIQUE_PAIRS && mode != B200MD_GRAPH_PER_IMAGE))
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_atoms > 4096) return B200MD_ERR_UNSUPPORTED;
  if (n_samples == 0) return B200MD_OK;
  int threads = n_atoms >= 256 ? 256 : (n_atoms >= 64 ? 128 : 64);
  if (mode == B200MD_GRAPH_PER_IMAGE && n_atoms > 512) threads = 128;
  size_t smem = (size_t)(3 * n_atoms + 81) * sizeof(float);
  if (mode == B200MD_GRAPH_PER_IMAGE) smem += (size_t)(threads / 32) * 2 * n_atoms * sizeof(float);
  if (smem > 200 * 1024) return B200MD_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(radius_graph_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, as_stream(stream)); radius_graph_fill_kernel<<<(unsigned)n_samples, threads, smem, as_stream(stream)>>>(
      positions, cell, positions_are_relative, radial_cutoff, radial_cutoff * radial_cutoff, n_atoms, mode, offsets,
      node_index_offset, src_out, dst_out, shifts_out, edge_batch); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

}  // extern "C"
