// Device helpers of the periodic radius graph (see graph.cu).
#pragma once
#include "common.cuh"

namespace b200md {

// squared distance to one periodic image, reference rounding chain (utils/neighbors.py:242-244,170):
//   shifted = p_j + lat ; diff = p_i - shifted ; d2 = (dx*dx + dy*dy) + dz*dz     -- fp32, no FMA.
__device__ __forceinline__ bool image_hit(float xi, float yi, float zi, float xj, float yj, float zj,
                                          float lx, float ly, float lz, float rc2, float& d2) {
  const float dx = __fsub_rn(xi, __fadd_rn(xj, lx));
  const float dy = __fsub_rn(yi, __fadd_rn(yj, ly));
  const float dz = __fsub_rn(zi, __fadd_rn(zj, lz));
  d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return (d2 > 0.0f) && (d2 <= rc2);  // neighbors.py:188-190
}

__device__ __forceinline__ int first_image_hit(float xi, float yi, float zi, float xj, float yj, float zj,
                                               const float* __restrict__ lat, float rc2) {
#pragma unroll 1
  for (int img = 0; img < 27; ++img) {
    float d2;
    if (image_hit(xi, yi, zi, xj, yj, zj, lat[img * 3], lat[img * 3 + 1], lat[img * 3 + 2], rc2, d2)) return img;
  }
  return -1;
}

__device__ __forceinline__ bool any_image_hit(float xi, float yi, float zi, float xj, float yj, float zj,
                                              const float* __restrict__ lat, float rc2) {
  return first_image_hit(xi, yi, zi, xj, yj, zj, lat, rc2) >= 0;
}

// Orthogonal cell with rc < L / 2 on every axis (the EGNN box: L' = max(L, 2.2 rc), egnn_score_network.py:236-240):
// at most ONE periodic image of j can lie inside the cutoff of i, and it is the one that minimises |dx|, |dy|, |dz|
// axis by axis.  With a diagonal cell the image vector is (n0 Lx, n1 Ly, n2 Lz) exactly, so d2(n) = fl(fl(a_n0 + b_n1) +
// c_n2) with a_n = fl(fl(x_i - fl(x_j + n Lx))^2) depends on one image index per axis and is monotone in each term:
// the minimum over the 27 images is reached at the per-axis minima, with the same rounding chain as the full sweep.
// "Some image has 0 < d2 <= rc2" is then "0 < min d2 <= rc2": if the minimum is 0 (coincident atoms, or i == j) every
// other image is at least L - |d| > rc away.  One test instead of 27 -- the pruning a cell list would give at larger
// L / rc, obtained here from the minimum-image property.
__device__ __forceinline__ float min_image_square(float xi, float xj, float L) {
  const float dm = __fsub_rn(xi, __fadd_rn(xj, -L));
  const float d0 = __fsub_rn(xi, __fadd_rn(xj, 0.0f));
  const float dp = __fsub_rn(xi, __fadd_rn(xj, L));
  return fminf(__fmul_rn(d0, d0), fminf(__fmul_rn(dm, dm), __fmul_rn(dp, dp)));
}
__device__ __forceinline__ bool min_image_hit(float xi, float yi, float zi, float xj, float yj, float zj, float Lx,
                                              float Ly, float Lz, float rc2) {
  const float d2 = __fadd_rn(__fadd_rn(min_image_square(xi, xj, Lx), min_image_square(yi, yj, Ly)),
                             min_image_square(zi, zj, Lz));
  return (d2 > 0.0f) && (d2 <= rc2);
}

// Shortest distance between opposite faces of the cell (utils/neighbors.py:310-351).
__device__ __forceinline__ float shortest_cell_crossing_distance(const float* a) {
  const float c12x = a[1] * a[5] - a[2] * a[4], c12y = a[2] * a[3] - a[0] * a[5], c12z = a[0] * a[4] - a[1] * a[3];
  const float c13x = a[1] * a[8] - a[2] * a[7], c13y = a[2] * a[6] - a[0] * a[8], c13z = a[0] * a[7] - a[1] * a[6];
  const float c23x = a[4] * a[8] - a[5] * a[7], c23y = a[5] * a[6] - a[3] * a[8], c23z = a[3] * a[7] - a[4] * a[6];
  const float vol = fabsf(c12x * a[6] + c12y * a[7] + c12z * a[8]);
  const float d12 = vol / sqrtf(c12x * c12x + c12y * c12y + c12z * c12z);
  const float d13 = vol / sqrtf(c13x * c13x + c13y * c13y + c13z * c13z);
  const float d23 = vol / sqrtf(c23x * c23x + c23y * c23y + c23z * c23z);
  return fminf(d12, fminf(d13, d23));
}

// Stage one structure: Cartesian positions (SoA) and the 27 lattice image vectors n @ cell with n in
// {-1,0,1}^3 in itertools.product order (utils/lattice_utils.py:24-27).  s_cell must already be loaded.
__device__ __forceinline__ void stage_structure_with_cell(const float* __restrict__ pos, int relative,
                                                          int N, float* px, float* py, float* pz,
                                                          float* lat, const float* s_cell) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float a = pos[i * 3], b = pos[i * 3 + 1], c = pos[i * 3 + 2];
    if (relative) {
      // p = x @ cell (utils/basis_transformations.py:56), left-to-right fp32 accumulation
      px[i] = __fadd_rn(__fadd_rn(__fmul_rn(a, s_cell[0]), __fmul_rn(b, s_cell[3])), __fmul_rn(c, s_cell[6]));
      py[i] = __fadd_rn(__fadd_rn(__fmul_rn(a, s_cell[1]), __fmul_rn(b, s_cell[4])), __fmul_rn(c, s_cell[7]));
      pz[i] = __fadd_rn(__fadd_rn(__fmul_rn(a, s_cell[2]), __fmul_rn(b, s_cell[5])), __fmul_rn(c, s_cell[8]));
    } else {
      px[i] = a; py[i] = b; pz[i] = c;
    }
  }
  for (int t = threadIdx.x; t < 81; t += blockDim.x) {
    const int img = t / 3, c = t - img * 3;
    const float n0 = (float)(img / 9 - 1), n1 = (float)((img / 3) % 3 - 1), n2 = (float)(img % 3 - 1);
    lat[t] = __fadd_rn(__fadd_rn(__fmul_rn(n0, s_cell[c]), __fmul_rn(n1, s_cell[3 + c])), __fmul_rn(n2, s_cell[6 + c]));
  }
}

__device__ __forceinline__ void stage_structure(const float* __restrict__ pos, const float* __restrict__ cell,
                                                int relative, int N, float* px, float* py, float* pz,
                                                float* lat, float* s_cell) {
  if (threadIdx.x < 9) s_cell[threadIdx.x] = cell[threadIdx.x];
  __syncthreads();
  stage_structure_with_cell(pos, relative, N, px, py, pz, lat, s_cell);
}

int launch_egnn_graph(const float* x_rel, const float* lattice, float clip_min, float rc, int64_t B,
                      int32_t N, int fully_connected, int64_t capacity, uint32_t* bitmask, int32_t* deg,
                      int32_t* struct_total, int32_t* struct_off, int32_t* n_edges, int32_t* row_ptr,
                      int32_t* edge_src, int32_t* edge_dst, int32_t* status, cudaStream_t st);

}  // namespace b200md
