// Device helpers of the periodic radius graph (see graph.cu).
#pragma once
#include <math_constants.h>

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

// ---- periodic-image culling: the slab form of a cell list -----------------------------------------------------------
// The reference tests every (i, j) pair against all 27 images.  Most (atom, image) combinations cannot produce a hit: with
// fractional coordinates s = p . inverse(cell), the image n of atom j lies at s_j + n, and two points whose fractional
// coordinates differ by ds along axis a are at least |ds| h_a apart (h_a = distance between the cell faces normal to
// reciprocal vector b_a = 1 / |b_a|).  With [smin_a, smax_a] the range of the structure's own fractional coordinates,
//     gap_a(i, n_a) = max(0, smin_a + n_a - s_i,a , s_i,a - smax_a - n_a) h_a
// bounds the distance from atom i to EVERY atom of image n from below; if it exceeds the cutoff on any axis the whole
// (i, image) row is skipped.  Valid for triclinic cells and for positions outside the cell (no wrapping is assumed);
// conservative (1 % slack on the cutoff, far above the fp32 error of s), so the surviving tests -- made with the
// reference's rounding chain -- give bit-identical edge sets.  An atom is within rc of a given face with probability
// rc / L, so (1 + 2 rc / L)^3 of the 27 images survive: 2.4 at L / rc = 5.8 (Stillinger-Weber cutoff in a 4x4x4 cell),
// 4.8 at L / rc = 2.9.
struct ImageCull {
  float b[9];              // reciprocal vectors, row a: s_a = p . b_a
  float h[3];              // 1 / |b_a|
  int smin[3], smax[3];    // bounds of the fractional coordinates (order-preserving integer encoding)
};
__device__ __forceinline__ int cull_encode(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float cull_decode(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// All threads of the CTA; px / py / pz and s_cell must be staged.  Ends with a __syncthreads().
__device__ __forceinline__ void image_cull_setup(ImageCull* c, const float* px, const float* py, const float* pz, int N,
                                                 const float* a /* s_cell */) {
  if (threadIdx.x == 0) {
    // inverse(cell): column a of the inverse = (a_{a+1} x a_{a+2}) / det
    const float c0x = a[4] * a[8] - a[5] * a[7], c0y = a[5] * a[6] - a[3] * a[8], c0z = a[3] * a[7] - a[4] * a[6];
    const float c1x = a[7] * a[2] - a[8] * a[1], c1y = a[8] * a[0] - a[6] * a[2], c1z = a[6] * a[1] - a[7] * a[0];
    const float c2x = a[1] * a[5] - a[2] * a[4], c2y = a[2] * a[3] - a[0] * a[5], c2z = a[0] * a[4] - a[1] * a[3];
    const float det = a[0] * c0x + a[1] * c0y + a[2] * c0z;
    const float inv = 1.0f / det;
    c->b[0] = c0x * inv; c->b[1] = c0y * inv; c->b[2] = c0z * inv;
    c->b[3] = c1x * inv; c->b[4] = c1y * inv; c->b[5] = c1z * inv;
    c->b[6] = c2x * inv; c->b[7] = c2y * inv; c->b[8] = c2z * inv;
    for (int k = 0; k < 3; ++k) {
      c->h[k] = rsqrtf(c->b[3 * k] * c->b[3 * k] + c->b[3 * k + 1] * c->b[3 * k + 1] + c->b[3 * k + 2] * c->b[3 * k + 2]);
      c->smin[k] = 0x7fffffff;
      c->smax[k] = (int)0x80000000;
    }
  }
  __syncthreads();
  float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    for (int k = 0; k < 3; ++k) {
      const float s = px[i] * c->b[3 * k] + py[i] * c->b[3 * k + 1] + pz[i] * c->b[3 * k + 2];
      lo[k] = fminf(lo[k], s); hi[k] = fmaxf(hi[k], s);
    }
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&c->smin[k], cull_encode(lo[k])); atomicMax(&c->smax[k], cull_encode(hi[k])); }
  }
  __syncthreads();
}
// Bit img of the result is set when image img of the structure MAY hold an atom within `rc` of the point (warp-uniform).
// A degenerate cell gives NaN gaps, which compare false: nothing is culled.
#ifndef B200MD_IMAGE_CULL
#define B200MD_IMAGE_CULL 1   // 0: every image is a candidate (the plain 27-image sweep, for A/B timing)
#endif
__device__ __forceinline__ unsigned image_candidates(const ImageCull* c, float xi, float yi, float zi, float rc) {
  if (!B200MD_IMAGE_CULL) return 0x7ffffffu;
  const float limit = rc * 1.01f;
  unsigned axis_ok[3];   // bit n: image index n - 1 along this axis survives
  for (int k = 0; k < 3; ++k) {
    const float s = xi * c->b[3 * k] + yi * c->b[3 * k + 1] + zi * c->b[3 * k + 2];
    const float smin = cull_decode(c->smin[k]), smax = cull_decode(c->smax[k]);
    unsigned ok = 0;
    for (int n = -1; n <= 1; ++n) {
      const float gap = fmaxf(fmaxf(smin + (float)n - s, s - smax - (float)n), 0.0f) * c->h[k];
      if (!(gap > limit)) ok |= 1u << (n + 1);
    }
    axis_ok[k] = ok;
  }
  unsigned mask = 0;
  for (int img = 0; img < 27; ++img)   // image index = 9 (n0 + 1) + 3 (n1 + 1) + (n2 + 1), as in stage_structure
    if (((axis_ok[0] >> (img / 9)) & (axis_ok[1] >> ((img / 3) % 3)) & (axis_ok[2] >> (img % 3))) & 1u) mask |= 1u << img;
  return mask;
}
// The same test for ONE image (the per-image row loops): three dot products instead of the whole 27-bit mask.
__device__ __forceinline__ bool image_can_hit(const ImageCull* c, float xi, float yi, float zi, int img, float rc) {
  if (!B200MD_IMAGE_CULL) return true;
  const float limit = rc * 1.01f;
  const int n[3] = {img / 9 - 1, (img / 3) % 3 - 1, img % 3 - 1};
  bool ok = true;
  for (int k = 0; k < 3; ++k) {
    const float s = xi * c->b[3 * k] + yi * c->b[3 * k + 1] + zi * c->b[3 * k + 2];
    const float gap = fmaxf(fmaxf(cull_decode(c->smin[k]) + (float)n[k] - s, s - cull_decode(c->smax[k]) - (float)n[k]), 0.0f) * c->h[k];
    ok = ok && !(gap > limit);
  }
  return ok;
}
// first hit among the candidate images, in image order (the image whose shift the reference reports)
__device__ __forceinline__ int first_image_hit_among(float xi, float yi, float zi, float xj, float yj, float zj,
                                                     const float* __restrict__ lat, float rc2, unsigned candidates) {
  for (unsigned m = candidates; m != 0; m &= m - 1) {
    const int img = __ffs(m) - 1;
    float d2;
    if (image_hit(xi, yi, zi, xj, yj, zj, lat[img * 3], lat[img * 3 + 1], lat[img * 3 + 2], rc2, d2)) return img;
  }
  return -1;
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
