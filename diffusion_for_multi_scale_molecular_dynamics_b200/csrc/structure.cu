// Structure statistics built on the periodic-image sweep of graph.cu.
//
// b200md_pair_distances_{count,fill} replace compute_distances_in_batch (utils/structure_utils.py:41-121): all
// interatomic distances 0 < d <= max_distance over the 27 periodic images, in the reference's native order
// (structure, image, source atom, ascending distance).  It is the sample statistic of the reference's structure
// Kolmogorov-Smirnov metric (models/axl_diffusion_lightning_model.py:531,633) and of the trajectory-level parity tests.
// Same rounding chain as the radius graph (d2 in fp32 without FMA contraction); the reference filters on the distance
// itself, `0 < sqrt(d2) <= max_distance` (structure_utils.py:113-121), so the test is made on the IEEE square root.
#include "common.cuh"
#include "graph.cuh"

namespace b200md {

__device__ __forceinline__ bool distance_hit(float xi, float yi, float zi, float xj, float yj, float zj, float lx,
                                             float ly, float lz, float rc, float& d2, float& d) {
  image_hit(xi, yi, zi, xj, yj, zj, lx, ly, lz, 0.f, d2);
  d = __fsqrt_rn(d2);
  return (d > 0.0f) && (d <= rc);
}

__global__ void pair_distances_count_kernel(const float* __restrict__ positions, const float* __restrict__ cell,
                                            int relative, float rc, int32_t N, int32_t* __restrict__ counts,
                                            int64_t* __restrict__ per_structure) {
  extern __shared__ float smem[];
  float* px = smem; float* py = px + N; float* pz = py + N;
  float* lat = pz + N;  // [27][3]
  __shared__ float s_cell[9];
  __shared__ unsigned long long s_total;
  __shared__ ImageCull s_cull;
  const int64_t b = blockIdx.x;
  stage_structure(positions + b * N * 3, cell + b * 9, relative, N, px, py, pz, lat, s_cell);
  if (threadIdx.x == 0) s_total = 0ull;
  __syncthreads();
  image_cull_setup(&s_cull, px, py, pz, N, s_cell);   // graph.cuh: rows (image, atom) that cannot reach the cutoff are skipped
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  unsigned long long local = 0;
  for (int r = warp; r < 27 * N; r += nwarps) {
    const int img = r / N, i = r - img * N;
    const float xi = px[i], yi = py[i], zi = pz[i];
    const float lx = lat[img * 3], ly = lat[img * 3 + 1], lz = lat[img * 3 + 2];
    int n = 0;
    if (!image_can_hit(&s_cull, xi, yi, zi, img, rc)) {
      if (lane == 0) counts[(b * 27 + img) * N + i] = 0;
      continue;
    }
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      float d2, d;
      if (j < N) hit = distance_hit(xi, yi, zi, px[j], py[j], pz[j], lx, ly, lz, rc, d2, d);
      n += __popc(__ballot_sync(0xffffffffu, hit));
    }
    if (lane == 0) { counts[(b * 27 + img) * N + i] = n; local += n; }
  }
  if (lane == 0 && local) atomicAdd(&s_total, local);
  __syncthreads();
  if (threadIdx.x == 0) per_structure[b] = (int64_t)s_total;
}

__global__ void pair_distances_fill_kernel(const float* __restrict__ positions, const float* __restrict__ cell,
                                           int relative, float rc, int32_t N, const int64_t* __restrict__ offsets,
                                           float* __restrict__ distances_out) {
  extern __shared__ float smem[];
  float* px = smem; float* py = px + N; float* pz = py + N;
  float* lat = pz + N;          // [27][3]
  float* scratch = lat + 81;    // per-warp [N] d2, [N] d, [N] j
  __shared__ float s_cell[9];
  __shared__ ImageCull s_cull;
  const int64_t b = blockIdx.x;
  stage_structure(positions + b * N * 3, cell + b * 9, relative, N, px, py, pz, lat, s_cell);
  __syncthreads();
  image_cull_setup(&s_cull, px, py, pz, N, s_cell);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  float* wd2 = scratch + (size_t)warp * 3 * N;
  float* wd = wd2 + N;
  int* wj = reinterpret_cast<int*>(wd + N);
  for (int r = warp; r < 27 * N; r += nwarps) {
    const int img = r / N, i = r - img * N;
    const float xi = px[i], yi = py[i], zi = pz[i];
    const float lx = lat[img * 3], ly = lat[img * 3 + 1], lz = lat[img * 3 + 2];
    int cnt = 0;
    if (!image_can_hit(&s_cull, xi, yi, zi, img, rc)) continue;   // an empty row
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      float d2 = 0.f, d = 0.f;
      if (j < N) hit = distance_hit(xi, yi, zi, px[j], py[j], pz[j], lx, ly, lz, rc, d2, d);
      const unsigned ballot = __ballot_sync(0xffffffffu, hit);
      if (hit) { const int k = cnt + __popc(ballot & lt_mask); wd2[k] = d2; wd[k] = d; wj[k] = j; }
      cnt += __popc(ballot);
    }
    __syncwarp();
    const int64_t pos = offsets[(b * 27 + img) * N + i];
    // rank by (d2, j): ascending squared distance like Kmin_argKmin (structure_utils.py:104-106)
    for (int k = lane; k < cnt; k += 32) {
      const float dk = wd2[k]; const int jk = wj[k];
      int rank = 0;
      for (int m = 0; m < cnt; ++m) {
        const float dm = wd2[m];
        rank += (dm < dk) || (dm == dk && wj[m] < jk);
      }
      distances_out[pos + rank] = wd[k];
    }
    __syncwarp();
  }
}

// ---- repulsive pseudo-force of ForceFieldAugmentedScoreNetwork --------------------------------------------------
// models/score_networks/force_field_augmented_score_network.py:44-236:
//   cell = diag(clip(L[:3], min_box_size)); p = x @ cell; for every edge (i -> image of j) with 0 < d2 <= r0^2:
//   disp = (p_j - p_i) + shift, r = |disp|, F_i += 2 s (r - r0) / (r + 1e-8) disp; out = [add_to +] F @ inverse(cell).
// The reference builds the per-image adjacency (3 KeOps reductions + nonzero), gathers, and scatter_adds one Cartesian
// component at a time; here one warp owns a source atom of a structure staged in shared memory and accumulates its
// force in registers over the 27 images -- no edge list exists.
__global__ void repulsive_pseudo_force_kernel(const float* __restrict__ x_rel, const float* __restrict__ lattice,
                                              float min_box, float rc, float rc2, float strength, int32_t N,
                                              const float* __restrict__ add_to, float* __restrict__ out,
                                              int32_t* __restrict__ status) {
  extern __shared__ float smem[];
  float* px = smem; float* py = px + N; float* pz = py + N;
  float* lat = pz + N;  // [27][3]
  __shared__ float s_cell[9];
  __shared__ ImageCull s_cull;
  const int64_t b = blockIdx.x;
  if (threadIdx.x < 9) {
    const int r = threadIdx.x / 3, c = threadIdx.x % 3;
    s_cell[threadIdx.x] = (r == c) ? fmaxf(lattice[b * 6 + r], min_box) : 0.0f;   // basis_transformations.py:253-256
  }
  __syncthreads();
  stage_structure_with_cell(x_rel + b * N * 3, 1, N, px, py, pz, lat, s_cell);
  if (threadIdx.x == 0 && !(shortest_cell_crossing_distance(s_cell) > rc))
    atomicOr(status, (int)B200MD_STATUS_CUTOFF_TOO_LARGE);
  __syncthreads();
  image_cull_setup(&s_cull, px, py, pz, N, s_cell);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = warp; i < N; i += nwarps) {
    const float xi = px[i], yi = py[i], zi = pz[i];
    float fx = 0.f, fy = 0.f, fz = 0.f;
    for (unsigned m = image_candidates(&s_cull, xi, yi, zi, rc); m != 0; m &= m - 1) {   // images in ascending order
      const int img = __ffs(m) - 1;
      const float lx = lat[img * 3], ly = lat[img * 3 + 1], lz = lat[img * 3 + 2];
      for (int j = lane; j < N; j += 32) {
        float d2;
        if (image_hit(xi, yi, zi, px[j], py[j], pz[j], lx, ly, lz, rc2, d2)) {
          const float dx = (px[j] - xi) + lx, dy = (py[j] - yi) + ly, dz = (pz[j] - zi) + lz;   // :150-154
          const float r = sqrtf(dx * dx + dy * dy + dz * dz);
          const float pref = 2.0f * strength * (r - rc) / (r + 1.0e-8f);                       // :104-108
          fx += pref * dx; fy += pref * dy; fz += pref * dz;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      fx += __shfl_down_sync(0xffffffffu, fx, o);
      fy += __shfl_down_sync(0xffffffffu, fy, o);
      fz += __shfl_down_sync(0xffffffffu, fz, o);
    }
    if (lane == 0) {
      const int64_t o3 = (b * N + i) * 3;
      // relative force = F @ inverse(cell): the cell is diagonal here (reference: "Only orthogonal boxes")
      float rx = fx / s_cell[0], ry = fy / s_cell[4], rz = fz / s_cell[8];
      if (add_to) { rx += add_to[o3]; ry += add_to[o3 + 1]; rz += add_to[o3 + 2]; }
      out[o3] = rx; out[o3 + 1] = ry; out[o3 + 2] = rz;
    }
  }
}

// ---- Stillinger-Weber potential energy --------------------------------------------------------------------------
// The reference evaluates sample energies by running LAMMPS (`pair_style sw`) one structure at a time
// (oracle/lammps_energy_oracle.py:60-193); the energies feed its energy Kolmogorov-Smirnov metric.  Same functional
// form here (LAMMPS pair_sw, Stillinger & Weber, PRB 31, 5262 (1985)), one CTA per structure:
//   E = sum_{i<j} phi2(r_ij) + sum_i sum_{j<k} phi3(r_ij, r_ik, theta_jik)
//   phi2 = A eps [B (sigma/r)^p - (sigma/r)^q] exp(sigma / (r - a sigma))                      entry (i, j, j)
//   phi3 = lambda eps (cos theta - cos theta0)^2 exp(gamma_ij sigma_ij / (r_ij - a_ij sigma_ij))
//                                                exp(gamma_ik sigma_ik / (r_ik - a_ik sigma_ik))    entries (i,j,k),(i,j,j),(i,k,k)
// A warp owns a centre atom: it compacts the neighbours inside the pair cutoffs over the 27 images into a shared-memory
// list (ballot), then its lanes split the pair terms and the j < k triplets.
constexpr int kSwMaxNeighbours = 96;
constexpr int kSwEntryFloats = 10;   // cut, cutsq, sigma, sigma*gamma, lambda*eps, costheta0, c5, c6, p, q
struct SwEntry { float cut, cutsq, sigma, sigma_gamma, lambda_epsilon, costheta, c5, c6, powerp, powerq; };

__global__ void stillinger_weber_energy_kernel(const float* __restrict__ x_rel, const float* __restrict__ lattice,
                                               const int64_t* __restrict__ atom_types, const float* __restrict__ params,
                                               int32_t n_types, float max_cut, float min_box, int32_t N,
                                               float* __restrict__ energies, int32_t* __restrict__ status) {
  extern __shared__ float smem[];
  float* px = smem; float* py = px + N; float* pz = py + N;
  float* lat = pz + N;                                    // [27][3]
  int* types = reinterpret_cast<int*>(lat + 81);          // [N]
  float* lists = reinterpret_cast<float*>(types + N);     // per warp: dx, dy, dz, r [kSwMaxNeighbours] + type
  __shared__ float s_cell[9];
  __shared__ double s_energy;
  __shared__ int s_skip;
  __shared__ ImageCull s_cull;
  const int64_t b = blockIdx.x;
  if (threadIdx.x < 9) {
    const int r = threadIdx.x / 3, c = threadIdx.x % 3;
    s_cell[threadIdx.x] = (r == c) ? lattice[b * 6 + r] : 0.0f;
  }
  if (threadIdx.x == 0) { s_energy = 0.0; s_skip = 0; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float shortest = fminf(s_cell[0], fminf(s_cell[4], s_cell[8]));
    if (shortest < min_box) s_skip = 1;   // "Got a box with a side length smaller than 3.0 Angstrom": energy 0 (:121-127)
    else if (!(shortest > max_cut)) atomicOr(status, (int)B200MD_STATUS_CUTOFF_TOO_LARGE);
  }
  stage_structure_with_cell(x_rel + b * N * 3, 1, N, px, py, pz, lat, s_cell);
  bool bad_type = false;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int64_t t = atom_types[b * N + i];
    bad_type |= (t < 0 || t >= n_types);
    types[i] = (int)(t < 0 ? 0 : (t >= n_types ? n_types - 1 : t));
  }
  if (bad_type) atomicOr(status, (int)B200MD_STATUS_A_OUT_OF_RANGE);
  __syncthreads();
  if (s_skip) { if (threadIdx.x == 0) energies[b] = 0.0f; return; }
  image_cull_setup(&s_cull, px, py, pz, N, s_cell);
  const SwEntry* table = reinterpret_cast<const SwEntry*>(params);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  float* ndx = lists + (size_t)warp * 5 * kSwMaxNeighbours;
  float* ndy = ndx + kSwMaxNeighbours; float* ndz = ndy + kSwMaxNeighbours; float* nr = ndz + kSwMaxNeighbours;
  int* nt = reinterpret_cast<int*>(nr + kSwMaxNeighbours);
  float local = 0.f;
  for (int i = warp; i < N; i += nwarps) {
    const float xi = px[i], yi = py[i], zi = pz[i];
    const int ti = types[i];
    int cnt = 0;
    for (unsigned m = image_candidates(&s_cull, xi, yi, zi, max_cut); m != 0; m &= m - 1) {   // images in ascending order
      const int img = __ffs(m) - 1;
      const float lx = lat[img * 3], ly = lat[img * 3 + 1], lz = lat[img * 3 + 2];
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        float dx = 0.f, dy = 0.f, dz = 0.f, d2 = 0.f;
        int tj = 0;
        if (j < N) {
          tj = types[j];
          dx = (px[j] + lx) - xi; dy = (py[j] + ly) - yi; dz = (pz[j] + lz) - zi;
          d2 = dx * dx + dy * dy + dz * dz;
          hit = (d2 > 0.f) && (d2 < table[(ti * n_types + tj) * n_types + tj].cutsq);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const int k = cnt + __popc(ballot & lt_mask);
          if (k < kSwMaxNeighbours) { ndx[k] = dx; ndy[k] = dy; ndz[k] = dz; nr[k] = sqrtf(d2); nt[k] = tj; }
        }
        cnt += __popc(ballot);
      }
    }
    if (cnt > kSwMaxNeighbours) { if (lane == 0) atomicOr(status, (int)B200MD_STATUS_EDGE_OVERFLOW); cnt = kSwMaxNeighbours; }
    __syncwarp();
    for (int k = lane; k < cnt; k += 32) {   // pair terms, half to each end
      const SwEntry e = table[(ti * n_types + nt[k]) * n_types + nt[k]];
      const float r = nr[k];
      const float rp = (e.powerp == 4.0f) ? 1.0f / ((r * r) * (r * r)) : powf(r, -e.powerp);
      const float rq = (e.powerq == 0.0f) ? 1.0f : powf(r, -e.powerq);
      local += 0.5f * (e.c5 * rp - e.c6 * rq) * expf(e.sigma / (r - e.cut));
    }
    for (int ja = 0; ja < cnt; ++ja) {       // triplets centred on i
      const int ta = nt[ja];
      const SwEntry eij = table[(ti * n_types + ta) * n_types + ta];
      const float r1 = nr[ja], ax = ndx[ja], ay = ndy[ja], az = ndz[ja];
      const float exp1 = expf(eij.sigma_gamma / (r1 - eij.cut));
      for (int kb = ja + 1 + lane; kb < cnt; kb += 32) {
        const int tb = nt[kb];
        const SwEntry eik = table[(ti * n_types + tb) * n_types + tb];
        const SwEntry eijk = table[(ti * n_types + ta) * n_types + tb];
        const float r2 = nr[kb];
        const float cs = (ax * ndx[kb] + ay * ndy[kb] + az * ndz[kb]) / (r1 * r2);
        const float delcs = cs - eijk.costheta;
        local += eijk.lambda_epsilon * exp1 * expf(eik.sigma_gamma / (r2 - eik.cut)) * delcs * delcs;
      }
    }
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if (lane == 0) atomicAdd(&s_energy, (double)local);
  __syncthreads();
  if (threadIdx.x == 0) energies[b] = (float)s_energy;
}

}  // namespace b200md

using namespace b200md;

extern "C" {

int b200md_pair_distances_count(const float* positions, const float* cell, int32_t positions_are_relative,
                                float max_distance, int64_t n_samples, int32_t n_atoms, int32_t* counts,
                                int64_t* distances_per_structure, b200md_stream_t stream) {
  if (n_samples == 0) return B200MD_OK;
  if (!positions || !cell || !counts || !distances_per_structure || n_samples < 0 || n_atoms <= 0 || !(max_distance > 0.f))
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_atoms > 4096) return B200MD_ERR_UNSUPPORTED;
  const int threads = n_atoms >= 256 ? 256 : (n_atoms >= 64 ? 128 : 64);
  const size_t smem = (size_t)(3 * n_atoms + 81) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(pair_distances_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, as_stream(stream)); pair_distances_count_kernel<<<(unsigned)n_samples, threads, smem, as_stream(stream)>>>(
      positions, cell, positions_are_relative, max_distance, n_atoms, counts, distances_per_structure); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_pair_distances_fill(const float* positions, const float* cell, int32_t positions_are_relative,
                               float max_distance, int64_t n_samples, int32_t n_atoms, const int64_t* offsets,
                               float* distances_out, b200md_stream_t stream) {
  if (n_samples == 0) return B200MD_OK;
  if (!positions || !cell || !offsets || !distances_out || n_samples < 0 || n_atoms <= 0 || !(max_distance > 0.f))
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_atoms > 4096) return B200MD_ERR_UNSUPPORTED;
  int threads = n_atoms >= 256 ? 256 : (n_atoms >= 64 ? 128 : 64);
  if (n_atoms > 512) threads = 128;
  const size_t smem = (size_t)(3 * n_atoms + 81) * sizeof(float) + (size_t)(threads / 32) * 3 * n_atoms * sizeof(float);
  if (smem > 200 * 1024) return B200MD_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(pair_distances_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, as_stream(stream)); pair_distances_fill_kernel<<<(unsigned)n_samples, threads, smem, as_stream(stream)>>>(
      positions, cell, positions_are_relative, max_distance, n_atoms, offsets, distances_out); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_repulsive_pseudo_force(const float* x_rel, const float* lattice, float min_box_size, float radial_cutoff,
                                  float strength, int64_t n_samples, int32_t n_atoms, const float* add_to_or_null,
                                  float* out, int32_t* status, b200md_stream_t stream) {
  if (!x_rel || !lattice || !out || !status || n_samples < 0 || n_atoms <= 0 || !(radial_cutoff > 0.f) || !(strength > 0.f))
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_atoms > 4096) return B200MD_ERR_UNSUPPORTED;
  if (n_samples == 0) return B200MD_OK;
  const int threads = n_atoms >= 256 ? 256 : (n_atoms >= 64 ? 128 : 64);
  const size_t smem = (size_t)(3 * n_atoms + 81) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(repulsive_pseudo_force_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, as_stream(stream)); repulsive_pseudo_force_kernel<<<(unsigned)n_samples, threads, smem, as_stream(stream)>>>(
      x_rel, lattice, min_box_size, radial_cutoff, radial_cutoff * radial_cutoff, strength, n_atoms, add_to_or_null, out, status); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_stillinger_weber_energy(const float* x_rel, const float* lattice, const int64_t* atom_types,
                                   const float* params, int32_t n_types, float max_cutoff, float min_box_size,
                                   int64_t n_samples, int32_t n_atoms, float* energies_out, int32_t* status,
                                   b200md_stream_t stream) {
  if (!x_rel || !lattice || !atom_types || !params || !energies_out || !status || n_samples < 0 || n_atoms <= 0 ||
      n_types <= 0 || !(max_cutoff > 0.f))
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_atoms > 4096) return B200MD_ERR_UNSUPPORTED;
  if (n_samples == 0) return B200MD_OK;
  const int threads = n_atoms >= 256 ? 256 : 128;
  const size_t smem = (size_t)(3 * n_atoms + 81) * sizeof(float) + (size_t)n_atoms * sizeof(int) +
                      (size_t)(threads / 32) * 5 * kSwMaxNeighbours * sizeof(float);
  if (smem > 200 * 1024) return B200MD_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(stillinger_weber_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  { ProfileScope ps__(B200MD_KERNEL_GRAPH, as_stream(stream)); stillinger_weber_energy_kernel<<<(unsigned)n_samples, threads, smem, as_stream(stream)>>>(
      x_rel, lattice, atom_types, params, n_types, max_cutoff, min_box_size, n_atoms, energies_out, status); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

}  // extern "C"
