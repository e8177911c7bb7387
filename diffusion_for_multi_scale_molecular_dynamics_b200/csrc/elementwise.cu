// HBM-bound kernels of the predictor-corrector step: noise draws, unit-cell wrap, coordinate / lattice
// update, atom-type (D3PM) update, RePaint overwrite, wrapped-Gaussian score, segment reductions.
// Reference functions are cited next to each kernel (paths relative to the reference package).
#include <math_constants.h>

#include "common.cuh"
#include "sampler.cuh"

namespace b200md {

// -------------------------------------------------------------------------------------------------
// Noise draws.  One thread per Philox block (4 values).
// -------------------------------------------------------------------------------------------------
enum DrawKind { kNormal = 0, kUniform = 1, kGumbel = 2 };

template <int KIND>
__global__ void draw_kernel(float* __restrict__ out, int64_t n_samples, int32_t units, int32_t unit_size,
                            float small_epsilon, uint64_t seed, uint32_t step_id, uint32_t draw_id,
                            uint64_t sample_offset) {
  const int32_t blocks_per_unit = (unit_size + 3) >> 2;
  const int64_t per_sample = (int64_t)units * blocks_per_unit;
  const int64_t total = n_samples * per_sample;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = t / per_sample;
    const int64_t r = t - s * per_sample;
    const int32_t unit = (int32_t)(r / blocks_per_unit);
    const int32_t blk = (int32_t)(r - (int64_t)unit * blocks_per_unit);
    Philox4 p = draw_block(seed, sample_offset + (uint64_t)s, (uint32_t)r, step_id, draw_id);
    float v[4];
    if (KIND == kNormal) {
      normal4(p, v);
    } else {
      v[0] = u01(p.x); v[1] = u01(p.y); v[2] = u01(p.z); v[3] = u01(p.w);
      if (KIND == kGumbel) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = gumbel_from_uniform(v[i], small_epsilon);
      }
    }
    float* o = out + (s * units + unit) * (int64_t)unit_size + blk * 4;
    const int n = min(4, unit_size - blk * 4);
    for (int i = 0; i < n; ++i) o[i] = v[i];
  }
}

template <int KIND>
static int launch_draw(float* out, int64_t n_samples, int32_t units, int32_t unit_size, float eps,
                       uint64_t seed, uint32_t step_id, uint32_t draw_id, uint64_t sample_offset,
                       b200md_stream_t stream) {
  if (!out || n_samples < 0 || units <= 0 || unit_size <= 0) return B200MD_ERR_INVALID_ARGUMENT;
  if (n_samples == 0) return B200MD_OK;
  int64_t total = n_samples * (int64_t)units * ((unit_size + 3) / 4);
  int threads = 256;
  int64_t blocks = ceil_div64(total, threads);
  if (blocks > 148 * 16) blocks = 148 * 16;
  { ProfileScope ps__(B200MD_KERNEL_OTHER, as_stream(stream)); draw_kernel<KIND><<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(
      out, n_samples, units, unit_size, eps, seed, step_id, draw_id, sample_offset); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

// -------------------------------------------------------------------------------------------------
// map_relative_coordinates_to_unit_cell   utils/basis_transformations.py:95-119
// -------------------------------------------------------------------------------------------------
__global__ void wrap_kernel(const float* x, float* out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = wrap_unit(x[i]);
}
// same, four values per thread with 128-bit accesses (n % 4 == 0, 16-byte aligned pointers)
__global__ void wrap_kernel_v4(const float4* x, float4* out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    out[i] = make_float4(wrap_unit(v.x), wrap_unit(v.y), wrap_unit(v.z), wrap_unit(v.w));
  }
}

// -------------------------------------------------------------------------------------------------
// Coordinate update, generators/langevin_generator.py:194-200:
//   updated = (x + (score_weight * s) / sigma) + noise_weight * z ; wrap
// One thread per atom.  Rounding chain kept explicit (no FMA contraction).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pc_update_value(float x, float s, float z, float sw, float sigma,
                                                 float nw) {
  float a = __fdiv_rn(__fmul_rn(sw, s), sigma);
  float b = __fmul_rn(nw, z);
  return wrap_unit(__fadd_rn(__fadd_rn(x, a), b));
}

// step-dependent arguments of a graph-replayed launch (sampler.cuh)
__device__ __forceinline__ void pc_step_weights(const PcStepScalars* cur, int phase, float& sw, float& sigma, float& nw,
                                                uint32_t& step_id) {
  const bool predictor = phase == 0;
  sw = predictor ? cur->g2 : cur->eps;
  sigma = predictor ? cur->sigma_pred : cur->sigma_corr;
  nw = predictor ? cur->g : cur->sqrt_2eps;
  step_id = cur->step_id + (uint32_t)phase;
}

template <int D>
__global__ void pc_coordinates_update_kernel(const float* x, const float* __restrict__ s,
                                             const float* __restrict__ z, float sw, float sigma,
                                             float nw, int64_t n_samples, int32_t n_atoms,
                                             uint64_t seed, uint32_t step_id,
                                             uint64_t sample_offset, float* out, const PcStepScalars* cur, int phase) {
  if (cur != nullptr) pc_step_weights(cur, phase, sw, sigma, nw, step_id);
  const int64_t total = n_samples * n_atoms;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    float zz[4];
    if (z != nullptr) {
#pragma unroll
      for (int d = 0; d < D; ++d) zz[d] = z[t * D + d];
    } else {
      const int64_t smp = t / n_atoms;
      const int32_t atom = (int32_t)(t - smp * n_atoms);
      Philox4 p = draw_block(seed, sample_offset + (uint64_t)smp, (uint32_t)atom, step_id,
                             B200MD_DRAW_COORD_NORMAL);
      normal4(p, zz);
    }
    float xv[D], sv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { xv[d] = x[t * D + d]; sv[d] = s[t * D + d]; }
#pragma unroll
    for (int d = 0; d < D; ++d) out[t * D + d] = pc_update_value(xv[d], sv[d], zz[d], sw, sigma, nw);
  }
}

// D = 3, four atoms (twelve consecutive floats = three float4) per thread: every access is a full 128-bit vector, so
// the kernel streams at HBM speed instead of issuing stride-3 scalar loads.  Same arithmetic and the same Philox
// counters (one block per (sample, atom)) as pc_coordinates_update_kernel<3>; needs (n_samples * n_atoms) % 4 == 0.
__global__ void pc_coordinates_update_v4_kernel(const float4* x, const float4* __restrict__ s,
                                                const float4* __restrict__ z, float sw, float sigma, float nw,
                                                int64_t n_groups, int32_t n_atoms, uint64_t seed, uint32_t step_id,
                                                uint64_t sample_offset, float4* out, const PcStepScalars* cur, int phase) {
  if (cur != nullptr) pc_step_weights(cur, phase, sw, sigma, nw, step_id);
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    float xv[12], sv[12], zv[12];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 a = x[g * 3 + k], b = s[g * 3 + k];
      xv[4 * k] = a.x; xv[4 * k + 1] = a.y; xv[4 * k + 2] = a.z; xv[4 * k + 3] = a.w;
      sv[4 * k] = b.x; sv[4 * k + 1] = b.y; sv[4 * k + 2] = b.z; sv[4 * k + 3] = b.w;
    }
    if (z != nullptr) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 c = z[g * 3 + k];
        zv[4 * k] = c.x; zv[4 * k + 1] = c.y; zv[4 * k + 2] = c.z; zv[4 * k + 3] = c.w;
      }
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int64_t t = g * 4 + a;
        const int64_t smp = t / n_atoms;
        const int32_t atom = (int32_t)(t - smp * n_atoms);
        float zz[4];
        normal4(draw_block(seed, sample_offset + (uint64_t)smp, (uint32_t)atom, step_id, B200MD_DRAW_COORD_NORMAL), zz);
        zv[3 * a] = zz[0]; zv[3 * a + 1] = zz[1]; zv[3 * a + 2] = zz[2];
      }
    }
    float r[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) r[k] = pc_update_value(xv[k], sv[k], zv[k], sw, sigma, nw);
#pragma unroll
    for (int k = 0; k < 3; ++k) out[g * 3 + k] = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
  }
}

// ---- adaptive corrector (generators/adaptive_corrector.py:98-148) ------------------------------------------------
// eps = 2 (r ||z|| / max(||sigma s|| / sigma, small_eps))^2 with
//   ||sigma s|| = mean over samples of the L2 norm of the sample's whole score block (units x d values),
//   ||z||       = mean over (sample, unit) of the L2 norm of the d-vector.
// Two launches, both with a fixed reduction order (deterministic): per-sample partial norms, then one CTA.
__global__ void adaptive_norms_kernel(const float* __restrict__ s, const float* __restrict__ z, int32_t units, int32_t d,
                                      float* __restrict__ partial, int64_t n_samples) {
  __shared__ float sh_s[32], sh_z[32];
  const int64_t b = blockIdx.x;
  const float* sb = s + b * units * d;
  const float* zb = z + b * units * d;
  float ss = 0.f, zn = 0.f;
  for (int u = threadIdx.x; u < units; u += blockDim.x) {
    float zz = 0.f;
    for (int k = 0; k < d; ++k) {
      const float sv = sb[u * d + k], zv = zb[u * d + k];
      ss += sv * sv;
      zz += zv * zv;
    }
    zn += sqrtf(zz);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) { ss += __shfl_down_sync(0xffffffffu, ss, o); zn += __shfl_down_sync(0xffffffffu, zn, o); }
  if (lane == 0) { sh_s[warp] = ss; sh_z[warp] = zn; }
  __syncthreads();
  if (warp == 0) {
    const int nwarps = blockDim.x >> 5;
    ss = lane < nwarps ? sh_s[lane] : 0.f;
    zn = lane < nwarps ? sh_z[lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) { ss += __shfl_down_sync(0xffffffffu, ss, o); zn += __shfl_down_sync(0xffffffffu, zn, o); }
    if (lane == 0) { partial[b] = sqrtf(ss); partial[n_samples + b] = zn; }
  }
}

__global__ void adaptive_step_size_kernel(const float* __restrict__ partial, int64_t n_samples, int32_t units, float sigma,
                                          float corrector_r, float small_epsilon, float* __restrict__ step_out) {
  __shared__ float sh_s[32], sh_z[32];
  float ss = 0.f, zn = 0.f;
  for (int64_t b = threadIdx.x; b < n_samples; b += blockDim.x) { ss += partial[b]; zn += partial[n_samples + b]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) { ss += __shfl_down_sync(0xffffffffu, ss, o); zn += __shfl_down_sync(0xffffffffu, zn, o); }
  if (lane == 0) { sh_s[warp] = ss; sh_z[warp] = zn; }
  __syncthreads();
  if (warp == 0) {
    const int nwarps = blockDim.x >> 5;
    ss = lane < nwarps ? sh_s[lane] : 0.f;
    zn = lane < nwarps ? sh_z[lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) { ss += __shfl_down_sync(0xffffffffu, ss, o); zn += __shfl_down_sync(0xffffffffu, zn, o); }
    if (lane == 0) {
      const float score_norm = (ss / (float)n_samples) / sigma;               // adaptive_corrector.py:130-134
      const float z_norm = zn / ((float)n_samples * (float)units);            // :136
      const float ratio = corrector_r * z_norm / fmaxf(score_norm, small_epsilon);
      const float eps = 2.0f * ratio * ratio;                                  // :138-146
      step_out[0] = eps;
      step_out[1] = sqrtf(2.0f * eps);
    }
  }
}

// out = [wrap]((x + (w[0] * s) / sigma) + w[1] * z) with the two weights read from device memory
__global__ void pc_update_device_weights_kernel(const float* x, const float* __restrict__ s, const float* __restrict__ z,
                                                const float* __restrict__ weights, float sigma, int64_t n, int wrap,
                                                float* out) {
  const float sw = weights[0], nw = weights[1];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const float a = __fdiv_rn(__fmul_rn(sw, s[t]), sigma);
    const float v = __fadd_rn(__fadd_rn(x[t], a), __fmul_rn(nw, z[t]));
    out[t] = wrap ? wrap_unit(v) : v;
  }
}

// Lattice update, generators/langevin_generator.py:485-490 (no wrap).  One thread per (sample, Philox block of 4),
// the same block layout as draw_kernel with units = 1, unit_size = n_lp.
__global__ void pc_lattice_update_kernel(const float* l, const float* __restrict__ s,
                                         const float* __restrict__ z, float sw, float sigma_n, float nw,
                                         int64_t n_samples, int32_t n_lp, uint64_t seed,
                                         uint32_t step_id, uint64_t sample_offset,
                                         float* out) {
  const int32_t units = (n_lp + 3) / 4;
  const int64_t total = n_samples * units;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t smp = t / units;
    const int32_t unit = (int32_t)(t - smp * units);
    float zz[4] = {0.f, 0.f, 0.f, 0.f};
    if (z == nullptr) {
      Philox4 p = draw_block(seed, sample_offset + (uint64_t)smp, (uint32_t)unit, step_id,
                             B200MD_DRAW_LATTICE_NORMAL);
      normal4(p, zz);
    }
    for (int d = 0; d < 4; ++d) {
      int c = unit * 4 + d;
      if (c >= n_lp) break;
      int64_t i = smp * n_lp + c;
      float zv = (z != nullptr) ? z[i] : zz[d];
      float a = __fdiv_rn(__fmul_rn(sw, s[i]), sigma_n);
      out[i] = __fadd_rn(__fadd_rn(l[i], a), __fmul_rn(nw, zv));
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Atom-type update.  One CTA per sample (per-sample reductions: all-MASK flag, arg-max over atoms).
//   generators/langevin_generator.py:247-439, utils/d3pm_utils.py:64-150
// -------------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 32;

// CT > 0: the number of classes at compile time (every per-atom array in registers, loops unrolled); CT = 0: run time
// (the arrays are indexed dynamically and live in local memory: 3.4 x slower at C = 3).  Same operations in the same order.
template <int CT>
__global__ void atom_types_update_kernel(const float* __restrict__ logits, const int64_t* a_in,
                                         const float* __restrict__ q, const float* __restrict__ qbar,
                                         const float* __restrict__ qbar_tm1,
                                         const float* __restrict__ gumbel, const float* __restrict__ binary,
                                         int32_t flags, float eps, int32_t n_atoms, int32_t C_runtime,
                                         uint64_t seed, uint32_t step_id, uint64_t sample_offset,
                                         int64_t* a_out, float* __restrict__ probs_out,
                                         int32_t* __restrict__ status, const PcStepScalars* cur, int phase) {
  const int C = CT > 0 ? CT : C_runtime;
  constexpr int kArr = CT > 0 ? CT : kMaxClasses;
  if (cur != nullptr) {   // graph replay: flags, table rows and the draw key of THIS step (sampler.cuh)
    const bool predictor = phase == 0;
    flags = predictor ? cur->atom_flags_pred : cur->atom_flags_corr;
    const int64_t row = (int64_t)(predictor ? cur->q_index_pred : cur->q_index_corr) * C * C;
    q += row; qbar += row; qbar_tm1 += row;
    step_id = cur->step_id + (uint32_t)phase;
  }
  extern __shared__ unsigned char smem_raw[];
  float* s_q = reinterpret_cast<float*>(smem_raw);  // [3][C*C]
  float* s_best = s_q + 3 * C * C;                  // [n_atoms]  winning Gumbel value
  int32_t* s_new = reinterpret_cast<int32_t*>(s_best + n_atoms);  // [n_atoms] sampled type
  __shared__ int s_not_all_masked;
  __shared__ float s_red_val[32];
  __shared__ int s_red_idx[32];

  const int64_t smp = blockIdx.x;
  const int mask_class = C - 1;
  const bool greedy = flags & B200MD_ATOM_GREEDY;
  const bool one_transition = (flags & B200MD_ATOM_ONE_TRANSITION) != 0;
  const bool last_step = (flags & B200MD_ATOM_LAST_STEP) != 0;

  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    s_q[i] = q[i]; s_q[C * C + i] = qbar[i]; s_q[2 * C * C + i] = qbar_tm1[i];
  }
  if (threadIdx.x == 0) s_not_all_masked = 0;
  __syncthreads();
  const int64_t* a_row = a_in + smp * n_atoms;
  if (greedy) {
    int any = 0;
    for (int n = threadIdx.x; n < n_atoms; n += blockDim.x) any |= (a_row[n] != mask_class);
    if (any) s_not_all_masked = 1;  // benign race: all writers store 1
  }
  __syncthreads();
  const bool all_masked = (s_not_all_masked == 0);
  unsigned local_status = 0;

  for (int n = threadIdx.x; n < n_atoms; n += blockDim.x) {
    const int64_t node = smp * n_atoms + n;
    const int at = (int)a_row[n];
    float p0[kArr];
    // softmax -> clip -> renormalise   utils/d3pm_utils.py:140-149
    float mx = -CUDART_INF_F;
    _Pragma("unroll") for (int c = 0; c < C; ++c) { p0[c] = logits[node * C + c]; mx = fmaxf(mx, p0[c]); }
    float sum = 0.f;
    _Pragma("unroll") for (int c = 0; c < C; ++c) { p0[c] = expf(p0[c] - mx); sum += p0[c]; }
    float check = 0.f;
    _Pragma("unroll") for (int c = 0; c < C; ++c) { p0[c] = p0[c] / sum; check += p0[c]; }
    if (!(fabsf(check - 1.0f) <= 1e-5f + 1.3e-6f)) local_status |= B200MD_STATUS_BAD_LOGITS;
    float csum = 0.f;
    _Pragma("unroll") for (int c = 0; c < C; ++c) { p0[c] = fmaxf(p0[c], eps); csum += p0[c]; }
    _Pragma("unroll") for (int c = 0; c < C; ++c) p0[c] = p0[c] / csum;
    // p(a_{t-1} | a_t)   utils/d3pm_utils.py:102-124
    float den = 0.f;
    _Pragma("unroll") for (int j = 0; j < C; ++j) den += p0[j] * s_q[C * C + j * C + at];
    float prob[kArr];
    _Pragma("unroll") for (int i = 0; i < C; ++i) {
      float num1 = 0.f;
      _Pragma("unroll") for (int j = 0; j < C; ++j) num1 += p0[j] * s_q[2 * C * C + j * C + i];
      float num2 = s_q[i * C + at];
      prob[i] = (num1 * num2) / den;
    }
    // Gumbel noise
    float gum[kArr];
    if (gumbel != nullptr) {
      _Pragma("unroll") for (int c = 0; c < C; ++c) gum[c] = gumbel[node * C + c];
    } else {
      const int nblk = (C + 3) >> 2;
      for (int b = 0; b < nblk; ++b) {
        Philox4 p = draw_block(seed, sample_offset + (uint64_t)smp, (uint32_t)(n * nblk + b), step_id,
                               B200MD_DRAW_GUMBEL);
        float u[4] = {u01(p.x), u01(p.y), u01(p.z), u01(p.w)};
        for (int i = 0; i < 4 && b * 4 + i < C; ++i) gum[b * 4 + i] = gumbel_from_uniform(u[i], eps);
      }
    }
    if (greedy) {
      // generators/langevin_generator.py:382-439
      float bin;
      if (binary != nullptr) {
        bin = binary[node];
      } else {
        Philox4 p = draw_block(seed, sample_offset + (uint64_t)smp, (uint32_t)n, step_id,
                               B200MD_DRAW_BINARY);
        bin = u01(p.x);
      }
      const bool do_greedy = (!all_masked) && (bin > prob[mask_class]) && (at == mask_class);
      if (do_greedy) prob[mask_class] = 0.f;
      if (!all_masked)
        _Pragma("unroll") for (int c = 0; c < C; ++c) gum[c] = 0.f;
    }
    if (probs_out != nullptr)
      _Pragma("unroll") for (int c = 0; c < C; ++c) probs_out[node * C + c] = prob[c];
    float best = -CUDART_INF_F;
    int best_c = 0;
    _Pragma("unroll") for (int c = 0; c < C; ++c) {
      float v = logf(prob[c] + eps) + gum[c];
      if (v > best || c == 0) { best = v; best_c = c; }  // first maximum wins (torch.max)
    }
    s_best[n] = best;
    s_new[n] = best_c;
  }
  __syncthreads();

  int64_t* out_row = a_out + smp * n_atoms;
  if (one_transition) {
    // generators/langevin_generator.py:339-380: only the most likely proposed change happens.
    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int n = threadIdx.x; n < n_atoms; n += blockDim.x) {
      float v = (s_new[n] != (int)a_row[n]) ? s_best[n] : -CUDART_INF_F;
      if (v > bv || (v == bv && n < bi)) { bv = v; bi = n; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_down_sync(0xffffffffu, bv, o);
      int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_red_val[warp] = bv; s_red_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      const int nw = (blockDim.x + 31) >> 5;
      bv = lane < nw ? s_red_val[lane] : -CUDART_INF_F;
      bi = lane < nw ? s_red_idx[lane] : 0x7fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_down_sync(0xffffffffu, bv, o);
        int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) s_red_idx[0] = bi;
    }
    __syncthreads();
    const int winner = s_red_idx[0];  // argmax of an all -inf row is atom 0, whose type is unchanged
    for (int n = threadIdx.x; n < n_atoms; n += blockDim.x) {
      int v = (n == winner) ? s_new[n] : (int)a_row[n];
      out_row[n] = v;
      if (last_step && v == mask_class) local_status |= B200MD_STATUS_MASK_REMAINS;
    }
  } else {
    for (int n = threadIdx.x; n < n_atoms; n += blockDim.x) {
      int v = s_new[n];
      out_row[n] = v;
      if (last_step && v == mask_class) local_status |= B200MD_STATUS_MASK_REMAINS;
    }
  }
  if (local_status && status != nullptr) atomicOr(status, (int)local_status);
}

// -------------------------------------------------------------------------------------------------
// RePaint overwrite.  One thread per (sample, constrained atom).
//   generators/constrained_langevin_generator.py:74-82,120-163; data/diffusion/noising_transform.py:153-181;
//   noisers/relative_coordinates_noiser.py:58-67; noisers/atom_types_noiser.py:50-60
// -------------------------------------------------------------------------------------------------
__global__ void repaint_kernel(float* __restrict__ x, int64_t* __restrict__ a, const float* __restrict__ x0,
                               const int64_t* __restrict__ a0, const int64_t* __restrict__ idx,
                               int32_t nc, float sigma_i, const float* __restrict__ qbar,
                               const float* __restrict__ z, const float* __restrict__ u,
                               int64_t n_samples, int32_t n_atoms, int32_t D, int32_t C, uint64_t seed,
                               uint32_t step_id, uint64_t sample_offset, const PcStepScalars* cur) {
  if (cur != nullptr) {   // graph replay (sampler.cuh)
    sigma_i = cur->repaint_sigma;
    if (cur->q_index_repaint >= 0) qbar += (int64_t)cur->q_index_repaint * C * C;
    step_id = cur->step_id;
  }
  const int64_t total = n_samples * nc;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t smp = t / nc;
    const int32_t k = (int32_t)(t - smp * nc);
    const int64_t atom = idx ? idx[k] : k;
    const int64_t node = smp * n_atoms + atom;
    const int a_known = (int)a0[k];
    if (sigma_i < 0.f) {  // hard constraint (_apply_constraint)
      for (int d = 0; d < D; ++d) x[node * D + d] = x0[k * D + d];
      a[node] = a_known;
      continue;
    }
    float zz[4];
    if (z != nullptr) {
      for (int d = 0; d < D; ++d) zz[d] = z[t * D + d];
    } else {
      Philox4 p = draw_block(seed, sample_offset + (uint64_t)smp, (uint32_t)k, step_id,
                             B200MD_DRAW_REPAINT_NORMAL);
      normal4(p, zz);
    }
    for (int d = 0; d < D; ++d)
      x[node * D + d] = wrap_unit(__fadd_rn(x0[k * D + d], __fmul_rn(sigma_i, zz[d])));
    // a_t = argmax_c log(Qbar[a0][c]) - log(-log(u_c))
    float best = -CUDART_INF_F;
    int best_c = 0;
    const int nblk = (C + 3) >> 2;
    for (int b = 0; b < nblk; ++b) {
      float uu[4];
      if (u != nullptr) {
        for (int i = 0; i < 4 && b * 4 + i < C; ++i) uu[i] = u[t * C + b * 4 + i];
      } else {
        Philox4 p = draw_block(seed, sample_offset + (uint64_t)smp, (uint32_t)(k * nblk + b), step_id,
                               B200MD_DRAW_REPAINT_UNIFORM);
        uu[0] = u01(p.x); uu[1] = u01(p.y); uu[2] = u01(p.z); uu[3] = u01(p.w);
      }
      for (int i = 0; i < 4 && b * 4 + i < C; ++i) {
        int c = b * 4 + i;
        float v = logf(qbar[a_known * C + c]) + (-logf(-logf(uu[i])));
        if (c == 0 || v > best) { best = v; best_c = c; }
      }
    }
    a[node] = best_c;
  }
}

// -------------------------------------------------------------------------------------------------
// Wrapped-Gaussian sigma-normalised score.   score/wrapped_gaussian_score.py:131-198,255-419
// -------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T wrapped_gaussian_score_value(T u, T sigma, int32_t kmax) {
  const T pi = (T)3.14159265358979323846;
  // SIGMA_THRESHOLD = torch.Tensor([1/sqrt(2 pi)]) is a float32 tensor (:37); comparing an f64 sigma with
  // it promotes the float32 value, so the threshold is the float32-rounded number in both precisions.
  const T sigma_threshold = (T)(float)0.3989422804014327;
  if (sigma <= sigma_threshold) {
    // small sigma: sigma*S = (-u - sum_k k e_k / sum_k e_k) / sigma            (:301-367)
    const bool small_u = u < (T)0.5;
    const T s2 = sigma * sigma;
    T num = 0, den = 0;
    for (int k = -kmax; k <= kmax; ++k) {
      const T kk = (T)k;
      T arg;
      if (small_u) arg = (T)(-0.5) * (kk * kk + (T)2.0 * u * kk) / s2;                       // (:271-273)
      else arg = (T)(-0.5) * ((kk * kk - (T)1.0) + (T)2.0 * u * (kk + (T)1.0)) / s2;         // (:294-297)
      const T e = exp(arg);
      num += kk * e;
      den += e;
    }
    return (-u - num / den) / sigma;
  }
  // large sigma, Ewald form (:369-419)
  T z2a = 0, z2b = 0, d1 = 0, d2 = 0;
  for (int k = -kmax; k <= kmax; ++k) {
    const T g = (T)k;
    const T upk = u + g;
    const T gu = u * g;
    const T sg = sigma * g;
    const T exp_upk = exp(-pi * upk * upk);
    const T exp_sg = exp((T)(-2.0) * pi * pi * sg * sg);
    const T exp_g = exp(-pi * g * g);
    const T comb = sqrt((T)2.0 * pi) * sigma * exp_sg - exp_g;
    T sn, cs;
    sincos((T)2.0 * pi * gu, &sn, &cs);
    z2a += exp_upk;
    z2b += comb * cs;
    d1 += upk * exp_upk;
    d2 += g * comb * sn;
  }
  const T z2 = z2a + z2b;
  const T deriv = (T)(-2.0) * pi * (d1 + d2);
  return sigma * deriv / z2;
}

// fp32 specialisation of the same two branches, restructured for throughput (the generic version spends ~400
// instructions per element on accurate libm calls and per-term divisions):
//  * small sigma: c = -1 / (2 sigma^2) once, arg_k = c (k^2 + 2 u k) by FMA, exp through the SFU (ex2.approx);
//  * large sigma: the u-independent factors of the reciprocal-space sum are built from ONE exponential,
//    exp(-2 pi^2 sigma^2)^(k^2) by repeated multiplication, exp(-pi k^2) is a constant table, cos / sin(2 pi k u) come
//    from one sincospi and the angle-addition recurrence, and the +-k terms are paired.
// Differences from the generic evaluation are a few 1e-7 relative (summation order), far inside the 5e-5 parity bound
// of the fp32 golden vectors.
__device__ __forceinline__ float wrapped_gaussian_score_value_f32(float u, float sigma, int32_t kmax) {
  const float pi = 3.14159265358979323846f;
  if (sigma <= 0.3989422804014327f) {
    const bool small_u = u < 0.5f;
    const float c = -0.5f / (sigma * sigma);
    const float two_u = 2.0f * u;
    float num = 0.f, den = 0.f;
    for (int k = -kmax; k <= kmax; ++k) {
      const float kk = (float)k;
      // u < 1/2: k^2 + 2 u k (:271-273); else (k^2 - 1) + 2 u (k + 1) (:294-297)
      const float q = small_u ? fmaf(two_u, kk, kk * kk) : fmaf(two_u, kk + 1.0f, kk * kk - 1.0f);
      const float e = __expf(c * q);
      num = fmaf(kk, e, num);
      den += e;
    }
    return (-u - num / den) / sigma;
  }
  // large sigma, Ewald form
  float sn1, cs1;
  sincospif(2.0f * u, &sn1, &cs1);
  const float e1 = __expf(-2.0f * pi * pi * sigma * sigma);      // exp(-2 pi^2 sigma^2 k^2) = e1^(k^2)
  const float amp = 2.5066282746310002f * sigma;                 // sqrt(2 pi) sigma
  const float e0 = __expf(-pi * u * u);
  float z2 = e0 + (amp - 1.0f);                                  // k = 0: exp(-pi u^2) + comb_0
  float d = u * e0;                                              // k = 0 term of d1 (d2 vanishes at k = 0)
  float cs = 1.0f, sn = 0.0f, epow = 1.0f, estep = e1;           // e1^(k^2): multiply by e1^(2k-1) each step
  const float e1sq = e1 * e1;
  for (int k = 1; k <= kmax; ++k) {
    const float kk = (float)k;
    const float cn = cs * cs1 - sn * sn1;
    sn = sn * cs1 + cs * sn1;
    cs = cn;
    epow *= estep;
    estep *= e1sq;
    const float comb = amp * epow - __expf(-pi * kk * kk);
    const float up = u + kk, um = u - kk;
    const float ep = __expf(-pi * up * up), em = __expf(-pi * um * um);
    z2 += (ep + em) + 2.0f * comb * cs;
    d += (up * ep + um * em) + 2.0f * kk * comb * sn;
  }
  return sigma * (-2.0f * pi * d) / z2;
}

template <typename T>
__device__ __forceinline__ T wrapped_gaussian_score_dispatch(T u, T sigma, int32_t kmax) {
  return wrapped_gaussian_score_value<T>(u, sigma, kmax);
}
template <>
__device__ __forceinline__ float wrapped_gaussian_score_dispatch<float>(float u, float sigma, int32_t kmax) {
  return wrapped_gaussian_score_value_f32(u, sigma, kmax);
}

template <typename T>
__global__ void wrapped_gaussian_score_kernel(const T* __restrict__ u_in, const T* __restrict__ s_in,
                                              int64_t n, int32_t kmax, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = wrapped_gaussian_score_dispatch<T>(u_in[i], s_in[i], kmax);
}

// AnalyticalScoreNetwork without permutation symmetry (models/score_networks/analytical_score_network.py:134-298):
//   u = wrap(x - x0);  sigma_eff = sqrt(sigma_d^2 + sigma_t^2);  out = sigma_t * (S(u, sigma_eff) / sigma_eff)
__global__ void analytical_score_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                                        const float* __restrict__ sigma_t, float sigma_d_square,
                                        int64_t n_samples, int32_t per_sample, int32_t kmax,
                                        float* __restrict__ out) {
  const int64_t total = n_samples * per_sample;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t smp = i / per_sample;
    const int32_t k = (int32_t)(i - smp * per_sample);
    const float st = sigma_t[smp];
    const float u = wrap_unit(__fsub_rn(x[i], x0[k]));
    const float se = sqrtf(__fadd_rn(sigma_d_square, __fmul_rn(st, st)));
    const float s = wrapped_gaussian_score_value_f32(u, se, kmax);
    out[i] = __fmul_rn(st, __fdiv_rn(s, se));
  }
}

// -------------------------------------------------------------------------------------------------
// unsorted_segment_sum / mean   models/egnn_utils.py:11-70   (atomic scatter-add; ids need not be sorted)
// -------------------------------------------------------------------------------------------------
__global__ void segment_scatter_kernel(const float* __restrict__ data, const int64_t* __restrict__ ids,
                                       int64_t n_elements, int32_t F, float* __restrict__ out,
                                       float* __restrict__ counts) {
  const int64_t total = n_elements * F;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = t / F;
    const int32_t f = (int32_t)(t - e * F);
    const int64_t seg = ids[e];
    atomicAdd(out + seg * F + f, data[t]);
    if (counts != nullptr && f == 0) atomicAdd(counts + seg, 1.0f);
  }
}
__global__ void segment_mean_finish_kernel(float* __restrict__ out, const float* __restrict__ counts,
                                           int64_t num_segments, int32_t F) {
  const int64_t total = num_segments * F;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x)
    out[t] = out[t] / fmaxf(counts[t / F], 1.0f);
}

static inline unsigned grid_for(int64_t total, int threads) {
  int64_t b = ceil_div64(total, threads);
  if (b > 148 * 32) b = 148 * 32;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int pc_coordinates_update_impl(const float* x, const float* s, const float* z, float score_weight, float sigma,
                               float noise_weight, int64_t n_samples, int32_t n_atoms, int32_t D, uint64_t seed,
                               uint32_t step_id, uint64_t sample_offset, float* x_out, const PcStepScalars* cur,
                               int phase, cudaStream_t st) {
  if (n_samples == 0) return B200MD_OK;
  if (!x || !s || !x_out || n_samples < 0 || n_atoms <= 0 || D < 1 || D > 3)
    return B200MD_ERR_INVALID_ARGUMENT;
  const int64_t total = n_samples * n_atoms;
  const unsigned g = grid_for(total, 256);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(z) |
                         reinterpret_cast<uintptr_t>(x_out)) & 15u) == 0;
  if (D == 3 && (total & 3) == 0 && aligned) {
    ProfileScope ps__(B200MD_KERNEL_UPDATE, st);
    pc_coordinates_update_v4_kernel<<<grid_for(total / 4, 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(s), reinterpret_cast<const float4*>(z),
        score_weight, sigma, noise_weight, total / 4, n_atoms, seed, step_id, sample_offset, reinterpret_cast<float4*>(x_out),
        cur, phase);
  } else {
    ProfileScope ps__(B200MD_KERNEL_UPDATE, st);
    switch (D) {
      case 1: pc_coordinates_update_kernel<1><<<g, 256, 0, st>>>(x, s, z, score_weight, sigma, noise_weight, n_samples, n_atoms, seed, step_id, sample_offset, x_out, cur, phase); break;
      case 2: pc_coordinates_update_kernel<2><<<g, 256, 0, st>>>(x, s, z, score_weight, sigma, noise_weight, n_samples, n_atoms, seed, step_id, sample_offset, x_out, cur, phase); break;
      default: pc_coordinates_update_kernel<3><<<g, 256, 0, st>>>(x, s, z, score_weight, sigma, noise_weight, n_samples, n_atoms, seed, step_id, sample_offset, x_out, cur, phase); break;
    }
  }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int atom_types_update_impl(const float* logits, const int64_t* atom_types, const float* q, const float* q_bar,
                           const float* q_bar_tm1, const float* gumbel, const float* binary, int32_t flags,
                           float small_epsilon, int64_t n_samples, int32_t n_atoms, int32_t n_classes, uint64_t seed,
                           uint32_t step_id, uint64_t sample_offset, int64_t* atom_types_out, float* probs_out,
                           int32_t* status, const PcStepScalars* cur, int phase, cudaStream_t st) {
  if (!logits || !atom_types || !q || !q_bar || !q_bar_tm1 || !atom_types_out || n_samples < 0 ||
      n_atoms <= 0 || n_classes < 2)
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_classes > kMaxClasses) return B200MD_ERR_UNSUPPORTED;
  if (n_samples == 0) return B200MD_OK;
  int threads = ((n_atoms + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  size_t smem = (size_t)3 * n_classes * n_classes * sizeof(float) + (size_t)n_atoms * 8;
  if (smem > 200 * 1024) return B200MD_ERR_UNSUPPORTED;
  auto kernel = atom_types_update_kernel<0>;
  if (n_classes == 2) kernel = atom_types_update_kernel<2>;        // Si (+ MASK)
  else if (n_classes == 3) kernel = atom_types_update_kernel<3>;   // SiGe (+ MASK)
  else if (n_classes == 4) kernel = atom_types_update_kernel<4>;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  { ProfileScope ps__(B200MD_KERNEL_UPDATE, st); kernel<<<(unsigned)n_samples, threads, smem, st>>>(
      logits, atom_types, q, q_bar, q_bar_tm1, gumbel, binary, flags, small_epsilon, n_atoms, n_classes,
      seed, step_id, sample_offset, atom_types_out, probs_out, status, cur, phase); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int repaint_impl(float* x, int64_t* atom_types, const float* x0_known, const int64_t* a0_known,
                 const int64_t* constrained_indices, int32_t n_constrained, float sigma_i, const float* q_bar,
                 const float* z, const float* u, int64_t n_samples, int32_t n_atoms, int32_t D, int32_t n_classes,
                 uint64_t seed, uint32_t step_id, uint64_t sample_offset, const PcStepScalars* cur, cudaStream_t st) {
  if (!x || !atom_types || !x0_known || !a0_known || n_constrained < 0 || n_samples < 0 || D < 1 ||
      D > 3 || ((cur != nullptr || sigma_i >= 0.f) && !q_bar))
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_samples == 0 || n_constrained == 0) return B200MD_OK;
  const int64_t total = n_samples * n_constrained;
  { ProfileScope ps__(B200MD_KERNEL_UPDATE, st); repaint_kernel<<<grid_for(total, 128), 128, 0, st>>>(
      x, atom_types, x0_known, a0_known, constrained_indices, n_constrained, sigma_i, q_bar, z, u,
      n_samples, n_atoms, D, n_classes, seed, step_id, sample_offset, cur); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

}  // namespace b200md

using namespace b200md;

extern "C" {

int b200md_draw_normal(float* out, int64_t n_samples, int32_t units, int32_t unit_size, uint64_t seed,
                       uint32_t step_id, uint32_t draw_id, uint64_t sample_offset,
                       b200md_stream_t stream) {
  return launch_draw<kNormal>(out, n_samples, units, unit_size, 0.f, seed, step_id, draw_id,
                              sample_offset, stream);
}
int b200md_draw_uniform(float* out, int64_t n_samples, int32_t units, int32_t unit_size, uint64_t seed,
                        uint32_t step_id, uint32_t draw_id, uint64_t sample_offset,
                        b200md_stream_t stream) {
  return launch_draw<kUniform>(out, n_samples, units, unit_size, 0.f, seed, step_id, draw_id,
                               sample_offset, stream);
}
int b200md_draw_gumbel(float* out, int64_t n_samples, int32_t units, int32_t unit_size,
                       float small_epsilon, uint64_t seed, uint32_t step_id, uint32_t draw_id,
                       uint64_t sample_offset, b200md_stream_t stream) {
  return launch_draw<kGumbel>(out, n_samples, units, unit_size, small_epsilon, seed, step_id, draw_id,
                              sample_offset, stream);
}

int b200md_wrap_to_unit_cell(const float* x, float* out, int64_t n, b200md_stream_t stream) {
  if (n < 0 || (n > 0 && (!x || !out))) return B200MD_ERR_INVALID_ARGUMENT;
  if (n == 0) return B200MD_OK;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    ProfileScope ps__(B200MD_KERNEL_OTHER, as_stream(stream));
    wrap_kernel_v4<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), n / 4);
  } else {
    ProfileScope ps__(B200MD_KERNEL_OTHER, as_stream(stream));
    wrap_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, out, n);
  }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_pc_coordinates_update(const float* x, const float* s, const float* z, float score_weight,
                                 float sigma, float noise_weight, int64_t n_samples, int32_t n_atoms,
                                 int32_t D, uint64_t seed, uint32_t step_id, uint64_t sample_offset,
                                 float* x_out, b200md_stream_t stream) {
  return pc_coordinates_update_impl(x, s, z, score_weight, sigma, noise_weight, n_samples, n_atoms, D, seed, step_id,
                                    sample_offset, x_out, nullptr, 0, as_stream(stream));
}

int b200md_adaptive_corrector_step_size(const float* s, const float* z, int64_t n_samples, int32_t units, int32_t d,
                                        float sigma, float corrector_r, float small_epsilon, float* partial,
                                        float* step_out, b200md_stream_t stream) {
  if (!s || !z || !partial || !step_out || n_samples <= 0 || units <= 0 || d <= 0 || !(sigma > 0.f))
    return B200MD_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  const int threads = units >= 256 ? 256 : (units >= 64 ? 128 : 32);
  { ProfileScope ps__(B200MD_KERNEL_UPDATE, st); adaptive_norms_kernel<<<(unsigned)n_samples, threads, 0, st>>>(s, z, units, d, partial, n_samples); }
  B200MD_CHECK_LAUNCH();
  { ProfileScope ps__(B200MD_KERNEL_UPDATE, st); adaptive_step_size_kernel<<<1, 1024, 0, st>>>(partial, n_samples, units, sigma, corrector_r, small_epsilon, step_out); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_pc_update_device_weights(const float* x, const float* s, const float* z, const float* weights, float sigma,
                                    int64_t n_values, int32_t wrap, float* out, b200md_stream_t stream) {
  if (!x || !s || !z || !weights || !out || n_values < 0 || !(sigma > 0.f)) return B200MD_ERR_INVALID_ARGUMENT;
  if (n_values == 0) return B200MD_OK;
  cudaStream_t st = as_stream(stream);
  { ProfileScope ps__(B200MD_KERNEL_UPDATE, st); pc_update_device_weights_kernel<<<grid_for(n_values, 256), 256, 0, st>>>(x, s, z, weights, sigma, n_values, wrap, out); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_pc_lattice_update(const float* l, const float* s, const float* z, float score_weight,
                             float sigma_n, float noise_weight, int64_t n_samples, int32_t n_lp,
                             uint64_t seed, uint32_t step_id, uint64_t sample_offset, float* l_out,
                             b200md_stream_t stream) {
  if (!l || !s || !l_out || n_samples < 0 || n_lp <= 0) return B200MD_ERR_INVALID_ARGUMENT;
  if (n_samples == 0) return B200MD_OK;
  const int64_t total = n_samples * ((n_lp + 3) / 4);
  { ProfileScope ps__(B200MD_KERNEL_UPDATE, as_stream(stream)); pc_lattice_update_kernel<<<grid_for(total, 128), 128, 0, as_stream(stream)>>>(
      l, s, z, score_weight, sigma_n, noise_weight, n_samples, n_lp, seed, step_id, sample_offset, l_out); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_atom_types_update(const float* logits, const int64_t* atom_types, const float* q,
                             const float* q_bar, const float* q_bar_tm1, const float* gumbel,
                             const float* binary, int32_t flags, float small_epsilon,
                             int64_t n_samples, int32_t n_atoms, int32_t n_classes, uint64_t seed,
                             uint32_t step_id, uint64_t sample_offset, int64_t* atom_types_out,
                             float* probs_out, int32_t* status, b200md_stream_t stream) {
  return atom_types_update_impl(logits, atom_types, q, q_bar, q_bar_tm1, gumbel, binary, flags, small_epsilon, n_samples,
                                n_atoms, n_classes, seed, step_id, sample_offset, atom_types_out, probs_out, status,
                                nullptr, 0, as_stream(stream));
}

int b200md_repaint(float* x, int64_t* atom_types, const float* x0_known, const int64_t* a0_known,
                   const int64_t* constrained_indices, int32_t n_constrained, float sigma_i,
                   const float* q_bar_i, const float* z, const float* u, int64_t n_samples,
                   int32_t n_atoms, int32_t D, int32_t n_classes, uint64_t seed, uint32_t step_id,
                   uint64_t sample_offset, b200md_stream_t stream) {
  return repaint_impl(x, atom_types, x0_known, a0_known, constrained_indices, n_constrained, sigma_i, q_bar_i, z, u,
                      n_samples, n_atoms, D, n_classes, seed, step_id, sample_offset, nullptr, as_stream(stream));
}

int b200md_wrapped_gaussian_score_f32(const float* u, const float* sigma, int64_t n, int32_t kmax,
                                      float* out, b200md_stream_t stream) {
  if (n < 0 || kmax < 0 || (n > 0 && (!u || !sigma || !out))) return B200MD_ERR_INVALID_ARGUMENT;
  if (n == 0) return B200MD_OK;
  { ProfileScope ps__(B200MD_KERNEL_OTHER, as_stream(stream)); wrapped_gaussian_score_kernel<float><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(u, sigma, n, kmax, out); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}
int b200md_wrapped_gaussian_score_f64(const double* u, const double* sigma, int64_t n, int32_t kmax,
                                      double* out, b200md_stream_t stream) {
  if (n < 0 || kmax < 0 || (n > 0 && (!u || !sigma || !out))) return B200MD_ERR_INVALID_ARGUMENT;
  if (n == 0) return B200MD_OK;
  { ProfileScope ps__(B200MD_KERNEL_OTHER, as_stream(stream)); wrapped_gaussian_score_kernel<double><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(u, sigma, n, kmax, out); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_analytical_score_f32(const float* x, const float* x0, const float* sigma_t, float sigma_d,
                                int64_t n_samples, int32_t n_atoms, int32_t spatial_dimension, int32_t kmax,
                                float* out, b200md_stream_t stream) {
  if (n_samples < 0 || n_atoms <= 0 || spatial_dimension <= 0 || kmax < 0 || !(sigma_d > 0.f) ||
      (n_samples > 0 && (!x || !x0 || !sigma_t || !out)))
    return B200MD_ERR_INVALID_ARGUMENT;
  if (n_samples == 0) return B200MD_OK;
  const int32_t per_sample = n_atoms * spatial_dimension;
  { ProfileScope ps__(B200MD_KERNEL_OTHER, as_stream(stream)); analytical_score_kernel<<<grid_for(n_samples * per_sample, 256), 256, 0, as_stream(stream)>>>(
        x, x0, sigma_t, sigma_d * sigma_d, n_samples, per_sample, kmax, out); }
  B200MD_CHECK_LAUNCH();
  return B200MD_OK;
}

int b200md_segment_reduce(const float* data, const int64_t* segment_ids, int64_t n_elements,
                          int32_t n_features, int64_t num_segments, int32_t mean, float* out,
                          float* scratch_counts, b200md_stream_t stream) {
  if (n_elements < 0 || n_features <= 0 || num_segments < 0 || !out || (mean && !scratch_counts))
    return B200MD_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  if (num_segments == 0) return B200MD_OK;
  cudaMemsetAsync(out, 0, sizeof(float) * num_segments * n_features, st);
  if (mean) cudaMemsetAsync(scratch_counts, 0, sizeof(float) * num_segments, st);
  if (n_elements > 0) {
    if (!data || !segment_ids) return B200MD_ERR_INVALID_ARGUMENT;
    { ProfileScope ps__(B200MD_KERNEL_OTHER, st); segment_scatter_kernel<<<grid_for(n_elements * n_features, 256), 256, 0, st>>>(
        data, segment_ids, n_elements, n_features, out, mean ? scratch_counts : nullptr); }
    B200MD_CHECK_LAUNCH();
  }
  if (mean) {
    { ProfileScope ps__(B200MD_KERNEL_OTHER, st); segment_mean_finish_kernel<<<grid_for(num_segments * n_features, 256), 256, 0, st>>>(
        out, scratch_counts, num_segments, n_features); }
    B200MD_CHECK_LAUNCH();
  }
  return B200MD_OK;
}

}  // extern "C"
