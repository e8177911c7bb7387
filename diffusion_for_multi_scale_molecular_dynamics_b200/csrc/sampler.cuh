// Per-step scalars of the predictor-corrector loop, resident in device memory so that ONE captured CUDA graph of a whole
// PC step (b200md_pc_sample, sampler.cu) can be replayed for every time index: the kernels of the step read their
// step-dependent arguments from `current` instead of from launch parameters, and the last node of the graph copies the
// next row of the table into `current`.
// Reference loop the table unrolls: generators/predictor_corrector_axl_generator.py:112-161,
// generators/langevin_generator.py:536-645 (predictor), :693-805 (corrector).
#pragma once
#include "common.cuh"

namespace b200md {

struct PcStepScalars {
  float sigma_pred, g2, g;             // predictor: sigma_i of the score network, score weight g_i^2, noise weight g_i
  float sigma_corr, eps, sqrt_2eps;    // correctors: sigma_i (sigma_min at i = 0), eps_i, sqrt(2 eps_i)
  float repaint_sigma;                 // RePaint noising after the predictor: sigma_{i-1}, < 0 = hard constraint (i = 0)
  uint32_t step_id;                    // counter key of the predictor's draws; corrector k uses step_id + 1 + k
  int32_t atom_flags_pred, atom_flags_corr;
  int32_t q_index_pred, q_index_corr, q_index_repaint;   // rows of the [T][C][C] transition tables
  int32_t pad[3];
};
static_assert(sizeof(PcStepScalars) == 64, "one 64-byte row per step");

// `phase`: 0 = predictor, k + 1 = k-th corrector.  `cur` == nullptr: the scalar arguments are used as given.
int pc_coordinates_update_impl(const float* x, const float* s, const float* z, float score_weight, float sigma,
                               float noise_weight, int64_t n_samples, int32_t n_atoms, int32_t D, uint64_t seed,
                               uint32_t step_id, uint64_t sample_offset, float* x_out, const PcStepScalars* cur,
                               int phase, cudaStream_t st);
// with `cur`, q / q_bar / q_bar_tm1 are the BASES of the [T][C][C] tables
int atom_types_update_impl(const float* logits, const int64_t* atom_types, const float* q, const float* q_bar,
                           const float* q_bar_tm1, const float* gumbel, const float* binary, int32_t flags,
                           float small_epsilon, int64_t n_samples, int32_t n_atoms, int32_t n_classes, uint64_t seed,
                           uint32_t step_id, uint64_t sample_offset, int64_t* atom_types_out, float* probs_out,
                           int32_t* status, const PcStepScalars* cur, int phase, cudaStream_t st);
// with `cur`, q_bar is the BASE of the [T][C][C] table
int repaint_impl(float* x, int64_t* atom_types, const float* x0_known, const int64_t* a0_known,
                 const int64_t* constrained_indices, int32_t n_constrained, float sigma_i, const float* q_bar,
                 const float* z, const float* u, int64_t n_samples, int32_t n_atoms, int32_t D, int32_t n_classes,
                 uint64_t seed, uint32_t step_id, uint64_t sample_offset, const PcStepScalars* cur, cudaStream_t st);

}  // namespace b200md
