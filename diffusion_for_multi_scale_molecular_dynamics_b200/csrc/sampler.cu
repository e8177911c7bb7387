// Whole predictor-corrector trajectory enqueued from one host call.
// Reference control flow: generators/predictor_corrector_axl_generator.py:112-161 (time loop),
// generators/langevin_generator.py:536-645 (predictor_step), :693-805 (corrector_step),
// generators/constrained_langevin_generator.py:94-163 (RePaint after each predictor step).
#include "egnn.cuh"

using namespace b200md;

extern "C" {

int64_t b200md_pc_sample_workspace_bytes(const b200md_egnn_config* cfg, int64_t B, int32_t N, int64_t cap) {
  if (!cfg || B < 0 || N <= 0 || cap < 0) return B200MD_ERR_INVALID_ARGUMENT;
  EgnnWorkspace ws;
  int64_t base = carve_workspace(*cfg, B, N, cap, nullptr, ws);
  const int64_t C = cfg->num_atom_types + 1;
  int64_t logits = (B * N * C * 4 + 255) & ~int64_t(255);
  int64_t score = (B * N * 3 * 4 + 255) & ~int64_t(255);
  return base + logits + score;
}

int b200md_pc_sample(const b200md_egnn_config* cfg, const float* packed, const b200md_pc_schedule* s,
                     int64_t* atom_types, float* x, const float* lattice, int64_t B, int32_t N,
                     uint64_t seed, uint64_t sample_offset, void* workspace, int64_t workspace_bytes,
                     int64_t edge_capacity, int32_t* status, int32_t impl, b200md_stream_t stream) {
  if (!cfg || !packed || !s || !atom_types || !x || !lattice || !workspace || !status || B < 0 || N <= 0)
    return B200MD_ERR_INVALID_ARGUMENT;
  if (!s->sigma_host || !s->g_host || !s->g2_host || !s->epsilon_host || !s->sqrt_2_epsilon_host || !s->q ||
      !s->q_bar || !s->q_bar_tm1)
    return B200MD_ERR_INVALID_ARGUMENT;
  const int T = s->total_time_steps;
  if (T < 2 || s->start_step_index > T || s->start_step_index <= s->end_step_index || s->start_step_index <= 0 ||
      s->end_step_index < 0 || s->number_of_corrector_steps < 0)
    return B200MD_ERR_INVALID_ARGUMENT;
  if (B == 0) return B200MD_OK;
  EgnnLayout lay;
  int rc = build_layout(*cfg, lay);
  if (rc != B200MD_OK) return rc;
  EgnnWorkspace ws;
  const int64_t base = carve_workspace(*cfg, B, N, edge_capacity, workspace, ws);
  const int C = cfg->num_atom_types + 1;
  const int64_t logits_bytes = (B * N * C * 4 + 255) & ~int64_t(255);
  const int64_t score_bytes = (B * N * 3 * 4 + 255) & ~int64_t(255);
  if (base + logits_bytes + score_bytes > workspace_bytes) return B200MD_ERR_WORKSPACE_TOO_SMALL;
  float* logits = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + base);
  float* score = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + base + logits_bytes);
  cudaStream_t st = as_stream(stream);
  const int ncorr = s->number_of_corrector_steps;
  const uint32_t phases = (uint32_t)(1 + ncorr);
  const int64_t CC = (int64_t)C * C;

  for (int i = s->start_step_index - 1; i >= s->end_step_index; --i) {
    // ---- predictor step with index_i = i + 1, idx = i ----
    {
      const int idx = i;
      const float sigma_i = s->sigma_host[idx];
      const uint32_t step_id = (uint32_t)i * phases;
      rc = egnn_forward_impl(*cfg, lay, packed, atom_types, x, lattice, nullptr, sigma_i, B, N, 1, logits, score, ws,
                             edge_capacity, status, impl, st);
      if (rc != B200MD_OK) return rc;
      const bool last = (idx == 0);
      int flags = s->atom_flags & B200MD_ATOM_GREEDY;
      if ((s->atom_flags & B200MD_ATOM_ONE_TRANSITION) && !last) flags |= B200MD_ATOM_ONE_TRANSITION;
      if (last) flags |= B200MD_ATOM_LAST_STEP;
      rc = b200md_atom_types_update(logits, atom_types, s->q + idx * CC, s->q_bar + idx * CC, s->q_bar_tm1 + idx * CC,
                                    nullptr, nullptr, flags, s->small_epsilon, B, N, C, seed, step_id, sample_offset,
                                    atom_types, nullptr, status, stream);
      if (rc != B200MD_OK) return rc;
      rc = b200md_pc_coordinates_update(x, score, nullptr, s->g2_host[idx], sigma_i, s->g_host[idx], B, N, 3, seed,
                                        step_id, sample_offset, x, stream);
      if (rc != B200MD_OK) return rc;
      if (s->n_constrained > 0) {
        // _repaint_composition(index_i - 1 = i): noise the known atoms to time index i (none at i == 0)
        const float sig = (i == 0) ? -1.0f : s->sigma_host[i - 1];
        const float* qb = (i == 0) ? nullptr : s->q_bar + (int64_t)(i - 1) * CC;
        rc = b200md_repaint(x, atom_types, s->x0_known, s->a0_known, s->constrained_indices, s->n_constrained, sig,
                            qb, nullptr, nullptr, B, N, 3, C, seed, step_id, sample_offset, stream);
        if (rc != B200MD_OK) return rc;
      }
    }
    // ---- corrector steps with index_i = i ----
    for (int k = 0; k < ncorr; ++k) {
      const int idx = (i == 0) ? 0 : i - 1;
      const float sigma_i = (i == 0) ? s->sigma_min : s->sigma_host[idx];
      const uint32_t step_id = (uint32_t)i * phases + 1u + (uint32_t)k;
      rc = egnn_forward_impl(*cfg, lay, packed, atom_types, x, lattice, nullptr, sigma_i, B, N, 1, logits, score, ws,
                             edge_capacity, status, impl, st);
      if (rc != B200MD_OK) return rc;
      rc = b200md_pc_coordinates_update(x, score, nullptr, s->epsilon_host[i], sigma_i, s->sqrt_2_epsilon_host[i], B, N,
                                        3, seed, step_id, sample_offset, x, stream);
      if (rc != B200MD_OK) return rc;
      if (s->atom_type_transition_in_corrector) {
        rc = b200md_atom_types_update(logits, atom_types, s->q + idx * CC, s->q_bar + idx * CC,
                                      s->q_bar_tm1 + idx * CC, nullptr, nullptr, s->atom_flags, s->small_epsilon, B, N,
                                      C, seed, step_id, sample_offset, atom_types, nullptr, status, stream);
        if (rc != B200MD_OK) return rc;
      }
    }
  }
  return B200MD_OK;
}

}  // extern "C"
