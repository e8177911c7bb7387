/*
 * b200md.h -- C ABI of the B200-native predictor-corrector sampling hot path.
 *
 * This is the drop-in boundary for ONE path of mila-iqia/diffusion_for_multi_scale_molecular_dynamics:
 * LangevinGenerator (PC sampler) -> EGNNScoreNetwork -> periodic radius graph / EGNN message passing /
 * wrapped-Gaussian score / fused predictor-corrector update.  The reference has no FFI layer of its own
 * (it is pure Python/PyTorch); every entry point below therefore cites the reference *Python* interface
 * it replaces (paths relative to src/diffusion_for_multi_scale_molecular_dynamics/ in the reference).
 *
 * Conventions
 *  - plain C, no torch types; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - memory is owned by the caller (PyTorch); the library never allocates device memory;
 *  - every launch goes on the `stream` argument (a cudaStream_t passed as void*);
 *  - return value: 0 on success, negative b200md_error on failure (see b200md_error_string);
 *  - data-dependent failures (the reference's `assert tensor.all()` host syncs) are reported through a
 *    device-side int32 `status` word (bitwise OR of B200MD_STATUS_*) that the host wrapper reads once per
 *    sample() instead of once per step;
 *  - floating point is fp32, atom types / edge indices at the boundary are int64 (as in the reference).
 */
#ifndef B200MD_H_
#define B200MD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MD_ABI_VERSION 1

typedef void* b200md_stream_t; /* cudaStream_t */

typedef enum {
  B200MD_OK = 0,
  B200MD_ERR_INVALID_ARGUMENT = -1,
  B200MD_ERR_CUDA = -2,
  B200MD_ERR_UNSUPPORTED = -3,
  B200MD_ERR_WORKSPACE_TOO_SMALL = -4
} b200md_error;

/* bits of the device-side status word */
#define B200MD_STATUS_X_OUT_OF_RANGE 1u     /* models/score_networks/score_network.py:110-114 */
#define B200MD_STATUS_A_OUT_OF_RANGE 2u     /* models/score_networks/score_network.py:159-165 */
#define B200MD_STATUS_CUTOFF_TOO_LARGE 4u   /* utils/neighbors.py:107-113 */
#define B200MD_STATUS_EDGE_OVERFLOW 8u      /* edge workspace capacity exceeded (no reference analogue) */
#define B200MD_STATUS_MASK_REMAINS 16u      /* generators/langevin_generator.py:616-620 */
#define B200MD_STATUS_BAD_LOGITS 32u        /* utils/d3pm_utils.py:143-145 */
#define B200MD_STATUS_SIGMA_OUT_OF_RANGE 64u
#define B200MD_STATUS_TC_RANGE 128u         /* an activation exceeded the fp16 range of B200MD_IMPL_TCGEN05_F16X3 */

const char* b200md_error_string(int code);
int b200md_abi_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t b200md_launch_count(void);

/* Optional in-stream kernel timing (bench.py's roofline leg): when enabled every launch is bracketed by a pair of
 * CUDA events recorded on the launch stream; b200md_profile_collect synchronises them, adds the elapsed
 * milliseconds and launch counts per kernel class into the caller's HOST arrays (length n_classes >=
 * B200MD_KERNEL_CLASSES) and clears the record. */
#define B200MD_KERNEL_GRAPH 0   /* radius graph: count / scan / fill */
#define B200MD_KERNEL_NODE 1    /* EGNN node kernel: embed, node MLP, read-out */
#define B200MD_KERNEL_EDGE 2    /* EGNN edge kernel: message + coordinate MLPs, segmented reduction */
#define B200MD_KERNEL_UPDATE 3  /* predictor/corrector coordinate, atom-type, lattice, repaint updates */
#define B200MD_KERNEL_OTHER 4   /* noise draws, wrap, wrapped-Gaussian score, segment reductions, packing */
#define B200MD_KERNEL_CLASSES 5
int b200md_profile_enable(int32_t on);
int b200md_profile_collect(double* ms_by_class_host, int64_t* launches_by_class_host, int32_t n_classes);

/* ------------------------------------------------------------------------------------------------
 * Counter-based noise (replaces the CPU torch.rand/randn + H2D copy of
 * generators/langevin_generator.py:92-111 and generators/trajectory_initializer.py:109-119).
 * Value (sample s, unit u, lane l) depends only on (seed, s + sample_offset, u, l, step_id, draw_id):
 * results are independent of batch chunking and of the number of GPUs.
 * `unit_size` = values per unit (3 for coordinates, num_classes for Gumbel, 1 for the greedy binary draw).
 * ---------------------------------------------------------------------------------------------- */
#define B200MD_DRAW_COORD_NORMAL 0u
#define B200MD_DRAW_GUMBEL 1u
#define B200MD_DRAW_BINARY 2u
#define B200MD_DRAW_LATTICE_NORMAL 3u
#define B200MD_DRAW_INIT_UNIFORM 4u
#define B200MD_DRAW_REPAINT_NORMAL 5u
#define B200MD_DRAW_REPAINT_UNIFORM 6u
#define B200MD_DRAW_REPAINT_INIT 7u
#define B200MD_DRAW_LATTICE_NORMAL_2 8u /* the corrector's second lattice draw (langevin_generator.py:771-777) */
#define B200MD_DRAW_INIT_LATTICE 9u     /* the starting lattice of a free-lattice trajectory (trajectory_initializer.py:114-121) */

int b200md_draw_normal(float* out, int64_t n_samples, int32_t units_per_sample, int32_t unit_size,
                       uint64_t seed, uint32_t step_id, uint32_t draw_id, uint64_t sample_offset,
                       b200md_stream_t stream);
int b200md_draw_uniform(float* out, int64_t n_samples, int32_t units_per_sample, int32_t unit_size,
                        uint64_t seed, uint32_t step_id, uint32_t draw_id, uint64_t sample_offset,
                        b200md_stream_t stream);
/* -log(-log(clip(u, small_epsilon)))  (langevin_generator.py:100-107) */
int b200md_draw_gumbel(float* out, int64_t n_samples, int32_t units_per_sample, int32_t unit_size,
                       float small_epsilon, uint64_t seed, uint32_t step_id, uint32_t draw_id,
                       uint64_t sample_offset, b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * map_relative_coordinates_to_unit_cell (utils/basis_transformations.py:95-119):
 * out = remainder(x, 1); out[out == 1] = 0.   `out` may alias `x`.
 * ---------------------------------------------------------------------------------------------- */
int b200md_wrap_to_unit_cell(const float* x, float* out, int64_t n, b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LangevinGenerator._relative_coordinates_update (generators/langevin_generator.py:155-201), used by both
 * predictor_step (:623-629) and corrector_step (:740-758):
 *   x_out = wrap( (x + (score_weight * s) / sigma) + noise_weight * z )
 * z is read from `z_or_null` when given, otherwise drawn in-kernel
 * (draw_id B200MD_DRAW_COORD_NORMAL, unit = atom).  `x_out` may alias `x`.
 * ---------------------------------------------------------------------------------------------- */
int b200md_pc_coordinates_update(const float* x, const float* sigma_normalized_score,
                                 const float* z_or_null, float score_weight, float sigma,
                                 float noise_weight, int64_t n_samples, int32_t n_atoms,
                                 int32_t spatial_dimension, uint64_t seed, uint32_t step_id,
                                 uint64_t sample_offset, float* x_out, b200md_stream_t stream);

/* LangevinGenerator._lattice_parameters_update (langevin_generator.py:441-490):
 *   l_out = (l + (score_weight * s) / sigma_n) + noise_weight * z          (no wrap)            */
int b200md_pc_lattice_update(const float* l, const float* sigma_normalized_score,
                             const float* z_or_null, float score_weight, float sigma_n,
                             float noise_weight, int64_t n_samples, int32_t n_lattice_parameters,
                             uint64_t seed, uint32_t step_id, uint64_t sample_offset, float* l_out,
                             b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * AdaptiveCorrectorGenerator._generic_corrector_step_size (generators/adaptive_corrector.py:98-148):
 *   eps = 2 (r ||z|| / max(||sigma s|| / sigma, small_epsilon))^2
 * ||sigma s|| = mean over samples of the L2 norm of the sample's score block [units, d]; ||z|| = mean over
 * (sample, unit) of the L2 norm of the d-vector (coordinates: units = atoms, d = 3; lattice: units = 1, d = 6).
 * partial: workspace of 2 * n_samples floats.  step_out [2] (device) = (eps, sqrt(2 eps)): the score and noise
 * weights of the corrector update, consumed by b200md_pc_update_device_weights without a host round trip.
 * ---------------------------------------------------------------------------------------------- */
int b200md_adaptive_corrector_step_size(const float* sigma_normalized_score, const float* z,
                                        int64_t n_samples, int32_t units, int32_t d, float sigma,
                                        float corrector_r, float small_epsilon, float* partial,
                                        float* step_out, b200md_stream_t stream);
/* out = [wrap]((x + (weights[0] * s) / sigma) + weights[1] * z), elementwise over n_values; weights on the device. */
int b200md_pc_update_device_weights(const float* x, const float* sigma_normalized_score, const float* z,
                                    const float* weights, float sigma, int64_t n_values, int32_t wrap,
                                    float* out, b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LangevinGenerator._atom_types_update (langevin_generator.py:247-439) fused with
 * get_probability_at_previous_time_step / get_probability_from_logits (utils/d3pm_utils.py:64-150).
 * q / q_bar / q_bar_tm1 are the [C,C] matrices of the current time index (device pointers into the
 * NoiseScheduler tables).  gumbel [B,N,C] and binary [B,N] are injected when non-NULL, else drawn.
 * flags: bit0 greedy sampling, bit1 one transition per step, bit2 last time step (assert no MASK -> status).
 * Optional outputs (may be NULL): probs_out [B,N,C] = p(a_{t-1}|a_t) after the greedy adjustment.
 * ---------------------------------------------------------------------------------------------- */
#define B200MD_ATOM_GREEDY 1
#define B200MD_ATOM_ONE_TRANSITION 2
#define B200MD_ATOM_LAST_STEP 4
int b200md_atom_types_update(const float* logits, const int64_t* atom_types, const float* q,
                             const float* q_bar, const float* q_bar_tm1,
                             const float* gumbel_or_null, const float* binary_or_null,
                             int32_t flags, float small_epsilon, int64_t n_samples,
                             int32_t n_atoms, int32_t n_classes, uint64_t seed, uint32_t step_id,
                             uint64_t sample_offset, int64_t* atom_types_out, float* probs_out,
                             int32_t* status, b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * RePaint overwrite (generators/constrained_langevin_generator.py:74-82,120-163 with
 * data/diffusion/noising_transform.py:98-200, noisers/relative_coordinates_noiser.py:33-67,
 * noisers/atom_types_noiser.py:30-60):
 *   X[:, idx] = wrap(x0_known + sigma_i * z)          (sigma_i < 0  => hard constraint, no noise)
 *   A[:, idx] = argmax(log(onehot(a0_known) Qbar_i) + gumbel(u))
 * In place on X [B,N,d] and A [B,N].  z [B,Nc,d] / u [B,Nc,C] injected when non-NULL.
 * ---------------------------------------------------------------------------------------------- */
int b200md_repaint(float* x, int64_t* atom_types, const float* x0_known, const int64_t* a0_known,
                   const int64_t* constrained_indices, int32_t n_constrained, float sigma_i,
                   const float* q_bar_i, const float* z_or_null, const float* u_or_null,
                   int64_t n_samples, int32_t n_atoms, int32_t spatial_dimension,
                   int32_t n_classes, uint64_t seed, uint32_t step_id, uint64_t sample_offset,
                   b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * get_coordinates_sigma_normalized_score (score/wrapped_gaussian_score.py:131-198): sigma-normalised score
 * of the wrapped Gaussian, lattice-image sum k in [-kmax, kmax], three numerically stable branches
 * (:201-419).  u in [0,1), sigma > 0, both [n].  f32 and f64 variants.
 * ---------------------------------------------------------------------------------------------- */
int b200md_wrapped_gaussian_score_f32(const float* u, const float* sigma, int64_t n, int32_t kmax,
                                      float* out, b200md_stream_t stream);
int b200md_wrapped_gaussian_score_f64(const double* u, const double* sigma, int64_t n,
                                      int32_t kmax, double* out, b200md_stream_t stream);

/* AnalyticalScoreNetwork._forward_unchecked without permutation symmetry
 * (models/score_networks/analytical_score_network.py:134-298): the exact sigma-normalised score of isotropic Gaussian
 * displacements of width sigma_d around equilibrium positions x0 [N,d], noised to sigma_t [B]:
 *   u = wrap(x - x0), sigma_eff = sqrt(sigma_d^2 + sigma_t^2), out = sigma_t * S_wrapped(u, sigma_eff) / sigma_eff.   */
int b200md_analytical_score_f32(const float* x, const float* x0, const float* sigma_t, float sigma_d,
                                int64_t n_samples, int32_t n_atoms, int32_t spatial_dimension, int32_t kmax,
                                float* out, b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * unsorted_segment_sum / unsorted_segment_mean (models/egnn_utils.py:11-70).
 * data [E,F] f32, segment_ids [E] int64 in [0,num_segments); out [num_segments,F].
 * `scratch_counts` [num_segments] f32 is needed for mean.
 * ---------------------------------------------------------------------------------------------- */
int b200md_segment_reduce(const float* data, const int64_t* segment_ids, int64_t n_elements,
                          int32_t n_features, int64_t num_segments, int32_t mean, float* out,
                          float* scratch_counts, b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Periodic radius graph.
 *
 * b200md_radius_graph_count / _fill replace get_periodic_adjacency_information
 * (utils/neighbors.py:36-224), get_adj_matrix (models/graph_utils.py:10-50) and
 * get_edges_with_radial_cutoff (models/egnn_utils.py:107-144).
 *
 * positions [B,N,3]: Cartesian (positions_are_relative = 0) or relative coordinates (1; p = x @ cell is
 * then formed in-kernel, basis_transformations.py:56).  cell [B,3,3], rows are a1,a2,a3.
 * Edge test, in fp32 without FMA contraction (dense semantics of neighbors.py:170,188-190):
 *     shifted = p_j + (n @ cell);  diff = p_i - shifted;  d2 = (dx*dx + dy*dy) + dz*dz;  0 < d2 <= rc*rc
 * over the 27 images n in {-1,0,1}^3 in itertools.product order (utils/lattice_utils.py:24-27).
 *
 * mode B200MD_GRAPH_UNIQUE_PAIRS : one edge per ordered pair (src,dst) having >= 1 image inside the cutoff,
 *      sorted by (batch, src, dst) -- the torch.unique(dim=1) form of egnn_utils.py:139-142.
 *      counts are per (b, src): counts [B*N].
 * mode B200MD_GRAPH_PER_IMAGE : one edge per (image, src, dst), order (batch, image, src, ascending d2,
 *      ties by dst) -- the native AdjacencyInfo order.  counts [B*27*N].
 *
 * count: writes `counts` (int32) and per-structure totals `edges_per_structure` [B] (int64) and sets
 *        B200MD_STATUS_CUTOFF_TOO_LARGE when rc >= the shortest cell-crossing distance
 *        (neighbors.py:248-351).
 * fill : given `offsets` = exclusive scan of `counts` (int64, length = len(counts)), writes
 *        src/dst int64 (local atom indices + node_index_offset*b, so node_index_offset = N gives the
 *        graph-batched indices of neighbors.py:381-390), optional shifts [E,3] f32 (the lattice vector
 *        n @ cell of each edge) and optional edge_batch_indices [E] int64.
 * ---------------------------------------------------------------------------------------------- */
#define B200MD_GRAPH_UNIQUE_PAIRS 0
#define B200MD_GRAPH_PER_IMAGE 1
int b200md_radius_graph_count(const float* positions, const float* cell,
                              int32_t positions_are_relative, float radial_cutoff, int64_t n_samples,
                              int32_t n_atoms, int32_t mode, int32_t* counts,
                              int64_t* edges_per_structure, int32_t* status,
                              b200md_stream_t stream);
int b200md_radius_graph_fill(const float* positions, const float* cell,
                             int32_t positions_are_relative, float radial_cutoff, int64_t n_samples,
                             int32_t n_atoms, int32_t mode, const int64_t* offsets,
                             int64_t node_index_offset, int64_t* src_out, int64_t* dst_out,
                             float* shifts_out_or_null, int64_t* edge_batch_or_null,
                             b200md_stream_t stream);
/* One sweep instead of two for the unique-pair form without shifts (get_edges_with_radial_cutoff,
 * models/egnn_utils.py:107-144): the count pass also writes the hit bit row of every source atom,
 * hit_mask [B*N, ceil(N/32)] uint32 (bit j of row (b, i): j is a neighbour of i), and the fill expands those rows
 * (one warp per row; same output order) without the positions. */
int b200md_radius_graph_count_masked(const float* positions, const float* cell,
                                     int32_t positions_are_relative, float radial_cutoff, int64_t n_samples,
                                     int32_t n_atoms, int32_t* counts, int64_t* edges_per_structure,
                                     int32_t* status, uint32_t* hit_mask, b200md_stream_t stream);
int b200md_radius_graph_fill_from_mask(const uint32_t* hit_mask, int64_t n_samples, int32_t n_atoms,
                                       const int64_t* offsets, int64_t node_index_offset, int64_t* src_out,
                                       int64_t* dst_out, int64_t* edge_batch_or_null, b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pair distances: compute_distances_in_batch (utils/structure_utils.py:41-121), the sample statistic of the
 * reference's structure Kolmogorov-Smirnov metric (models/axl_diffusion_lightning_model.py:531,633).
 *
 * All distances d with 0 < d <= max_distance between atom i and the periodic images (n in {-1,0,1}^3) of
 * atom j, in the reference's order (structure, image, i, ascending d2, ties by j).  d2 follows the rounding
 * chain of the radius graph; the filter is made on d = sqrt(d2) as the reference does (:113-121).
 * count: counts [B*27*N] int32 and distances_per_structure [B] int64.
 * fill : offsets = exclusive scan of counts (int64) -> distances_out [sum(counts)] f32.
 * ---------------------------------------------------------------------------------------------- */
int b200md_pair_distances_count(const float* positions, const float* cell,
                                int32_t positions_are_relative, float max_distance, int64_t n_samples,
                                int32_t n_atoms, int32_t* counts, int64_t* distances_per_structure,
                                b200md_stream_t stream);
int b200md_pair_distances_fill(const float* positions, const float* cell,
                               int32_t positions_are_relative, float max_distance, int64_t n_samples,
                               int32_t n_atoms, const int64_t* offsets, float* distances_out,
                               b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * ForceFieldAugmentedScoreNetwork.get_relative_coordinates_pseudo_force
 * (models/score_networks/force_field_augmented_score_network.py:44-236): repulsion phi(r) = s (r - r0)^2 for r < r0.
 *   cell = diag(clip(lattice[:, :3], min_box_size)); p = x @ cell; for every (i, image of j) with 0 < d2 <= r0^2:
 *   disp = (p_j - p_i) + shift;  F_i += 2 s (|disp| - r0) / (|disp| + 1e-8) disp;  out = [add_to +] F @ inverse(cell)
 * x_rel [B,N,3], lattice [B,6], out [B,N,3] (relative-coordinate pseudo force, optionally added to a score).
 * Sets B200MD_STATUS_CUTOFF_TOO_LARGE like get_periodic_adjacency_information (utils/neighbors.py:107-113).
 * ---------------------------------------------------------------------------------------------- */
int b200md_repulsive_pseudo_force(const float* x_rel, const float* lattice, float min_box_size,
                                  float radial_cutoff, float strength, int64_t n_samples, int32_t n_atoms,
                                  const float* add_to_or_null, float* out, int32_t* status,
                                  b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stillinger-Weber potential energy of every structure of a batch: what the reference obtains from LAMMPS
 * (`pair_style sw`, oracle/lammps_energy_oracle.py:60-193) one structure at a time, for its energy
 * Kolmogorov-Smirnov metric (models/axl_diffusion_lightning_model.py:141,560-640).
 *   E = sum_{i<j} phi2(r_ij) + sum_i sum_{j<k} phi3(r_ij, r_ik, theta_jik)       (LAMMPS pair_sw)
 * x_rel [B,N,3], lattice [B,6] (orthogonal box diag(lattice[:, :3])), atom_types [B,N] int64 in [0, n_types).
 * params [n_types^3][10] f32, entry (i,j,k) = { a*sigma, (a*sigma)^2, sigma, sigma*gamma, lambda*epsilon, costheta0,
 * A*epsilon*B*sigma^p, A*epsilon*sigma^q, p, q }; pair terms use entry (i,j,j).  max_cutoff = largest a*sigma.
 * A box side below min_box_size gives energy 0 (the reference skips such boxes, :121-127); a side not larger
 * than max_cutoff raises B200MD_STATUS_CUTOFF_TOO_LARGE (neighbours beyond the first image shell).
 * energies_out [B] f32 (eV).
 * ---------------------------------------------------------------------------------------------- */
int b200md_stillinger_weber_energy(const float* x_rel, const float* lattice, const int64_t* atom_types,
                                   const float* params, int32_t n_types, float max_cutoff,
                                   float min_box_size, int64_t n_samples, int32_t n_atoms,
                                   float* energies_out, int32_t* status, b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * EGNN score network: EGNNScoreNetwork._forward_unchecked (models/score_networks/egnn_score_network.py:226-303)
 * + EGNN / E_GCL (models/egnn.py:266-288,362-385) + the MASK-logit rule of ScoreNetwork.forward
 * (models/score_networks/score_network.py:183-185,221).
 * ---------------------------------------------------------------------------------------------- */
#define B200MD_EGNN_RESIDUAL 1
#define B200MD_EGNN_ATTENTION 2
#define B200MD_EGNN_NORMALIZE 4
#define B200MD_EGNN_TANH 8
#define B200MD_EGNN_COORDS_MEAN 16
#define B200MD_EGNN_MESSAGE_MEAN 32
#define B200MD_EGNN_RADIAL_CUTOFF 64   /* edges="radial_cutoff" (else fully connected) */
#define B200MD_EGNN_KEEP_DUPLICATES 128 /* drop_duplicate_edges=False */

typedef struct {
  int32_t spatial_dimension;      /* 3 */
  int32_t num_atom_types;         /* excluding MASK */
  int32_t n_layers;
  int32_t n_bloch;                /* number of Bloch wave vectors; uplift dimension is 2*n_bloch */
  int32_t node_hidden;            /* node_hidden_dimensions_size  (H) */
  int32_t node_n_hidden;          /* node_n_hidden_dimensions */
  int32_t message_hidden;         /* message_hidden_dimensions_size (M) */
  int32_t message_n_hidden;
  int32_t coord_hidden;           /* coordinate_hidden_dimensions_size (Cx) */
  int32_t coord_n_hidden;
  int32_t flags;                  /* B200MD_EGNN_* */
  float radial_cutoff;            /* Angstrom; used when B200MD_EGNN_RADIAL_CUTOFF */
  float lattice_clip_min;         /* float32(2.2 * radial_cutoff), egnn_score_network.py:236-239 */
} b200md_egnn_config;

/* Parameters are handed over in state_dict order of the reference module (egnn_score_network.py /
 * egnn.py; names in SURVEY.md section 8 a-12):
 *   bloch_wave_reciprocal_lattice_vectors [n_bloch,3], projection_matrices [3,2n,2n],
 *   egnn.embedding_in.{weight,bias},
 *   per layer: message_mlp.{0,2,..}.{weight,bias}, node_mlp.{0,2,..,last}.{weight,bias},
 *              coord_mlp.{0,2,..}.{weight,bias}, coord_mlp.last.weight, [att_mlp.0.{weight,bias}],
 *   egnn.node_classification_layer.{weight,bias}.
 * b200md_egnn_param_count gives the expected number of tensors, b200md_egnn_packed_floats the size of the
 * packed image (fp32 words) the caller must allocate; b200md_egnn_pack_weights fills it (device->device).
 * `param_ptrs_host` is a HOST array of DEVICE pointers.                                            */
int32_t b200md_egnn_param_count(const b200md_egnn_config* cfg);
int64_t b200md_egnn_packed_floats(const b200md_egnn_config* cfg);
int b200md_egnn_pack_weights(const b200md_egnn_config* cfg, const float* const* param_ptrs_host,
                             int32_t n_params, float* packed, b200md_stream_t stream);

/* Workspace (bytes) for a forward over B structures of N atoms with room for `edge_capacity` directed
 * edges in total.  b200md_egnn_default_edge_capacity gives a safe default (all pairs when N is small).   */
int64_t b200md_egnn_default_edge_capacity(const b200md_egnn_config* cfg, int64_t n_samples,
                                          int32_t n_atoms, const float* cell_lengths_host_or_null);
int64_t b200md_egnn_workspace_bytes(const b200md_egnn_config* cfg, int64_t n_samples,
                                    int32_t n_atoms, int64_t edge_capacity);

#define B200MD_IMPL_AUTO 0
#define B200MD_IMPL_FFMA 1    /* fp32 CUDA-core tiles */
#define B200MD_IMPL_TCGEN05 2 /* 3xTF32 tcgen05.mma tiles, fp32 accumulate in TMEM */
#define B200MD_IMPL_TCGEN05_F16X3 3 /* 3xFP16 tcgen05.mma tiles (same 11-bit significands, half the operand bytes);
                                      activations beyond the fp16 range raise B200MD_STATUS_TC_RANGE */
#define B200MD_IMPL_TCGEN05_F16X3_TS 4 /* 3xFP16 with the edge kernel's A operand in tensor memory: accumulators are
                                         converted in place into the next layer's operand (no shared-memory round trip) */

/* The implementation B200MD_IMPL_AUTO resolves to for this configuration (B200MD_IMPL_FFMA or B200MD_IMPL_TCGEN05),
 * or a negative error when `impl` is not available for it. */
int b200md_egnn_resolve_impl(const b200md_egnn_config* cfg, int32_t impl);

/* atom_types [B,N] int64, x [B,N,3] in [0,1), lattice [B,6], sigma [B] ->
 * logits_out [B,N,C] (C = num_atom_types+1; MASK logit = -inf when mask_logit_to_neg_inf),
 * score_out [B,N,3] (sigma-normalised score).  The lattice output of the reference is identically zero
 * (egnn_score_network.py:296-300) and is produced by the host wrapper.
 * Optional debug output: n_edges_out (device int64[1]) total number of edges of the graph used.        */
int b200md_egnn_forward(const b200md_egnn_config* cfg, const float* packed,
                        const int64_t* atom_types, const float* x, const float* lattice,
                        const float* sigma, int64_t n_samples, int32_t n_atoms,
                        int32_t mask_logit_to_neg_inf, float* logits_out, float* score_out,
                        void* workspace, int64_t workspace_bytes, int64_t edge_capacity,
                        int64_t* n_edges_out_or_null, int32_t* status, int32_t impl,
                        b200md_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Whole predictor-corrector trajectory: PredictorCorrectorAXLGenerator.sample_from_noisy_composition
 * (generators/predictor_corrector_axl_generator.py:112-161) driving LangevinGenerator.predictor_step /
 * corrector_step (generators/langevin_generator.py:536-645,693-805) with the EGNN score network, every
 * launch enqueued from one host call, noise drawn in-kernel, no host synchronisation.
 * One predictor-corrector step (~27 kernels) is captured into a CUDA graph that reads its step-dependent scalars
 * from a device table (csrc/sampler.cuh) and is replayed once per time index; trajectories longer than
 * B200MD_PC_MAX_GRAPH_STEPS steps, B200MD_PC_DIRECT=1 in the environment, or per-class profiling use direct launches.
 * Schedule tables are DEVICE arrays from NoiseScheduler (noise_schedulers/noise_scheduler.py:348-378):
 * sigma,g,g2 [T]; epsilon [T]; q,q_bar,q_bar_tm1 [T,C,C]; the host copies `*_host` [T] drive kernel arguments.
 * In place on atom_types [B,N], x [B,N,3]; lattice [B,6] is fixed (use_fixed_lattice_parameters=True).
 * Optional constraint (RePaint, constrained_langevin_generator.py) when n_constrained > 0.
 * ---------------------------------------------------------------------------------------------- */
#define B200MD_PC_MAX_GRAPH_STEPS 4096   /* rows of the device-side step table in the sampler workspace (64 B each) */
typedef struct {
  int32_t total_time_steps;           /* T */
  int32_t number_of_corrector_steps;
  int32_t start_step_index;           /* starting_step_index (T for a full trajectory) */
  int32_t end_step_index;             /* ending_step_index (0) */
  int32_t atom_flags;                 /* B200MD_ATOM_GREEDY | B200MD_ATOM_ONE_TRANSITION */
  int32_t atom_type_transition_in_corrector;
  float small_epsilon;
  float sigma_min;                    /* corrector at i == 0 (langevin_generator.py:719-725) */
  const float* sigma_host;            /* [T] */
  const float* g_host;                /* [T] */
  const float* g2_host;               /* [T] */
  const float* epsilon_host;          /* [T] */
  const float* sqrt_2_epsilon_host;   /* [T] */
  const float* q;                     /* device [T,C,C] */
  const float* q_bar;                 /* device [T,C,C] */
  const float* q_bar_tm1;             /* device [T,C,C] */
  /* RePaint */
  int32_t n_constrained;
  const float* x0_known;              /* device [Nc,3] */
  const int64_t* a0_known;            /* device [Nc] */
  const int64_t* constrained_indices; /* device [Nc] */
} b200md_pc_schedule;

int b200md_pc_sample(const b200md_egnn_config* cfg, const float* packed,
                     const b200md_pc_schedule* schedule, int64_t* atom_types, float* x,
                     const float* lattice, int64_t n_samples, int32_t n_atoms, uint64_t seed,
                     uint64_t sample_offset, void* workspace, int64_t workspace_bytes,
                     int64_t edge_capacity, int32_t* status, int32_t impl, b200md_stream_t stream);
/* 1 if the last b200md_pc_sample call of this thread replayed the captured step graph, 0 if it launched directly */
int b200md_pc_sample_last_mode(void);
/* whole workspace of b200md_pc_sample: b200md_egnn_workspace_bytes + logits + scores + the device-side step table */
int64_t b200md_pc_sample_workspace_bytes(const b200md_egnn_config* cfg, int64_t n_samples,
                                         int32_t n_atoms, int64_t edge_capacity);

#ifdef __cplusplus
}
#endif
#endif /* B200MD_H_ */
