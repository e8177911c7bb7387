// Whole predictor-corrector trajectory enqueued from one host call.
// Reference control flow: generators/predictor_corrector_axl_generator.py:112-161 (time loop),
// generators/langevin_generator.py:536-645 (predictor_step), :693-805 (corrector_step),
// generators/constrained_langevin_generator.py:94-163 (RePaint after each predictor step).
//
// Two ways to enqueue the loop, same kernels and bit-identical arguments:
//   * graph replay (default): the step-dependent scalars of every time index (sigma_i, g_i^2, eps_i, draw keys, table
//     rows, last-step flags) are written once into a device table (sampler.cuh); ONE predictor-corrector step --
//     ~27 kernels -- is captured into a CUDA graph whose kernels read "the current row", and whose last node copies the
//     next row in.  The trajectory is then one cudaGraphLaunch per step: no per-kernel launch cost on the host, no
//     argument marshalling, and the launch gaps between the dependent kernels of a step shrink to the graph's own
//     (this is what bounds small batches: 32-64 samples per GPU, the reference's own job shapes);
//   * direct launches (B200MD_PC_DIRECT=1, while profiling per kernel class, or if capture is refused): the plain host
//     loop over the same launch functions with the scalars passed as kernel arguments.
#include <stdlib.h>

#include <vector>

#include "egnn.cuh"
#include "sampler.cuh"

using namespace b200md;

namespace {

__global__ void pc_advance_kernel(PcStepScalars* current, const PcStepScalars* __restrict__ table, int32_t* cursor,
                                  int32_t n_steps) {
  // one thread: current <- table[++cursor] (the row after the last one is never read)
  const int next = *cursor + 1;
  *cursor = next;
  if (next < n_steps) *current = table[next];
}

__global__ void pc_load_first_kernel(PcStepScalars* current, const PcStepScalars* __restrict__ table, int32_t* cursor) {
  *cursor = 0;
  *current = table[0];
}

struct PcDeviceState {
  PcStepScalars* table;
  PcStepScalars* current;
  int32_t* cursor;
};

int64_t pc_state_bytes(int n_steps) { return (int64_t)(n_steps + 1) * (int64_t)sizeof(PcStepScalars) + 256; }

// The arguments of one predictor-corrector step.  `cur` == nullptr: scalars of time index i passed as launch
// arguments (direct mode); otherwise every kernel reads them from *cur (graph mode) and `i` is ignored.
struct StepContext {
  const b200md_egnn_config* cfg; const EgnnLayout* lay; const float* packed; const b200md_pc_schedule* s;
  int64_t* atom_types; float* x; const float* lattice; int64_t B; int32_t N; int C;
  uint64_t seed, sample_offset;
  const EgnnWorkspace* ws; int64_t edge_capacity; float* logits; float* score; int32_t* status; int32_t impl;
  cudaStream_t st;
};

PcStepScalars scalars_of_step(const b200md_pc_schedule* s, int i) {
  PcStepScalars r{};
  const int ncorr = s->number_of_corrector_steps;
  const bool last = (i == 0);
  r.sigma_pred = s->sigma_host[i]; r.g2 = s->g2_host[i]; r.g = s->g_host[i];
  const int idx_corr = last ? 0 : i - 1;
  r.sigma_corr = last ? s->sigma_min : s->sigma_host[idx_corr];
  r.eps = s->epsilon_host[i]; r.sqrt_2eps = s->sqrt_2_epsilon_host[i];
  r.repaint_sigma = last ? -1.0f : s->sigma_host[i - 1];
  r.step_id = (uint32_t)i * (uint32_t)(1 + ncorr);
  int flags = s->atom_flags & B200MD_ATOM_GREEDY;
  if ((s->atom_flags & B200MD_ATOM_ONE_TRANSITION) && !last) flags |= B200MD_ATOM_ONE_TRANSITION;
  if (last) flags |= B200MD_ATOM_LAST_STEP;
  r.atom_flags_pred = flags;
  r.atom_flags_corr = s->atom_flags;
  r.q_index_pred = i; r.q_index_corr = idx_corr; r.q_index_repaint = last ? -1 : i - 1;
  return r;
}

int enqueue_step(const StepContext& c, int i, const PcStepScalars* cur) {
  const b200md_pc_schedule* s = c.s;
  const int ncorr = s->number_of_corrector_steps;
  const int64_t CC = (int64_t)c.C * c.C;
  const PcStepScalars v = cur ? PcStepScalars{} : scalars_of_step(s, i);
  int rc;
  // ---- predictor step with index_i = i + 1, idx = i ----
  rc = egnn_forward_impl(*c.cfg, *c.lay, c.packed, c.atom_types, c.x, c.lattice, cur ? &cur->sigma_pred : nullptr,
                         v.sigma_pred, 0, c.B, c.N, 1, c.logits, c.score, *c.ws, c.edge_capacity, c.status, c.impl, c.st);
  if (rc != B200MD_OK) return rc;
  const int64_t row_pred = cur ? 0 : (int64_t)v.q_index_pred * CC;
  rc = atom_types_update_impl(c.logits, c.atom_types, s->q + row_pred, s->q_bar + row_pred, s->q_bar_tm1 + row_pred, nullptr,
                              nullptr, v.atom_flags_pred, s->small_epsilon, c.B, c.N, c.C, c.seed, v.step_id,
                              c.sample_offset, c.atom_types, nullptr, c.status, cur, 0, c.st);
  if (rc != B200MD_OK) return rc;
  rc = pc_coordinates_update_impl(c.x, c.score, nullptr, v.g2, v.sigma_pred, v.g, c.B, c.N, 3, c.seed, v.step_id,
                                  c.sample_offset, c.x, cur, 0, c.st);
  if (rc != B200MD_OK) return rc;
  if (s->n_constrained > 0) {
    // _repaint_composition(index_i - 1 = i): noise the known atoms to time index i (hard constraint at i == 0)
    const float* qb = cur ? s->q_bar : (v.q_index_repaint >= 0 ? s->q_bar + (int64_t)v.q_index_repaint * CC : s->q_bar);
    rc = repaint_impl(c.x, c.atom_types, s->x0_known, s->a0_known, s->constrained_indices, s->n_constrained,
                      v.repaint_sigma, qb, nullptr, nullptr, c.B, c.N, 3, c.C, c.seed, v.step_id, c.sample_offset, cur, c.st);
    if (rc != B200MD_OK) return rc;
  }
  // ---- corrector steps with index_i = i ----
  for (int k = 0; k < ncorr; ++k) {
    const uint32_t step_id = v.step_id + 1u + (uint32_t)k;
    rc = egnn_forward_impl(*c.cfg, *c.lay, c.packed, c.atom_types, c.x, c.lattice, cur ? &cur->sigma_corr : nullptr,
                           v.sigma_corr, 0, c.B, c.N, 1, c.logits, c.score, *c.ws, c.edge_capacity, c.status, c.impl, c.st);
    if (rc != B200MD_OK) return rc;
    rc = pc_coordinates_update_impl(c.x, c.score, nullptr, v.eps, v.sigma_corr, v.sqrt_2eps, c.B, c.N, 3, c.seed, step_id,
                                    c.sample_offset, c.x, cur, 1 + k, c.st);
    if (rc != B200MD_OK) return rc;
    if (s->atom_type_transition_in_corrector) {
      const int64_t row = cur ? 0 : (int64_t)v.q_index_corr * CC;
      rc = atom_types_update_impl(c.logits, c.atom_types, s->q + row, s->q_bar + row, s->q_bar_tm1 + row, nullptr, nullptr,
                                  v.atom_flags_corr, s->small_epsilon, c.B, c.N, c.C, c.seed, step_id, c.sample_offset,
                                  c.atom_types, nullptr, c.status, cur, 1 + k, c.st);
      if (rc != B200MD_OK) return rc;
    }
  }
  return B200MD_OK;
}

thread_local int g_last_mode = 0;   // 1: the last b200md_pc_sample of this thread replayed a graph, 0: direct launches

// Stream capture is refused on the legacy default stream (which is what torch.cuda.current_stream() is unless the caller
// changes it), so the step is always captured on a private non-blocking stream -- nothing executes during capture --
// and the instantiated graph is launched on the caller's stream.  One capture stream per device, created on first use.
cudaStream_t capture_stream() {
  static cudaStream_t streams[PerDeviceOnce::kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= PerDeviceOnce::kMaxDevices) return nullptr;
  if (streams[dev] == nullptr && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    streams[dev] = nullptr;
  }
  return streams[dev];
}

bool graph_mode_wanted() {
  if (g_profile_on) return false;   // per-class timing records events around every launch
  const char* env = getenv("B200MD_PC_DIRECT");
  return !(env && env[0] == '1');
}

}  // namespace

extern "C" {

int64_t b200md_pc_sample_workspace_bytes(const b200md_egnn_config* cfg, int64_t B, int32_t N, int64_t cap) {
  if (!cfg || B < 0 || N <= 0 || cap < 0) return B200MD_ERR_INVALID_ARGUMENT;
  EgnnWorkspace ws;
  int64_t base = carve_workspace(*cfg, B, N, cap, nullptr, ws);
  const int64_t C = cfg->num_atom_types + 1;
  int64_t logits = (B * N * C * 4 + 255) & ~int64_t(255);
  int64_t score = (B * N * 3 * 4 + 255) & ~int64_t(255);
  return base + logits + score + pc_state_bytes(B200MD_PC_MAX_GRAPH_STEPS);
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
  g_last_mode = 0;
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
  char* wbase = reinterpret_cast<char*>(workspace);
  float* logits = reinterpret_cast<float*>(wbase + base);
  float* score = reinterpret_cast<float*>(wbase + base + logits_bytes);
  cudaStream_t st = as_stream(stream);

  StepContext c{};
  c.cfg = cfg; c.lay = &lay; c.packed = packed; c.s = s; c.atom_types = atom_types; c.x = x; c.lattice = lattice;
  c.B = B; c.N = N; c.C = C; c.seed = seed; c.sample_offset = sample_offset; c.ws = &ws; c.edge_capacity = edge_capacity;
  c.logits = logits; c.score = score; c.status = status; c.impl = impl; c.st = st;

  const int n_steps = s->start_step_index - s->end_step_index;
  const int64_t state_offset = base + logits_bytes + score_bytes;
  const bool table_fits = n_steps <= B200MD_PC_MAX_GRAPH_STEPS && state_offset + pc_state_bytes(n_steps) <= workspace_bytes;
  cudaStream_t cap = (graph_mode_wanted() && table_fits && n_steps >= 2) ? capture_stream() : nullptr;
  if (cap != nullptr) {
    PcDeviceState dev;
    dev.cursor = reinterpret_cast<int32_t*>(wbase + state_offset);
    dev.current = reinterpret_cast<PcStepScalars*>(wbase + state_offset + 128);
    dev.table = dev.current + 1;
    std::vector<PcStepScalars> rows((size_t)n_steps);
    for (int j = 0; j < n_steps; ++j) rows[(size_t)j] = scalars_of_step(s, s->start_step_index - 1 - j);
    // pageable source: the call returns once the rows are staged, so `rows` may go out of scope afterwards
    cudaError_t e = cudaMemcpyAsync(dev.table, rows.data(), sizeof(PcStepScalars) * (size_t)n_steps, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return set_cuda_error(e);
    pc_load_first_kernel<<<1, 1, 0, st>>>(dev.current, dev.table, dev.cursor);
    B200MD_CHECK_LAUNCH();

    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const unsigned long long launches_before = g_launch_count;
    e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed);
    if (e == cudaSuccess) {
      c.st = cap;
      rc = enqueue_step(c, 0, dev.current);
      c.st = st;
      if (rc == B200MD_OK) {
        pc_advance_kernel<<<1, 1, 0, cap>>>(dev.current, dev.table, dev.cursor, n_steps);
        ++g_launch_count;
      }
      e = cudaStreamEndCapture(cap, &graph);   // always end the capture, also after an error inside it
      if (rc != B200MD_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
      if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
    }
    if (e == cudaSuccess && exec != nullptr) {
      const unsigned long long per_step = g_launch_count - launches_before;
      g_launch_count = launches_before;
      for (int j = 0; j < n_steps; ++j) {
        e = cudaGraphLaunch(exec, st);
        if (e != cudaSuccess) break;
        g_launch_count += per_step;   // kernels executed, as in direct mode
      }
      // an executable graph that is still in flight is freed asynchronously when its launches complete
      cudaGraphExecDestroy(exec);
      if (e != cudaSuccess) return set_cuda_error(e);
      g_last_mode = 1;
      return B200MD_OK;
    }
    cudaGetLastError();   // capture or instantiation refused (e.g. the caller is itself capturing): direct launches
    g_launch_count = launches_before;
  }

  for (int i = s->start_step_index - 1; i >= s->end_step_index; --i) {
    rc = enqueue_step(c, i, nullptr);
    if (rc != B200MD_OK) return rc;
  }
  return B200MD_OK;
}

int b200md_pc_sample_last_mode(void) { return g_last_mode; }

}  // extern "C"
