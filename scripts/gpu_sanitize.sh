#!/bin/bash
# compute-sanitizer memcheck over one production-width forward (tensor-memory edge / node kernels, input warps, CUDA-graph
# free path) and a 2-step trajectory of a small batch; log in gpurun_out/sanitize.log
set -u
mkdir -p gpurun_out
cat > /tmp/sanitize_case.py <<'PY'
import sys
import torch
sys.path.insert(0, ".")
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import PredictorCorrectorSamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import EGNNScoreNetwork, EGNNScoreNetworkParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
dev = torch.device("cuda:0")
torch.manual_seed(0)
hp = dict(num_atom_types=2, edges="radial_cutoff", radial_cutoff=7.5, n_layers=4, message_n_hidden_dimensions=4,
          message_hidden_dimensions_size=256, node_n_hidden_dimensions=4, node_hidden_dimensions_size=256,
          coordinate_n_hidden_dimensions=4, coordinate_hidden_dimensions_size=256)
net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp)).to(dev)
B, N = 7, 72   # 504 nodes: three full tiles and a ragged one
sampling = PredictorCorrectorSamplingParameters(number_of_samples=B, number_of_atoms=N, num_atom_types=2,
                                                use_fixed_lattice_parameters=True, cell_dimensions=[11.3] * 3)
gen = LangevinGenerator(NoiseParameters(total_time_steps=2), sampling, net)
gen.seed = 5
out = gen.sample(B, dev)
torch.cuda.synchronize()
assert torch.isfinite(out.X).all()
print("sanitize case ok", net.last_edge_kernel_name(), tuple(out.X.shape))
PY
B200MD_PC_DIRECT=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/sanitize_case.py > gpurun_out/sanitize.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|sanitize case ok|Invalid|out of bounds" gpurun_out/sanitize.log | head -10
