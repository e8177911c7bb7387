"""EGNN score network on sm_100a (reference: models/score_networks/egnn_score_network.py:24-303).

Constructor arguments, dataclass fields/defaults, parameter names and the ``forward`` contract are those of
the reference.  ``_forward_unchecked`` hands the whole evaluation -- radius graph, Bloch-wave uplift, node
features, the E_GCL layers, classification logits and the score projection -- to ``b200md_egnn_forward``
(csrc/egnn.cu).  Parameters are read from a packed image that is rebuilt whenever a parameter changes
(``load_state_dict``, ``.to()``, in-place edits).
"""
import ctypes
from dataclasses import dataclass
from typing import AnyStr, Dict, List, Optional, Union

import numpy as np
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.models.egnn import EGNN
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import (
    ScoreNetwork, ScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL, NOISE, NOISY_AXL_COMPOSITION, TIME
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.lattice_utils import \
    get_cubic_point_group_positive_normalized_bloch_wave_vectors


@dataclass(kw_only=True)
class EGNNScoreNetworkParameters(ScoreNetworkParameters):
    """Hyper-parameters of the EGNN score network (reference :24-45)."""

    architecture: str = "egnn"
    number_of_bloch_wave_shells: int = 1
    message_n_hidden_dimensions: int = 1
    message_hidden_dimensions_size: int = 16
    node_n_hidden_dimensions: int = 1
    node_hidden_dimensions_size: int = 32
    coordinate_n_hidden_dimensions: int = 1
    coordinate_hidden_dimensions_size: int = 32
    residual: bool = True
    attention: bool = False
    normalize: bool = False
    tanh: bool = False
    coords_agg: str = "mean"
    message_agg: str = "mean"
    n_layers: int = 4
    edges: str = "fully_connected"
    radial_cutoff: Union[float, None] = None
    drop_duplicate_edges: bool = True


class EGNNScoreNetwork(ScoreNetwork):
    """Score network built on EGNN (Satorras et al.) with a Bloch-wave uplift of the torus coordinates."""

    def __init__(self, hyper_params: EGNNScoreNetworkParameters):
        super().__init__(hyper_params)
        self.number_of_features_per_node = self.num_atom_types + 2  # sigma + one-hot over types incl. MASK
        self.number_of_bloch_wave_shells = hyper_params.number_of_bloch_wave_shells
        bloch = get_cubic_point_group_positive_normalized_bloch_wave_vectors(
            number_of_complete_shells=self.number_of_bloch_wave_shells, spatial_dimension=self.spatial_dimension
        ).float()
        self.register_parameter("bloch_wave_reciprocal_lattice_vectors", torch.nn.Parameter(bloch, requires_grad=False))
        self.register_parameter(
            "projection_matrices",
            torch.nn.Parameter(self._create_block_diagonal_projection_matrices(bloch), requires_grad=False),
        )
        self.edges = hyper_params.edges
        assert self.edges in ["fully_connected", "radial_cutoff"], \
            f"Edges type should be fully_connected or radial_cutoff. Got {self.edges}"
        self.radial_cutoff = hyper_params.radial_cutoff
        if self.edges == "fully_connected":
            assert self.radial_cutoff is None, "Specifying a radial cutoff is inconsistent with edges=fully_connected."
        else:
            assert type(self.radial_cutoff) is float, \
                "A floating point value for the radial cutoff is needed for edges=radial_cutoff."
        self.drop_duplicate_edges = hyper_params.drop_duplicate_edges
        self.egnn = EGNN(
            input_size=self.number_of_features_per_node,
            message_n_hidden_dimensions=hyper_params.message_n_hidden_dimensions,
            message_hidden_dimensions_size=hyper_params.message_hidden_dimensions_size,
            node_n_hidden_dimensions=hyper_params.node_n_hidden_dimensions,
            node_hidden_dimensions_size=hyper_params.node_hidden_dimensions_size,
            coordinate_n_hidden_dimensions=hyper_params.coordinate_n_hidden_dimensions,
            coordinate_hidden_dimensions_size=hyper_params.coordinate_hidden_dimensions_size,
            residual=hyper_params.residual,
            attention=hyper_params.attention,
            normalize=hyper_params.normalize,
            tanh=hyper_params.tanh,
            coords_agg=hyper_params.coords_agg,
            message_agg=hyper_params.message_agg,
            n_layers=hyper_params.n_layers,
            num_classes=self.num_atom_types + 1,
        )
        # ---- sm_100a state (not part of the state_dict) ----
        self.kernel_impl = _cabi.IMPL_AUTO  # IMPL_FFMA / IMPL_TCGEN05 can be forced for A/B checks
        # Set (and kept) once an activation of THIS network left the fp16 range of the 3xFP16 kernels under IMPL_AUTO:
        # from then on AUTO means the 3xTF32 kernels, which have the fp32 exponent range.
        self._auto_fallback_impl: Optional[int] = None
        self.edge_capacity: Optional[int] = None  # None -> b200md_egnn_default_edge_capacity
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._workspace: Optional[torch.Tensor] = None
        self._workspace_key = None
        self._fuse_mask_logit = False
        self._n_edges: Optional[torch.Tensor] = None  # device int64[1]: edges of the last graph built by forward()

    def last_number_of_edges(self) -> int:
        """Total number of directed edges of the graph used by the last ``forward`` (one host sync)."""
        return 0 if self._n_edges is None else int(self._n_edges.item())

    def effective_kernel_impl(self) -> int:
        """``kernel_impl``, or the wider-range implementation IMPL_AUTO has fallen back to for this network."""
        if self.kernel_impl == _cabi.IMPL_AUTO and self._auto_fallback_impl is not None:
            return self._auto_fallback_impl
        return int(self.kernel_impl)

    def widen_range_after_overflow(self, word: Optional[int] = None) -> bool:
        """IMPL_AUTO only: if the status word says an activation overflowed fp16 (STATUS_TC_RANGE), clear the bit,
        switch this network to the 3xTF32 kernels for good and tell the caller to evaluate again.  A forced
        ``kernel_impl`` keeps the bit, which is then raised as an AssertionError."""
        if self.kernel_impl != _cabi.IMPL_AUTO or self._auto_fallback_impl is not None or self._status is None:
            return False
        cfg = self._config()
        lib = _cabi.lib()
        if lib.b200md_egnn_resolve_impl(ctypes.byref(cfg), _cabi.IMPL_AUTO) not in (_cabi.IMPL_TCGEN05_F16X3,
                                                                                   _cabi.IMPL_TCGEN05_F16X3_TS):
            return False
        if lib.b200md_egnn_resolve_impl(ctypes.byref(cfg), _cabi.IMPL_TCGEN05) != _cabi.IMPL_TCGEN05:
            return False
        if word is None:
            word = _cabi.status_word(self._status)
        if not word & _cabi.STATUS_TC_RANGE:
            return False
        self._status.zero_()  # inf / NaN downstream of the overflow may have raised other bits; the re-run raises the real ones
        self._auto_fallback_impl = _cabi.IMPL_TCGEN05
        return True

    def last_edge_kernel_name(self) -> str:
        """Which edge-MLP implementation ``kernel_impl`` resolves to for this configuration."""
        cfg = self._config()
        resolved = _cabi.lib().b200md_egnn_resolve_impl(ctypes.byref(cfg), self.effective_kernel_impl())
        return {_cabi.IMPL_FFMA: "egnn_edge_kernel (fp32 FFMA tiles)",
                _cabi.IMPL_TCGEN05: "egnn_edge_tc_kernel (tcgen05 3xTF32, fp32 accumulate in TMEM)",
                _cabi.IMPL_TCGEN05_F16X3: "egnn_edge_tc_kernel (tcgen05 3xFP16, fp32 accumulate in TMEM)",
                _cabi.IMPL_TCGEN05_F16X3_TS: "egnn_edge_ts_kernel (tcgen05 cta_group::2 3xFP16, A operand and fp32 "
                                             "accumulators in TMEM)"}.get(
                    resolved, f"unsupported ({resolved})")

    # ---- reference helpers kept for API compatibility ----------------------------------------------------
    @staticmethod
    def _create_block_diagonal_projection_matrices(bloch_wave_reciprocal_lattice_vectors: torch.Tensor) -> torch.Tensor:
        """Gamma^alpha = blockdiag_k( k_alpha [[0,-1],[1,0]] ), shape [d, 2 n_bloch, 2 n_bloch] (reference :131-162)."""
        rotation = torch.tensor([[0.0, -1.0], [1.0, 0.0]])
        per_axis = [torch.block_diag(*[k * rotation for k in column]) for column in
                    bloch_wave_reciprocal_lattice_vectors.transpose(0, 1)]
        return torch.stack(per_axis)

    # ---- configuration / packing -------------------------------------------------------------------------
    def _config(self) -> _cabi.EgnnConfig:
        hp = self._hyper_params
        flags = 0
        flags |= _cabi.EGNN_RESIDUAL if hp.residual else 0
        flags |= _cabi.EGNN_ATTENTION if hp.attention else 0
        flags |= _cabi.EGNN_NORMALIZE if hp.normalize else 0
        flags |= _cabi.EGNN_TANH if hp.tanh else 0
        flags |= _cabi.EGNN_COORDS_MEAN if hp.coords_agg == "mean" else 0
        flags |= _cabi.EGNN_MESSAGE_MEAN if hp.message_agg == "mean" else 0
        flags |= _cabi.EGNN_RADIAL_CUTOFF if self.edges == "radial_cutoff" else 0
        flags |= 0 if hp.drop_duplicate_edges else _cabi.EGNN_KEEP_DUPLICATES
        rc = float(self.radial_cutoff) if self.radial_cutoff is not None else 0.0
        return _cabi.EgnnConfig(
            spatial_dimension=self.spatial_dimension,
            num_atom_types=self.num_atom_types,
            n_layers=hp.n_layers,
            n_bloch=int(self.bloch_wave_reciprocal_lattice_vectors.shape[0]),
            node_hidden=hp.node_hidden_dimensions_size,
            node_n_hidden=hp.node_n_hidden_dimensions,
            message_hidden=hp.message_hidden_dimensions_size,
            message_n_hidden=hp.message_n_hidden_dimensions,
            coord_hidden=hp.coordinate_hidden_dimensions_size,
            coord_n_hidden=hp.coordinate_n_hidden_dimensions,
            flags=flags,
            radial_cutoff=rc,
            # `.clip(min=2.2 * radial_cutoff)` on a float32 tensor (reference :236-238)
            lattice_clip_min=float(np.float32(2.2 * rc)),
        )

    def _ordered_parameters(self) -> List[torch.Tensor]:
        """Parameters in the order b200md_egnn_pack_weights expects (include/b200md.h)."""
        out = [self.bloch_wave_reciprocal_lattice_vectors, self.projection_matrices,
               self.egnn.embedding_in.weight, self.egnn.embedding_in.bias]
        for layer in self.egnn.graph_layers:
            for mlp in (layer.message_mlp, layer.node_mlp, layer.coord_mlp):
                for module in mlp:
                    if isinstance(module, torch.nn.Linear):
                        out.append(module.weight)
                        if module.bias is not None:
                            out.append(module.bias)
            if layer.attention:
                out += [layer.att_mlp[0].weight, layer.att_mlp[0].bias]
        out += [self.egnn.node_classification_layer.weight, self.egnn.node_classification_layer.bias]
        return out

    def packed_weights(self, device: torch.device) -> torch.Tensor:
        """Packed weight image on ``device``; rebuilt when any parameter storage or version changes."""
        params = self._ordered_parameters()
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        lib = _cabi.lib()
        cfg = self._config()
        n_floats = lib.b200md_egnn_packed_floats(ctypes.byref(cfg))
        _cabi.check(n_floats)
        expected = lib.b200md_egnn_param_count(ctypes.byref(cfg))
        assert expected == len(params), f"parameter list mismatch: {expected} expected, {len(params)} found"
        staged = [_cabi.as_f32(p).to(device) for p in params]
        pointers = (ctypes.c_void_p * len(staged))(*[t.data_ptr() for t in staged])
        packed = torch.empty(int(n_floats), dtype=torch.float32, device=device)
        _cabi.check(lib.b200md_egnn_pack_weights(ctypes.byref(cfg), pointers, len(staged), _cabi.ptr(packed),
                                                 _cabi.stream_ptr(device)))
        torch.cuda.current_stream(device).synchronize()  # `staged` may be temporaries
        self._packed, self._packed_key = packed, key
        return packed

    def resolve_edge_capacity(self, batch_size: int, number_of_atoms: int) -> int:
        if self.edge_capacity is not None:
            return int(self.edge_capacity)
        cfg = self._config()
        cap = _cabi.lib().b200md_egnn_default_edge_capacity(ctypes.byref(cfg), batch_size, number_of_atoms, None)
        _cabi.check(cap)
        return int(cap)

    def workspace(self, device: torch.device, batch_size: int, number_of_atoms: int, capacity: int,
                  for_sampler: bool = False) -> torch.Tensor:
        key = (str(device), batch_size, number_of_atoms, capacity, for_sampler)
        if self._workspace is not None and self._workspace_key == key:
            return self._workspace
        cfg = self._config()
        fn = _cabi.lib().b200md_pc_sample_workspace_bytes if for_sampler else _cabi.lib().b200md_egnn_workspace_bytes
        n_bytes = fn(ctypes.byref(cfg), batch_size, number_of_atoms, capacity)
        _cabi.check(n_bytes)
        self._workspace = None
        self._workspace = torch.empty(int(n_bytes), dtype=torch.uint8, device=device)
        self._workspace_key = key
        return self._workspace

    # ---- forward ---------------------------------------------------------------------------------------------
    def _check_batch_values(self, batch: Dict[AnyStr, torch.Tensor]):
        """x in [0,1), atom types in range and sigma > 0 are checked inside the node kernel; only `time` here."""
        t = batch[TIME]
        if t.is_cuda:
            bad = torch.logical_or(t < 0.0, t > 1.0).any().to(torch.int32) * _cabi.STATUS_X_OUT_OF_RANGE
            self.device_status(t.device).bitwise_or_(bad)  # no host sync

    def _impose_non_mask_atomic_type_prediction(self, output: AXL):
        if not self._fuse_mask_logit:  # the kernel already wrote -inf when the flag is set
            super()._impose_non_mask_atomic_type_prediction(output)

    def forward(self, batch: Dict[AnyStr, torch.Tensor], conditional: Optional[bool] = None) -> AXL:
        self._fuse_mask_logit = True
        try:
            return super().forward(batch, conditional)
        finally:
            self._fuse_mask_logit = False

    def _forward_unchecked(self, batch: Dict[AnyStr, torch.Tensor], conditional: bool = False) -> AXL:
        composition = batch[NOISY_AXL_COMPOSITION]
        device = _cabi.require_cuda(composition.X, composition.A, composition.L)
        x = _cabi.as_f32(composition.X)
        a = _cabi.as_i64(composition.A)
        lattice = _cabi.as_f32(composition.L)
        sigma = _cabi.as_f32(batch[NOISE].to(device)).reshape(-1)
        batch_size, number_of_atoms, _ = x.shape
        num_classes = self.num_atom_types + 1
        cfg = self._config()
        packed = self.packed_weights(device)
        capacity = self.resolve_edge_capacity(batch_size, number_of_atoms)
        workspace = self.workspace(device, batch_size, number_of_atoms, capacity)
        logits = torch.empty(batch_size, number_of_atoms, num_classes, dtype=torch.float32, device=device)
        scores = torch.empty(batch_size, number_of_atoms, 3, dtype=torch.float32, device=device)
        status = self.device_status(device)
        if self._n_edges is None or self._n_edges.device != device:
            self._n_edges = torch.zeros(1, dtype=torch.int64, device=device)
        _cabi.check(_cabi.lib().b200md_egnn_forward(
            ctypes.byref(cfg), _cabi.ptr(packed), _cabi.ptr(a), _cabi.ptr(x), _cabi.ptr(lattice), _cabi.ptr(sigma),
            batch_size, number_of_atoms, int(self._fuse_mask_logit), _cabi.ptr(logits), _cabi.ptr(scores),
            _cabi.ptr(workspace), workspace.numel(), capacity, _cabi.ptr(self._n_edges), _cabi.ptr(status),
            self.effective_kernel_impl(), _cabi.stream_ptr(device)))
        return AXL(A=logits, X=scores, L=torch.zeros_like(composition.L))
