"""Parameter containers of the EGNN (reference: models/egnn.py:21-385).

``E_GCL`` and ``EGNN`` reproduce the reference's module tree -- same attribute names, same ``nn.Linear``
shapes, same construction order (so the same ``torch.manual_seed`` yields the same initial weights and a
reference ``state_dict`` loads unchanged).  They hold parameters only: the arithmetic of
``E_GCL.forward`` / ``EGNN.forward`` runs in the fused sm_100a kernels driven by ``EGNNScoreNetwork``
(csrc/egnn.cu, csrc/egnn_tc.cu), which read these parameters through a packed weight image.
"""
from typing import Callable

import torch
from torch import nn

from diffusion_for_multi_scale_molecular_dynamics_b200._cabi import B200MDError

_FUSED_ONLY = (
    "EGNN layers are evaluated by the fused sm_100a kernels through EGNNScoreNetwork.forward(); "
    "the module tree only holds the parameters."
)


def _mlp(sizes, act_fn, final_activation: bool, final_bias: bool = True) -> nn.Sequential:
    """Linear/act chain with the reference's Sequential indexing (Linear at even positions)."""
    layers = []
    last = len(sizes) - 2
    for i, (fan_in, fan_out) in enumerate(zip(sizes[:-1], sizes[1:])):
        layers.append(nn.Linear(fan_in, fan_out, bias=final_bias or i < last))
        if i < last or final_activation:
            layers.append(act_fn)
    return nn.Sequential(*layers)


class E_GCL(nn.Module):
    """One equivariant graph convolution layer: message phi_e, node phi_h and coordinate phi_x MLPs (:21-134)."""

    def __init__(
        self,
        input_size: int,
        output_size: int,
        message_n_hidden_dimensions: int,
        message_hidden_dimensions_size: int,
        node_n_hidden_dimensions: int,
        node_hidden_dimensions_size: int,
        coordinate_n_hidden_dimensions: int,
        coordinate_hidden_dimensions_size: int,
        act_fn: Callable = nn.SiLU(),
        residual: bool = True,
        attention: bool = False,
        normalize: bool = False,
        coords_agg: str = "mean",
        message_agg: str = "mean",
        tanh: bool = False,
    ):
        super().__init__()
        if coords_agg not in ["mean", "sum"]:
            raise ValueError(f"coords_agg should be mean or sum. Got {coords_agg}")
        if message_agg not in ["mean", "sum"]:
            raise ValueError(f"message_agg should be mean or sum. Got {message_agg}")
        self.residual, self.attention, self.normalize, self.tanh = residual, attention, normalize, tanh
        self.coords_agg, self.message_agg = coords_agg, message_agg
        self.epsilon = 1e-8
        m, h, c = message_hidden_dimensions_size, node_hidden_dimensions_size, coordinate_hidden_dimensions_size
        # phi_e: [h_i, h_j, |x_i - x_j|^2] -> m_ij, activation after every Linear
        self.message_mlp = _mlp([2 * input_size + 1] + [m] * (1 + message_n_hidden_dimensions), act_fn, True)
        # phi_h: [h_i, sum_j m_ij] -> h_i update, no activation after the last Linear
        self.node_mlp = _mlp([input_size + m] + [h] * (1 + node_n_hidden_dimensions) + [output_size], act_fn, False)
        # phi_x: m_ij -> scalar weight of (x_i - x_j); last Linear has no bias
        self.coord_mlp = _mlp([m] + [c] * (1 + coordinate_n_hidden_dimensions) + [1], act_fn, False, final_bias=False)
        if tanh:
            self.coord_mlp.append(nn.Tanh())
        if attention:
            self.att_mlp = nn.Sequential(nn.Linear(m, 1), nn.Sigmoid())

    def forward(self, h, edge_index, coord):
        raise B200MDError(_FUSED_ONLY)


class EGNN(nn.Module):
    """Stack of E_GCL layers between an input embedding and a classification read-out (:291-385)."""

    def __init__(
        self,
        input_size: int,
        num_classes: int,
        message_n_hidden_dimensions: int,
        message_hidden_dimensions_size: int,
        node_n_hidden_dimensions: int,
        node_hidden_dimensions_size: int,
        coordinate_n_hidden_dimensions: int,
        coordinate_hidden_dimensions_size: int,
        act_fn: Callable = nn.SiLU(),
        residual: bool = True,
        attention: bool = False,
        normalize: bool = False,
        tanh: bool = False,
        coords_agg: str = "mean",
        message_agg: str = "mean",
        n_layers: int = 4,
    ):
        super().__init__()
        self.n_layers = n_layers
        h = node_hidden_dimensions_size
        self.embedding_in = nn.Linear(input_size, h)
        self.graph_layers = nn.ModuleList([])
        self.node_classification_layer = nn.Linear(h, num_classes)
        for _ in range(n_layers):
            self.graph_layers.append(
                E_GCL(
                    input_size=h,
                    output_size=h,
                    message_n_hidden_dimensions=message_n_hidden_dimensions,
                    message_hidden_dimensions_size=message_hidden_dimensions_size,
                    node_n_hidden_dimensions=node_n_hidden_dimensions,
                    node_hidden_dimensions_size=h,
                    coordinate_n_hidden_dimensions=coordinate_n_hidden_dimensions,
                    coordinate_hidden_dimensions_size=coordinate_hidden_dimensions_size,
                    act_fn=act_fn,
                    residual=residual,
                    attention=attention,
                    normalize=normalize,
                    coords_agg=coords_agg,
                    message_agg=message_agg,
                    tanh=tanh,
                )
            )

    def forward(self, h, edges, x):
        raise B200MDError(_FUSED_ONLY)
