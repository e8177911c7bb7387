"""EGNN hyper-parameter variants shared by the golden generator (oracle/make_golden.py) and the tests."""
EGNN_VARIANTS = {
    "default_fc": dict(num_atom_types=1),
    "radial_sige": dict(num_atom_types=2, edges="radial_cutoff", radial_cutoff=3.0),
    "all_options": dict(num_atom_types=3, edges="radial_cutoff", radial_cutoff=2.5, number_of_bloch_wave_shells=2,
                        message_n_hidden_dimensions=2, message_hidden_dimensions_size=24,
                        node_n_hidden_dimensions=2, node_hidden_dimensions_size=40,
                        coordinate_n_hidden_dimensions=0, coordinate_hidden_dimensions_size=20, residual=False,
                        attention=True, normalize=True, tanh=True, coords_agg="sum", message_agg="sum", n_layers=2),
    "wide": dict(num_atom_types=1, edges="radial_cutoff", radial_cutoff=3.2, message_n_hidden_dimensions=2,
                 message_hidden_dimensions_size=64, node_n_hidden_dimensions=2, node_hidden_dimensions_size=64,
                 coordinate_n_hidden_dimensions=2, coordinate_hidden_dimensions_size=64, n_layers=3),
}

# dataclass defaults of EGNNScoreNetworkParameters (reference egnn_score_network.py:24-45)
EGNN_DEFAULTS = dict(number_of_bloch_wave_shells=1, message_n_hidden_dimensions=1, message_hidden_dimensions_size=16,
                     node_n_hidden_dimensions=1, node_hidden_dimensions_size=32, coordinate_n_hidden_dimensions=1,
                     coordinate_hidden_dimensions_size=32, residual=True, attention=False, normalize=False, tanh=False,
                     coords_agg="mean", message_agg="mean", n_layers=4, edges="fully_connected", radial_cutoff=None)


def hyper_parameters(variant: str) -> dict:
    hp = dict(EGNN_DEFAULTS)
    hp.update(EGNN_VARIANTS[variant])
    return hp
