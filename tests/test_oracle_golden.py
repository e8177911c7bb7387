"""CPU: pin ``oracle/pc_oracle.py`` against the golden vectors produced by the unmodified reference
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``).  Integer outputs must be bit-exact; floating point within
the tolerance written beside each check."""
import pytest
import torch

from oracle import pc_oracle as oracle
from tests.conftest import load_golden, sub_dict
from tests.egnn_variants import hyper_parameters

# fp32 tolerance for the oracle against the reference: same torch ops on the same ISA, so only the order of a few
# reductions differs.
RTOL, ATOL = 2e-6, 2e-6


@pytest.mark.parametrize("tag,kwargs,C", [
    ("exp_T10_C2", dict(total_time_steps=10), 2),
    ("lin_T7_C3", dict(total_time_steps=7, schedule_type="linear", sigma_min=1e-4, sigma_max=0.2,
                       corrector_step_epsilon=2.5e-8), 3),
    ("exp_T1000_C2", dict(total_time_steps=1000), 2),
])
def test_noise_tables(tag, kwargs, C):
    golden = sub_dict(load_golden("noise_tables"), tag)
    tables = oracle.noise_tables(num_classes=C, **kwargs)
    for key, expected in golden.items():
        torch.testing.assert_close(tables[key], expected, rtol=1e-6, atol=1e-12, msg=key)


def test_wrap():
    g = load_golden("wrap_and_wrapped_gaussian")
    out = oracle.wrap_to_unit_cell(g["wrap_in"].clone())
    assert torch.equal(out, g["wrap_out"])
    assert (out >= 0).all() and (out < 1).all()


@pytest.mark.parametrize("tag,rtol", [("f32", 2e-5), ("f64", 1e-12)])
@pytest.mark.parametrize("kmax", [1, 4])
def test_wrapped_gaussian_score(tag, rtol, kmax):
    g = load_golden("wrap_and_wrapped_gaussian")
    out = oracle.wrapped_gaussian_score(g[f"wg_{tag}_u"], g[f"wg_{tag}_sigma"], kmax)
    torch.testing.assert_close(out, g[f"wg_{tag}_k{kmax}"], rtol=rtol, atol=rtol)


@pytest.mark.parametrize("rc", [1.1, 2.2])
def test_periodic_adjacency(rc):
    g = load_golden("neighbors")
    adjacency, shifts, edge_batch, counts = oracle.periodic_adjacency(g["tri.cart"], g["tri.basis"], rc)
    assert torch.equal(counts, g[f"tri_rc{rc}.number_of_edges"])
    assert torch.equal(edge_batch, g[f"tri_rc{rc}.edge_batch"])
    assert torch.equal(adjacency, g[f"tri_rc{rc}.adjacency"])
    torch.testing.assert_close(shifts, g[f"tri_rc{rc}.shifts"], rtol=1e-6, atol=1e-6)


def test_cutoff_too_large_raises():
    g = load_golden("neighbors")
    with pytest.raises(AssertionError):
        oracle.periodic_adjacency(g["tri.cart"], g["tri.basis"], 6.0)


@pytest.mark.parametrize("rc", ["3.0", "3.3"])
def test_edges_with_radial_cutoff(rc):
    g = load_golden("neighbors")
    edges = oracle.edges_with_radial_cutoff(g["egnn.rel"], g["egnn.cell"], float(rc))
    assert torch.equal(edges, g[f"egnn.edges_rc{rc}"])


def test_fully_connected_edges():
    edges = oracle.fully_connected_edges(4, 2)
    assert edges.shape == (24, 2)
    assert edges[0].tolist() == [0, 1] and edges[12].tolist() == [4, 5]


def test_segment_reductions():
    g = load_golden("segment")
    torch.testing.assert_close(oracle.segment_sum(g["data"], g["ids"], 11), g["sum"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(oracle.segment_mean(g["data"], g["ids"], 11), g["mean"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("variant", ["default_fc", "radial_sige", "all_options", "wide"])
def test_egnn_forward(variant):
    g = sub_dict(load_golden("egnn_forward"), variant)
    logits, score = oracle.egnn_score_forward(sub_dict(g, "sd"), hyper_parameters(variant), g["a"], g["x"], g["l"],
                                              g["sigma"])
    torch.testing.assert_close(logits, g["logits"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(score, g["score"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("idx", [0, 3, 7])
def test_atom_types_update(idx):
    g = load_golden("atom_types")
    q, qb, qbm = g[f"idx{idx}.q"], g[f"idx{idx}.q_bar"], g[f"idx{idx}.q_bar_tm1"]
    probs = oracle.probability_at_previous_time_step(g["logits"], g["a"], q, qb, qbm, 1e-8)
    torch.testing.assert_close(probs, g[f"idx{idx}.probs"], rtol=1e-6, atol=1e-7, equal_nan=True)
    for greedy in (False, True):
        for one in (False, True):
            new, _ = oracle.atom_types_update(g["logits"], g["a"], q, qb, qbm, g["gumbel"], g["binary"], greedy, one,
                                              1e-8)
            assert torch.equal(new, g[f"idx{idx}.g{int(greedy)}o{int(one)}"]), (greedy, one)


TRAJECTORY_CASES = {
    "si_fc": ("default_fc", dict()),
    "sige_radial": ("radial_sige", dict()),
    "sige_free": ("radial_sige", dict(one_transition=False, greedy=False, atom_type_transition_in_corrector=True)),
    "repaint": ("radial_sige", dict()),
}


@pytest.mark.parametrize("name", list(TRAJECTORY_CASES))
def test_trajectory(name):
    variant, flags = TRAJECTORY_CASES[name]
    g = sub_dict(load_golden("trajectories"), name)
    hp = hyper_parameters(variant)
    sd = sub_dict(g, "sd")
    T, n_corr = int(g["T"]), int(g["n_corr"])
    tables = oracle.noise_tables(T, hp["num_atom_types"] + 1)

    def score_fn(a, x, lattice, sigma):
        return oracle.egnn_score_forward(sd, hp, a, x, lattice, torch.full((x.shape[0], 1), sigma))

    constraint = None
    if "c_x0" in g:
        constraint = dict(x0=g["c_x0"], a0=g["c_a0"], indices=g["c_idx"])
    a, x, _ = oracle.pc_sample(score_fn, tables, g["a0"], g["x0"], g["l"], oracle.NoiseTape(77),
                               number_of_corrector_steps=n_corr, constraint=constraint, **flags)
    assert torch.equal(a, g["final_a"])
    # trajectories amplify rounding differences; compare on the torus
    delta = (x - g["final_x"]).abs()
    delta = torch.minimum(delta, 1.0 - delta)
    assert delta.max() < 2e-5, delta.max()


def test_analytical_score_network():
    g = load_golden("analytical")
    score = oracle.analytical_score(g["x"], g["x0"], g["sigma"], 0.01, 4)
    torch.testing.assert_close(score, g["score"], rtol=2e-5, atol=2e-5)


def test_analytical_trajectory():
    """A whole trajectory driven by the analytical score: contractive dynamics, so coordinates agree tightly."""
    g = load_golden("analytical")
    T = int(g["T"])
    tables = oracle.noise_tables(T, 2)
    B = g["traj_x0"].shape[0]
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = 5.43

    def score_fn(a, x, lat, sigma):
        logits = torch.zeros(B, 8, 2)
        logits[..., -1] = -torch.inf
        return logits, oracle.analytical_score(x, g["x0"], torch.full((B,), sigma), 0.01, 4)

    a, x, _ = oracle.pc_sample(score_fn, tables, torch.ones(B, 8, dtype=torch.long), g["traj_x0"], lattice,
                               oracle.NoiseTape(31))
    assert torch.equal(a, g["traj_final_a"])
    delta = (x - g["traj_final_x"]).abs()
    assert torch.minimum(delta, 1.0 - delta).max() < 2e-5
    # the samples sit within a few sigma_d of the equilibrium positions
    d0 = (x - g["x0"]).abs()
    assert torch.minimum(d0, 1.0 - d0).max() < 0.08


def test_free_lattice_trajectory():
    """use_fixed_lattice_parameters=False: the oracle's lattice diffusion against the reference's trajectory."""
    g = load_golden("free_lattice")
    sd = sub_dict(g, "sd")
    hp = hyper_parameters("radial_sige")
    T, n_corr = int(g["T"]), int(g["n_corr"])
    tables = oracle.noise_tables(T, 3, sigma_min=0.005, sigma_max=0.5)
    B = g["x0"].shape[0]

    def score_fn(at, x, lat, sigma):
        return oracle.egnn_score_forward(sd, hp, at, x, lat, torch.full((B, 1), sigma))

    a, x, lattice = oracle.pc_sample(score_fn, tables, g["a0"], g["x0"], g["l"], oracle.NoiseTape(80),
                                     number_of_corrector_steps=n_corr, free_lattice=True)
    assert torch.equal(a, g["final_a"])
    torch.testing.assert_close(lattice, g["final_l"], rtol=1e-6, atol=1e-6)
    d = (x - g["final_x"]).abs()
    assert torch.minimum(d, 1 - d).max() < 2e-5
