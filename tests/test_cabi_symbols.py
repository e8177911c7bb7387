"""CPU: the C-ABI library loads and exports every entry point ``include/b200md.h`` declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200md.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200md_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_bound_signatures():
    assert set(declared_symbols()) == set(_cabi.SIGNATURES), \
        "include/b200md.h and _cabi.SIGNATURES must list the same entry points"


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_cabi.LIBRARY_PATH):
        import __graft_entry__
        __graft_entry__.build()
    library = _cabi.load_library()
    for name in declared_symbols():
        assert hasattr(library, name), f"{name} is declared in include/b200md.h but not exported"
    assert library.b200md_abi_version() == 1
    assert library.b200md_error_string(0) == b"ok"
    assert library.b200md_launch_count() == 0


def test_struct_layouts_match_the_header():
    # 11 int32 + 2 float
    assert ctypes.sizeof(_cabi.EgnnConfig) == 13 * 4
    # 6 int32 + 2 float + 8 pointers + (int32 + pad) + 3 pointers
    assert ctypes.sizeof(_cabi.PcSchedule) == 8 * 4 + 8 * 8 + 8 + 3 * 8


def test_compute_entry_points_refuse_cpu_tensors():
    import torch
    from diffusion_for_multi_scale_molecular_dynamics_b200.utils.basis_transformations import \
        map_relative_coordinates_to_unit_cell
    with pytest.raises(_cabi.B200MDError):
        map_relative_coordinates_to_unit_cell(torch.rand(4, 3))


def test_missing_library_is_a_loud_error(tmp_path):
    with pytest.raises(_cabi.B200MDError):
        _cabi.load_library(str(tmp_path / "nope.so"))
