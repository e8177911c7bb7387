"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference package from /root/reference.

The reference (mila-iqia/diffusion_for_multi_scale_molecular_dynamics) is pure Python but needs
four third-party modules that are not installed in this image.  This module injects minimal
stand-ins into ``sys.modules`` *before* the reference is imported so that the reference's own
source files run unchanged on CPU.  It is used by ``oracle/make_golden.py`` (fixture generation)
and by CPU tests that are skipped when /root/reference is absent (it is absent on the GPU box).

Shims:
  * ``e3nn.o3.Irreps`` / ``torch_geometric.data.Data``: only imported for annotations in
    models/mace_utils.py:6-7; never executed on the EGNN path.
  * ``pykeops.torch.LazyTensor``: dense torch stand-in for the four symbolic ops used at
    utils/neighbors.py:155-186 (``-``, ``** 2``, ``.sum(dim)``, ``<=``, ``.sum_reduction``,
    ``.Kmin_argKmin``).  pykeops 2.2.3 is the pinned third-party dependency; its source is not in
    /root/reference, so exact tie behaviour at d^2 == rc^2 is parity-unpinned by the reference.
  * ``numpy.NaN`` alias (utils/neighbors.py:21; removed in numpy 2).
  * ``pymatgen.core.Lattice`` / ``Structure``: imported at utils/structure_utils.py:6 for ``create_structure``,
    which nothing on this path calls.

Nothing in the product package imports this file.
"""
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("B200MD_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")
REF_PKG = "diffusion_for_multi_scale_molecular_dynamics"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, REF_PKG))


class _DenseLazyTensor:
    """Dense stand-in for pykeops.torch.LazyTensor restricted to what neighbors.py uses."""

    def __init__(self, t):
        self.t = t

    def __sub__(self, other):
        return _DenseLazyTensor(self.t - other.t)

    def __pow__(self, p):
        return _DenseLazyTensor(self.t ** p)

    def sum(self, dim):
        return _DenseLazyTensor(self.t.sum(dim=dim))

    def __le__(self, other):
        other = other.t if isinstance(other, _DenseLazyTensor) else other
        return _DenseLazyTensor((self.t <= other).to(torch.float32))

    def sum_reduction(self, dim):
        return self.t.sum(dim=dim, keepdim=True)

    def Kmin_argKmin(self, K, dim):
        values, indices = torch.topk(self.t, k=K, dim=dim, largest=False, sorted=True)
        return values, indices


def install_shims() -> None:
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    if "e3nn" not in sys.modules:
        e3nn = types.ModuleType("e3nn")
        o3 = types.ModuleType("e3nn.o3")
        o3.Irreps = object
        e3nn.o3 = o3
        sys.modules["e3nn"] = e3nn
        sys.modules["e3nn.o3"] = o3
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgd = types.ModuleType("torch_geometric.data")
        tgd.Data = object
        tg.data = tgd
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.data"] = tgd
    if "pykeops" not in sys.modules:
        pk = types.ModuleType("pykeops")
        pkt = types.ModuleType("pykeops.torch")
        pkt.LazyTensor = _DenseLazyTensor
        pk.torch = pkt
        sys.modules["pykeops"] = pk
        sys.modules["pykeops.torch"] = pkt
    if "pymatgen" not in sys.modules:  # utils/structure_utils.py:6 (create_structure only; never called here)
        pm = types.ModuleType("pymatgen")
        pmc = types.ModuleType("pymatgen.core")
        pmc.Lattice = object
        pmc.Structure = object
        pm.core = pmc
        sys.modules["pymatgen"] = pm
        sys.modules["pymatgen.core"] = pmc
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


def import_reference():
    """Return the reference top-level package (raises if /root/reference is absent)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    install_shims()
    import importlib

    return importlib.import_module(REF_PKG)
