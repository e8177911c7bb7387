"""B200-native predictor-corrector sampling path.

A from-scratch sm_100a implementation of ONE hot path of
mila-iqia/diffusion_for_multi_scale_molecular_dynamics: ``LangevinGenerator`` (PC sampler, exploding
variance schedule) driving ``EGNNScoreNetwork`` over batches of periodic supercells in AXL form.  The
Python modules mirror the reference's module layout, class names, dataclass fields and method
signatures for that path only; all arithmetic on the path runs in hand-written CUDA kernels behind
the C ABI declared in ``include/b200md.h`` (``csrc/libb200md.so``).  There is no CPU fallback: a
compute entry point called without the extension or with CPU tensors raises ``B200MDError``.
"""
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL  # noqa: F401

__version__ = "0.1.0"
