"""Two-sample Kolmogorov-Smirnov metric (reference: metrics/kolmogorov_smirnov_metrics.py:7-75).

Same interface as the reference class.  The samples (pair distances from ``compute_distances_in_batch``, energies) are
kept as a list of tensors and concatenated at compute time (the reference uses ``torchmetrics.CatMetric`` for that,
which is not a dependency of the sampling path); the statistic itself is scipy's ``ks_2samp``, as in the reference.
"""
from typing import List, Tuple

import scipy.stats as ss
import torch


class KolmogorovSmirnovMetrics:
    """Accumulates two families of scalar samples and compares their distributions (reference class of the same name)."""

    def __init__(self, maximum_number_of_samples: int = 1_000_000):
        """``maximum_number_of_samples`` caps each family: once a family holds that many values, further registrations
        for it are dropped (a whole registration is kept or dropped, so the cap can be exceeded by one batch)."""
        self._reference_samples: List[torch.Tensor] = []
        self._predicted_samples: List[torch.Tensor] = []
        self.maximum_count = maximum_number_of_samples
        self.reference_count = 0
        self.predicted_count = 0

    def _register(self, store: List[torch.Tensor], count: int, samples) -> int:
        if count >= self.maximum_count:
            return count
        store.append(torch.as_tensor(samples).detach().flatten())
        return count + len(samples)

    def register_reference_samples(self, reference_samples):
        """Samples of the data distribution (e.g. pair distances or energies of the validation set)."""
        self.reference_count = self._register(self._reference_samples, self.reference_count, reference_samples)

    def register_predicted_samples(self, predicted_samples):
        """Samples of the generated structures."""
        self.predicted_count = self._register(self._predicted_samples, self.predicted_count, predicted_samples)

    def reset(self):
        """Forget everything registered so far."""
        self._reference_samples, self._predicted_samples = [], []
        self.reference_count = self.predicted_count = 0

    def compute_kolmogorov_smirnov_distance_and_pvalue(self) -> Tuple[float, float]:
        """(KS statistic, p-value) of scipy's two-sided two-sample test, predicted vs reference."""
        reference = torch.cat(self._reference_samples).cpu().numpy()
        predicted = torch.cat(self._predicted_samples).cpu().numpy()
        result = ss.ks_2samp(predicted, reference, alternative="two-sided", method="auto")
        return result.statistic, result.pvalue
