"""Two-sample Kolmogorov-Smirnov metric (reference: metrics/kolmogorov_smirnov_metrics.py:7-75).

Same interface as the reference class.  The samples (pair distances from ``compute_distances_in_batch``, energies) are
kept as a list of tensors and concatenated at compute time (the reference uses ``torchmetrics.CatMetric`` for that,
which is not a dependency of the sampling path); the statistic itself is scipy's ``ks_2samp``, as in the reference.
"""
from typing import List, Tuple

import scipy.stats as ss
import torch


class KolmogorovSmirnovMetrics:
    """Kolmogorov Smirnov metrics."""

    def __init__(self, maximum_number_of_samples: int = 1_000_000):
        """Init method.

        Args:
            maximum_number_of_samples : maximum number of samples that will be aggregated. This is to avoid
                memory use explosion.
        """
        self._reference_samples: List[torch.Tensor] = []
        self._predicted_samples: List[torch.Tensor] = []
        self.maximum_count = maximum_number_of_samples
        self.reference_count = 0
        self.predicted_count = 0

    def register_reference_samples(self, reference_samples):
        """Register reference samples."""
        if self.reference_count < self.maximum_count:
            self.reference_count += len(reference_samples)
            self._reference_samples.append(torch.as_tensor(reference_samples).detach().flatten())

    def register_predicted_samples(self, predicted_samples):
        """Register predicted samples."""
        if self.predicted_count < self.maximum_count:
            self.predicted_count += len(predicted_samples)
            self._predicted_samples.append(torch.as_tensor(predicted_samples).detach().flatten())

    def reset(self):
        """reset."""
        self._reference_samples = []
        self._predicted_samples = []
        self.reference_count = 0
        self.predicted_count = 0

    def compute_kolmogorov_smirnov_distance_and_pvalue(self) -> Tuple[float, float]:
        """Two-sample, two-sided KS test between the predicted and the reference samples.

        Returns:
            ks_distance, p_value: the Kolmogorov-Smirnov test statistic (a "distance")
                and the statistical test's p-value.
        """
        reference_samples = torch.cat(self._reference_samples).cpu().numpy()
        predicted_samples = torch.cat(self._predicted_samples).cpu().numpy()
        test_result = ss.ks_2samp(predicted_samples, reference_samples, alternative="two-sided", method="auto")
        return test_result.statistic, test_result.pvalue
