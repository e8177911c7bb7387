"""Recorder for sampling trajectories (reference: utils/sample_trajectory.py:7-44)."""
from collections import defaultdict
from typing import Any, Dict

import torch


class SampleTrajectory:
    """Append-only store of per-step records kept on the host; ``write_to_pickle`` uses torch.save."""

    def __init__(self):
        self._internal_data = defaultdict(list)

    def reset(self):
        self._internal_data = defaultdict(list)

    def record(self, key: str, entry: Any):
        self._internal_data[key].append(entry)

    def write_to_pickle(self, path_to_pickle: str):
        with open(path_to_pickle, "wb") as fd:
            torch.save(self._internal_data, fd)

    @property
    def data(self) -> Dict[str, list]:
        return self._internal_data
