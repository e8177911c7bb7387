"""Chemical element names <-> integer atom-type indices (API mirror of the reference's data/element_types.py:7-90).

The sampling path only needs the mapping in two places: ``SamplingConstraint`` (element names of the constrained
atoms) and the Stillinger-Weber oracle (column order of the ``.sw`` file).  The contract that matters is the ORDER:
atom type k is the k-th element name in alphabetical order, and the padding element maps to -1.
"""
from typing import Dict, List

NULL_ELEMENT = "NULL_ELEMENT_FOR_PADDING"
NULL_ELEMENT_ID = -1


class ElementTypes:
    """Alphabetically ranked element names: ``get_element_id("Ge") == 0`` and ``get_element_id("Si") == 1`` for SiGe."""

    def __init__(self, elements: List[str]):
        self.validate_elements(elements)
        ranked = sorted(elements)
        self._elements = ranked
        self._ids = list(range(len(ranked)))
        self._element_to_id_map: Dict[str, int] = {name: rank for rank, name in enumerate(ranked)}
        self._id_to_element_map: Dict[int, str] = {rank: name for rank, name in enumerate(ranked)}
        # the padding element is addressable both ways, like in the reference
        self._element_to_id_map[NULL_ELEMENT] = NULL_ELEMENT_ID
        self._id_to_element_map[NULL_ELEMENT_ID] = NULL_ELEMENT

    @staticmethod
    def validate_elements(elements: List[str]):
        """Reserved name and duplicates are refused with the reference's assertion messages."""
        assert NULL_ELEMENT not in elements, f"The element '{NULL_ELEMENT}' is reserved and should not be used."
        assert len(set(elements)) == len(elements), "Each entry in the elements list should be unique."

    @property
    def number_of_atom_types(self) -> int:
        return len(self._elements)

    @property
    def elements(self) -> List[str]:
        return self._elements

    @property
    def element_ids(self) -> List[int]:
        return self._ids

    def get_element(self, element_id: int) -> str:
        return self._id_to_element_map[element_id]

    def get_element_id(self, element: str) -> int:
        return self._element_to_id_map[element]
