"""Element <-> integer index bookkeeping (reference: data/element_types.py:7-90)."""
from typing import Dict, List

NULL_ELEMENT = "NULL_ELEMENT_FOR_PADDING"
NULL_ELEMENT_ID = -1


class ElementTypes:
    """Sorted element names and their indices; the index of an element is its rank in alphabetical order."""

    def __init__(self, elements: List[str]):
        self.validate_elements(elements)
        self._elements = sorted(elements)
        self._ids = list(range(len(self._elements)))
        self._element_to_id_map: Dict[str, int] = dict(zip(self._elements, self._ids))
        self._id_to_element_map: Dict[int, str] = dict(zip(self._ids, self._elements))
        self._element_to_id_map[NULL_ELEMENT] = NULL_ELEMENT_ID
        self._id_to_element_map[NULL_ELEMENT_ID] = NULL_ELEMENT

    @staticmethod
    def validate_elements(elements: List[str]):
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
