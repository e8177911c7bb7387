"""Energy oracles (reference: oracle/__init__.py): where the Stillinger-Weber coefficient files live."""
from pathlib import Path

SW_COEFFICIENTS_DIR = Path(__file__).resolve().parent.parent / "data" / "stillinger_weber_coefficients"
