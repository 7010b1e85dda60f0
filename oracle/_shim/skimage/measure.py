"""``skimage.measure`` stand-in: forwards ``regionprops`` to the oracle's
restatement of scikit-image 0.15 (oracle/reference_path.py)."""
from oracle.reference_path import regionprops  # noqa: F401
