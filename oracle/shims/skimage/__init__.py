"""TEST INFRASTRUCTURE ONLY -- placeholder for scikit-image (not installed, no network): the reference's
sample-data/run-tsdf-reconstruction.py imports `skimage.measure` at module level for marching cubes, which the golden
generator never calls."""
from . import measure  # noqa: F401
