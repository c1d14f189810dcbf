"""TEST INFRASTRUCTURE ONLY (build container): golden vectors for SURVEY section 8 row f4 -- the reference's own TSDFVolume
(sample-data/run-tsdf-reconstruction.py:30-310), CPU mode (use_gpu=False: the numba / numpy path; pycuda is not installed),
executed from the UNMODIFIED script file on the seeded cases of oracle/tsdf_cases.py.  Writes tests/golden/tsdf.npz.

    python oracle/make_golden_tsdf.py

The script file is loaded with importlib from /root/reference (its `if __name__ == "__main__"` block does not run); its
module-level imports are served by the reference's own `dvmvs` package + the stand-ins of oracle/shims (path, kornia,
pytorch3d, skimage)."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_runner  # noqa: E402
import tsdf_cases  # noqa: E402


def main():
    ref_runner.import_reference()
    path = os.path.join(ref_runner.REFERENCE_ROOT, "sample-data", "run-tsdf-reconstruction.py")
    spec = importlib.util.spec_from_file_location("ref_tsdf", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for name in tsdf_cases.CASES:
        inp = tsdf_cases.inputs(name)
        vol = mod.TSDFVolume(inp["bounds"].copy(), inp["voxel"], use_gpu=False)
        for i, fr in enumerate(inp["frames"]):
            vol.integrate(fr["color"], fr["depth"], inp["K"], fr["pose"], obs_weight=fr["weight"])
            tsdf, color = vol.get_volume()
            out["%s/tsdf_after_%d" % (name, i)] = tsdf.copy()
            out["%s/color_after_%d" % (name, i)] = color.copy()
            out["%s/weight_after_%d" % (name, i)] = vol._weight_vol_cpu.copy()
        out["%s/frustum_bounds" % name] = mod.TSDFFusion.calculate_volume_bounds([f["depth"] for f in inp["frames"]], [f["pose"] for f in inp["frames"]], inp["K"])
        out["%s/vol_dim" % name] = np.asarray(vol._vol_dim)
        out["%s/vol_origin" % name] = np.asarray(vol._vol_origin)
        print(name, vol._vol_dim, "observed voxels:", int((vol._weight_vol_cpu > 0).sum()))
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "tsdf.npz"), **out)
    print("wrote tests/golden/tsdf.npz")


if __name__ == "__main__":
    main()
