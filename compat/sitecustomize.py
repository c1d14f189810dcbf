"""Environment compatibility overlay for running the reference's OWN scripts (dvmvs/fusionnet/run-testing.py, ...) unmodified
against this package on the software of this image.  Put this directory on PYTHONPATH *before* anything else:

    PYTHONPATH=<repo>/compat:<repo>/deep-video-mvs_b200  DVMVS_DATA=<sample-data>  DVMVS_RESULTS=<out dir> \
        python <reference>/dvmvs/fusionnet/run-testing.py          # cwd: a directory holding `weights/`

It touches nothing of the reference and nothing of the model path; it only bridges two incompatibilities between the
reference's pinned 2020 environment (README.md:56-71: numpy 1.18, path==15.0.0) and this one:

  * `np.loadtxt(file, dtype=str, delimiter="\n")` (run-testing.py:73, pairnet/run-testing.py:69, dataset_loader.py:256) raises
    "control character 'delimiter' cannot be a newline" under NumPy >= 1.23: answered with the lines of the file, which is
    what NumPy 1.18 returned;
  * `from path import Path` (run-testing.py:9): the `path` package is not installed and there is no network -- compat/path.py
    next to this file provides the handful of methods the scripts use.

Everything else the scripts need from their environment is configuration, and `dvmvs.config.Config` of this package reads it
from environment variables (DVMVS_DATA, DVMVS_RESULTS, DVMVS_SCENE; Config.test_visualize defaults to False: no display)."""
import numpy as _np

_loadtxt = _np.loadtxt


def _loadtxt_compat(fname, *args, **kwargs):
    if kwargs.get("delimiter") == "\n":
        dtype = kwargs.get("dtype", args[0] if args else float)
        with open(str(fname)) as fh:
            lines = [ln for ln in fh.read().splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
        return _np.atleast_1d(_np.array(lines, dtype=dtype))
    return _loadtxt(fname, *args, **kwargs)


_loadtxt_compat.__doc__ = _loadtxt.__doc__
_np.loadtxt = _loadtxt_compat
