"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference package from /root/reference inside THIS
build container (it does not exist on the GPU box) so golden vectors can be generated from the reference's
own code.  Puts the three dependency stand-ins of oracle/shims on sys.path (kornia 0.3.2, path 15.0.0,
pytorch3d 0.2.5 are not installed; README.md:56-71) and stops FeatureExtractor.__init__
(dvmvs/fusionnet/model.py:125) from downloading ImageNet weights.  Must run in its own process: the
reference package is also called `dvmvs`, like the product's drop-in package.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("DVMVS_REFERENCE_ROOT", "/root/reference")


def import_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "dvmvs")):
        raise RuntimeError("reference tree not found at %s (only present in the build container)" % REFERENCE_ROOT)
    for p in (os.path.join(here, "shims"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torchvision.models as tvm
    if not getattr(tvm.mnasnet1_0, "_dvmvs_no_download", False):
        orig = tvm.mnasnet1_0

        def mnasnet1_0(pretrained=False, **kw):          # same architecture, never touches the network
            return orig(weights=None, **kw)

        mnasnet1_0._dvmvs_no_download = True
        tvm.mnasnet1_0 = mnasnet1_0
    import dvmvs  # noqa: F401  (the reference)
    assert os.path.abspath(dvmvs.__path__[0]).startswith(os.path.abspath(REFERENCE_ROOT)), dvmvs.__path__
    return dvmvs
