"""TEST INFRASTRUCTURE ONLY -- see __init__.py."""


def marching_cubes_lewiner(*args, **kwargs):
    raise NotImplementedError("skimage is not installed in this container; the golden generator does not extract meshes")
