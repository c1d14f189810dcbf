"""TEST INFRASTRUCTURE ONLY -- copies what the keyframe-selection test needs out of the reference's sample data (build
container only; /root/reference is absent on the GPU box):

    python oracle/make_golden_keyframes.py

  tests/golden/keyframes/poses_000.npy         all 373 camera poses of fixture scene 000 (float64, as np.fromfile reads
                                               sample-data/hololens-dataset/000/poses.txt, simulate_keyframe_buffer.py:28)
  tests/golden/keyframes/image_names_000.txt   the scene's sorted image file names (index -> name)
  tests/golden/keyframes/keyframe+hololens-dataset+000+nmeas+{1,2,3}
                                               the reference's SHIPPED selection results (sample-data/indices/), produced by
                                               its own KeyframeBuffer through simulate_keyframe_buffer.py -- the golden
"""
import os
import shutil

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("DVMVS_REFERENCE_ROOT", "/root/reference")


def main():
    scene = os.path.join(REFERENCE_ROOT, "sample-data", "hololens-dataset", "000")
    out = os.path.join(REPO, "tests", "golden", "keyframes")
    os.makedirs(out, exist_ok=True)
    poses = np.fromfile(os.path.join(scene, "poses.txt"), dtype=float, sep="\n ").reshape((-1, 4, 4))
    np.save(os.path.join(out, "poses_000.npy"), poses)
    names = sorted(n for n in os.listdir(os.path.join(scene, "images")) if n.endswith(".png"))
    assert len(names) == len(poses)
    with open(os.path.join(out, "image_names_000.txt"), "w") as fh:
        fh.write("\n".join(names) + "\n")
    for n in (1, 2, 3):
        f = "keyframe+hololens-dataset+000+nmeas+%d" % n
        shutil.copyfile(os.path.join(REFERENCE_ROOT, "sample-data", "indices", f), os.path.join(out, f))
    print("wrote", out, poses.shape)


if __name__ == "__main__":
    main()
