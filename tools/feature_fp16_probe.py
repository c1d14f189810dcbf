"""CPU experiment for DESIGN section 9 item 2 (16-bit feature storage for the plane sweep): how much end-to-end error would
rounding the half-resolution FPN features to fp16 / bf16 -- only where the sweep reads them -- cost?  Runs the CPU oracle
(the reference restatement) twice, with and without the rounding injected in front of cost_volume_fusion, on (a) BASELINE c2
with seeded synthetic weights, 3 recurrent keyframes, and (b) the shipped fusionnet weights on the fixture scene, first 4
keyframes vs the reference's shipped golden.  Development tool: no GPU, nothing of the product path involved."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import numpy as np
import torch

import synth_data as synth
from oracle import dvmvs_oracle as oracle
from tests import scene_fixture

T = torch.from_numpy
_orig = oracle.cost_volume_fusion


def rounded(dtype, which):
    def fusion(image1, image2s, *a, **k):
        r = lambda t: t.to(dtype).to(torch.float32)
        return _orig(r(image1) if which in ("both", "ref") else image1, [r(t) for t in image2s] if which in ("both", "meas") else image2s, *a, **k)
    return fusion


def run_c2(weights, n=3):
    H, W, D, M = 256, 256, 64, 2
    clip = synth.make_clip(0, n, H, W, M)
    K = T(clip["K"])[None]
    st, out = oracle.FusionnetState(), []
    with torch.no_grad():
        for ref_i, meas_i in clip["frames"]:
            g, st = oracle.fusionnet_step(weights, st, T(clip["images"][ref_i])[None], T(clip["poses"][ref_i])[None],
                                          [T(clip["images"][j])[None] for j in meas_i], [T(clip["poses"][j])[None] for j in meas_i], K, n_depth_levels=D)
            out.append(g.numpy())
    return out


def run_scene(weights, n=4):
    frames, full_K, gold = scene_fixture.load_scene()
    st, out = oracle.FusionnetState(), []
    with torch.no_grad():
        for fr in frames[:n]:
            g, st = oracle.fusionnet_step(weights, st, T(fr["reference_image"])[None], T(fr["reference_pose"])[None],
                                          [T(x)[None] for x in fr["measurement_images"]], [T(p)[None] for p in fr["measurement_poses"]],
                                          T(full_K)[None])
            out.append(g[0].numpy())
    return out, gold[:n]


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    shapes = oracle.state_dict_shapes(64)
    w = {tag: {k: T(v) for k, v in synth.make_state_dict(shapes[tag], seed=7).items()} for tag in shapes}
    shipped = scene_fixture.load_shipped_weights("fusionnet")
    base_c2 = run_c2(w)
    base_scene = run_scene(shipped) if shipped is not None else None
    for name, dtype, which in (("fp16 both", torch.float16, "both"), ("fp16 measurement only", torch.float16, "meas"),
                               ("bf16 both", torch.bfloat16, "both")):
        oracle.cost_volume_fusion = rounded(dtype, which)
        try:
            res = {"features": name,
                   "c2_synthetic_vs_unrounded": [float(oracle.rel_l1_inverse_depth(a, b)) for a, b in zip(run_c2(w), base_c2)]}
            if base_scene is not None:
                got, gold = run_scene(shipped)
                res["shipped_vs_unrounded"] = [float(oracle.rel_l1_inverse_depth(a, b)) for a, b in zip(got, base_scene[0])]
                res["shipped_vs_shipped_golden"] = [float(oracle.rel_l1_inverse_depth(a, b)) for a, b in zip(got, gold)]
        finally:
            oracle.cost_volume_fusion = _orig
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
