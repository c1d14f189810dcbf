"""Operand-precision policies of the tensor-core path, measured end to end: for each policy (module family -> 1 | 3
terms) the rel-L1 error on inverse depth of (a) BASELINE config c2 with seeded synthetic weights over 3 recurrent
keyframes vs the CPU oracle and (b) the reference's shipped fusionnet weights on the fixture scene vs the reference's
shipped golden predictions (10 keyframes).  Budget: 1e-3 (BASELINE.json).  Prints one JSON line per policy."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import numpy as np
import torch

import synth_data as synth
from dvmvs import _ops as ops
from dvmvs import pipeline
from oracle import dvmvs_oracle as oracle
from tests import helpers, scene_fixture

H, W, D, M = 256, 256, 64, 2
T = torch.from_numpy
POLICIES = ["", "fe=1,fpn=1", "cve=1", "fe=1,fpn=1,cve=1", "fe=1,fpn=1,cve=1,lstm=1", "cvd=1", "fe=1,fpn=1,cve=1,lstm=1,cvd=1"]


def main():
    policies = sys.argv[1:] or POLICIES
    shapes = oracle.state_dict_shapes(D)
    w = {tag: {k: T(v) for k, v in synth.make_state_dict(shapes[tag], seed=7).items()} for tag in shapes}
    clip = synth.make_clip(0, 3, H, W, M)
    K = T(clip["K"])[None]
    st_o, golds = oracle.FusionnetState(), []
    with torch.no_grad():
        for ref_i, meas_i in clip["frames"]:
            g, st_o = oracle.fusionnet_step(w, st_o, T(clip["images"][ref_i])[None], T(clip["poses"][ref_i])[None],
                                            [T(clip["images"][j])[None] for j in meas_i], [T(clip["poses"][j])[None] for j in meas_i], K,
                                            n_depth_levels=D)
            golds.append(g.numpy())
    shipped = scene_fixture.load_shipped_weights("fusionnet")
    scene = scene_fixture.load_scene() if shipped is not None else None
    ops.set_conv_backend("tc", terms=3, stride2=True)
    for pol in policies:
        ops.set_precision_policy(pol)
        res = {"policy": pol or "all=3"}
        with torch.no_grad():
            mods = pipeline.build_modules(w, device="cuda", n_depth_levels=D)
            st, errs = pipeline.KeyframeState(), []
            for fi, (ref_i, meas_i) in enumerate(clip["frames"]):
                pred, st = pipeline.keyframe(mods, st, T(clip["images"][ref_i])[None].cuda(), T(clip["poses"][ref_i])[None].cuda(),
                                             [T(clip["images"][j])[None].cuda() for j in meas_i],
                                             [T(clip["poses"][j])[None].cuda() for j in meas_i], K.cuda(), n_depth_levels=D)
                errs.append(float(oracle.rel_l1_inverse_depth(pred.cpu().numpy(), golds[fi])))
            res["c2_synthetic_vs_oracle"] = errs
            if shipped is not None:
                mods = helpers.build_product_modules(shipped)
                frames, full_K, gold = scene
                state, errs = helpers.ProductState(), []
                cu = lambda a: T(np.ascontiguousarray(a)).cuda()
                for i, fr in enumerate(frames):
                    pred, state = helpers.product_fusionnet_step(mods, state, cu(fr["reference_image"])[None], cu(fr["reference_pose"])[None],
                                                                 [cu(x)[None] for x in fr["measurement_images"]],
                                                                 [cu(p)[None] for p in fr["measurement_poses"]], cu(full_K)[None])
                    errs.append(float(oracle.rel_l1_inverse_depth(pred[0].cpu().numpy(), gold[i])))
                res["shipped_weights_vs_shipped_golden"] = errs
        res["max"] = max(res["c2_synthetic_vs_oracle"] + res.get("shipped_weights_vs_shipped_golden", []))
        print(json.dumps(res), flush=True)
    ops.set_precision_policy(None)


if __name__ == "__main__":
    main()
