"""Development probe: error of the tensor-core backend against the CPU oracle along a LONG recurrent sequence over real
textures -- the committed fixture keyframes of scene 000 (shipped weights, 320x256, 1..3 measurement frames) played forwards and
backwards (ping-pong: every keyframe keeps its own measurement frames and poses, the recurrent state is carried throughout).

    python tools/drift_probe.py [--frames 60] [--policies "1;3;fe=1,fpn=1,cve=1,lstm=3,cvd=3"]"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "deep-video-mvs_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--policies", default="1;3;fe=1,fpn=1,cve=1,sweep=1,lstm=3,cvd=3;fe=1,fpn=1,cve=1,sweep=1,lstm=1,cvd=3")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from oracle import dvmvs_oracle as oracle
    from tests import helpers, scene_fixture
    from dvmvs import _ops as ops
    w = scene_fixture.load_shipped_weights("fusionnet")
    frames, full_K, _ = scene_fixture.load_scene()
    order = list(range(len(frames)))
    seq = []
    while len(seq) < args.frames:
        seq += order + order[-2:0:-1]
    seq = seq[:args.frames]
    T = torch.from_numpy
    K = T(full_K)[None]
    torch.set_num_threads(16)
    golds, st = [], oracle.FusionnetState()
    with torch.no_grad():
        for i in seq:
            fr = frames[i]
            g, st = oracle.fusionnet_step(w, st, T(fr["reference_image"])[None], T(fr["reference_pose"])[None], [T(x)[None] for x in fr["measurement_images"]],
                                          [T(p)[None] for p in fr["measurement_poses"]], K)
            golds.append(g.numpy())
    results = {}
    for pol in args.policies.split(";"):
        if pol in ("1", "3"):
            ops.set_precision_policy(None)
            ops.set_conv_backend("tc", terms=int(pol), stride2=True)
        else:
            ops.set_conv_backend("tc", terms=3, stride2=True)
            ops.set_precision_policy(pol)
        mods = helpers.build_product_modules(w)
        state = helpers.ProductState()
        errs = []
        c = lambda a: T(np.ascontiguousarray(a))[None].cuda()
        with torch.no_grad():
            for t, i in enumerate(seq):
                fr = frames[i]
                pred, state = helpers.product_fusionnet_step(mods, state, c(fr["reference_image"]), c(fr["reference_pose"]), [c(x) for x in fr["measurement_images"]],
                                                             [c(p) for p in fr["measurement_poses"]], K.cuda())
                errs.append(float(oracle.rel_l1_inverse_depth(pred.cpu().numpy(), golds[t])))
        results[pol] = errs
        print("policy %-50s max %.2e  every 6th: %s" % (pol, max(errs), ["%.1e" % e for e in errs[::6]]), flush=True)
    ops.set_precision_policy(None)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump({"sequence": seq, "rel_l1_inverse_depth_vs_oracle": results}, fh)


if __name__ == "__main__":
    main()
