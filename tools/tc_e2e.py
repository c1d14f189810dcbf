"""End-to-end parity of the tensor-core backend: fusionnet c2 (256x256, D=64, M=2), 3 recurrent keyframes, synthetic
weights, CUDA path with DVMVS_CONV_BACKEND=tc vs the CPU oracle; also times both backends (eager, per-keyframe)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import torch

import synth_data as synth
from dvmvs import _ops as ops
from dvmvs import pipeline
from oracle import dvmvs_oracle as oracle

H, W, D, M = 256, 256, 64, 2
T = torch.from_numpy


def main():
    terms = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    shapes = oracle.state_dict_shapes(D)
    w = {tag: {k: T(v) for k, v in synth.make_state_dict(shapes[tag], seed=7).items()} for tag in shapes}
    clip = synth.make_clip(0, 3, H, W, M)
    K = T(clip["K"])[None]
    st_o = oracle.FusionnetState()
    golds = []
    with torch.no_grad():
        for ref_i, meas_i in clip["frames"]:
            g, st_o = oracle.fusionnet_step(w, st_o, T(clip["images"][ref_i])[None], T(clip["poses"][ref_i])[None],
                                            [T(clip["images"][j])[None] for j in meas_i], [T(clip["poses"][j])[None] for j in meas_i], K,
                                            n_depth_levels=D)
            golds.append(g)
    for backend in ("fp32", "tc"):
        ops.set_conv_backend(backend, terms=terms, stride2=os.environ.get("DVMVS_TC_STRIDE2", "0") == "1")
        mods = pipeline.build_modules(w, device="cuda", n_depth_levels=D)
        st = pipeline.KeyframeState()
        errs, times = [], []
        with torch.no_grad():
            for rep in range(2):
                st = pipeline.KeyframeState()
                for fi, (ref_i, meas_i) in enumerate(clip["frames"]):
                    args = (T(clip["images"][ref_i])[None].cuda(), T(clip["poses"][ref_i])[None].cuda(),
                            [T(clip["images"][j])[None].cuda() for j in meas_i], [T(clip["poses"][j])[None].cuda() for j in meas_i], K.cuda())
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    pred, st = pipeline.keyframe(mods, st, *args, n_depth_levels=D)
                    torch.cuda.synchronize()
                    times.append(time.perf_counter() - t0)
                    if rep == 1:
                        errs.append(oracle.rel_l1_inverse_depth(pred.cpu().numpy(), golds[fi].numpy()))
        print("backend=%s terms=%d rel-L1(inv depth) vs oracle per frame: %s | ms/keyframe (eager, wall): %s"
              % (backend, terms, ["%.2e" % e for e in errs], ["%.2f" % (t * 1e3) for t in times[3:]]), flush=True)


if __name__ == "__main__":
    main()
