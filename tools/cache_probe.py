"""Development probe: the feature-cache paths (eager keyframe(cache=) and the pipelined engine) against the script sequence and the
oracle on the tensor-core backend, 1 and 3 terms (c2, seed-7 weights)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "deep-video-mvs_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import synth_data as synth  # noqa: E402
from oracle import dvmvs_oracle as oracle  # noqa: E402
from tests import helpers  # noqa: E402
from dvmvs import _ops as ops, pipeline  # noqa: E402

H, W, D, M, N = 256, 256, 64, 2, int(os.environ.get("NFRAMES", "24"))
w = helpers.oracle_weights(oracle, synth, 7, n_depth_levels=D)
clip = synth.make_clip(0, N, H, W, M)
T = torch.from_numpy
K = T(clip["K"])[None]
torch.set_num_threads(16)
golds, st = [], oracle.FusionnetState()
with torch.no_grad():
    for ref_i, meas_i in clip["frames"]:
        g, st = oracle.fusionnet_step(w, st, T(clip["images"][ref_i])[None], T(clip["poses"][ref_i])[None], [T(clip["images"][j])[None] for j in meas_i],
                                      [T(clip["poses"][j])[None] for j in meas_i], K, n_depth_levels=D)
        golds.append(g.numpy())
c = lambda a: T(np.ascontiguousarray(a))[None].cuda()
for terms in (1, 3):
    ops.set_conv_backend("tc", terms=terms, stride2=True)
    mods = helpers.build_product_modules(w, n_depth_levels=D)
    sa, sb = helpers.ProductState(), pipeline.KeyframeState()
    cache = pipeline.FeatureCache(capacity=8)
    pipe = pipeline.PipelinedFusionnet(mods, batch=1, height=H, width=W, n_measurement_frames=M, n_depth_levels=D, n_stages=5, feature_cache=8)
    ea, eb, ec, dab = [], [], [], []
    outs = []
    with torch.no_grad():
        for t, (ref_i, meas_i) in enumerate(clip["frames"]):
            args = (c(clip["images"][ref_i]), c(clip["poses"][ref_i]), [c(clip["images"][j]) for j in meas_i], [c(clip["poses"][j]) for j in meas_i], K.cuda())
            a, sa = helpers.product_fusionnet_step(mods, sa, *args, n_depth_levels=D)
            b, sb = pipeline.keyframe(mods, sb, *args, n_depth_levels=D, cache=cache, reference_id=ref_i, measurement_ids=meas_i)
            out = torch.empty((1, H, W), dtype=torch.float32, device="cuda")
            pipe.submit(*args, out=out, reference_id=ref_i, measurement_ids=meas_i)
            outs.append(out)
            ea.append(oracle.rel_l1_inverse_depth(a.cpu().numpy(), golds[t]))
            eb.append(oracle.rel_l1_inverse_depth(b.cpu().numpy(), golds[t]))
            dab.append(oracle.rel_l1_inverse_depth(b.cpu().numpy(), a.cpu().numpy()))
        pipe.synchronize()
    ec = [oracle.rel_l1_inverse_depth(o.cpu().numpy(), golds[t]) for t, o in enumerate(outs)]
    print("terms %d: script vs oracle max %.2e | keyframe(cache) vs oracle max %.2e | pipelined cache engine vs oracle max %.2e | cache vs script max %.2e"
          % (terms, max(ea), max(eb), max(ec), max(dab)))
