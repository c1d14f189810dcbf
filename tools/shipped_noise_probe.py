"""Development probe: drift of the tensor-core backend (1 / 3 terms) against the CPU oracle over a long recurrent synthetic clip,
with the shipped or the seeded synthetic weights (WEIGHTS=shipped|synthetic, NFRAMES=100)."""
import sys, os, numpy as np, torch
REPO='/root/repo'
sys.path[:0]=[REPO, os.path.join(REPO,'deep-video-mvs_b200')]
import synth_data as synth
from oracle import dvmvs_oracle as oracle
from tests import helpers, scene_fixture
from dvmvs import _ops as ops, pipeline
H,W,D,M=256,256,64,2
w = scene_fixture.load_shipped_weights("fusionnet") if os.environ.get("WEIGHTS","shipped")=="shipped" else helpers.oracle_weights(oracle, synth, 7, n_depth_levels=64)
N=int(os.environ.get("NFRAMES","100"))
clip = synth.make_clip(0, N, H, W, M)
T=torch.from_numpy
K = T(clip["K"])[None]
golds=[]; st=oracle.FusionnetState()
torch.set_num_threads(16)
with torch.no_grad():
    for ref_i, meas_i in clip["frames"]:
        g, st = oracle.fusionnet_step(w, st, T(clip["images"][ref_i])[None], T(clip["poses"][ref_i])[None], [T(clip["images"][j])[None] for j in meas_i], [T(clip["poses"][j])[None] for j in meas_i], K, n_depth_levels=D)
        golds.append(g.numpy())
for terms in (1,3):
    ops.set_conv_backend("tc", terms=terms, stride2=True)
    mods = helpers.build_product_modules(w, n_depth_levels=D)
    stp = pipeline.KeyframeState(); errs=[]
    with torch.no_grad():
        for t,(ref_i, meas_i) in enumerate(clip["frames"]):
            c=lambda a: T(np.ascontiguousarray(a))[None].cuda()
            pred, stp = pipeline.keyframe(mods, stp, c(clip["images"][ref_i]), c(clip["poses"][ref_i]), [c(clip["images"][j]) for j in meas_i], [c(clip["poses"][j]) for j in meas_i], K.cuda(), n_depth_levels=D)
            errs.append(oracle.rel_l1_inverse_depth(pred.cpu().numpy(), golds[t]))
    print("terms", terms, ["%.1e"%e for e in errs[::8]], "max %.2e"%max(errs))
print("depth stats", float(golds[-1].min()), float(golds[-1].max()), float(golds[-1].mean()))
