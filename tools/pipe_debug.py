import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import torch
import synth_data as synth
from dvmvs import pipeline, _ops as ops
from oracle import dvmvs_oracle as oracle
T = torch.from_numpy
H, W, D, M = 64, 96, 64, 2
ops.set_conv_backend(os.environ.get("BACKEND", "fp32"), terms=3, stride2=True)
shapes = oracle.state_dict_shapes(D)
w = {t: {k: T(v) for k, v in synth.make_state_dict(shapes[t], seed=11).items()} for t in shapes}
mods = pipeline.build_modules(w, device="cuda", n_depth_levels=D)
clip = synth.make_clip(5, 8, H, W, M)
K = T(clip["K"])[None].cuda()
st = pipeline.KeyframeState()
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pipe = pipeline.PipelinedFusionnet(mods, batch=1, height=H, width=W, n_measurement_frames=M, n_depth_levels=D, n_stages=ns)
exp, got = [], []
with torch.no_grad():
    for ref_i, meas_i in clip["frames"]:
        args = (T(clip["images"][ref_i])[None].cuda(), T(clip["poses"][ref_i])[None].cuda(), [T(clip["images"][j])[None].cuda() for j in meas_i],
                [T(clip["poses"][j])[None].cuda() for j in meas_i], K)
        a, st = pipeline.keyframe(mods, st, *args, n_depth_levels=D)
        exp.append(a.cpu().numpy())
        torch.cuda.synchronize()
        out = torch.empty((1, H, W), device="cuda")
        pipe.submit(*args, out=out)
        if os.environ.get("SYNC_EACH") == "1":
            pipe.synchronize()
        got.append(out)
    pipe.synchronize()
print("stages", ns, "serial", os.environ.get("DVMVS_PIPE_SERIAL"), "sync_each", os.environ.get("SYNC_EACH"),
      ["%.1e" % oracle.rel_l1_inverse_depth(g.cpu().numpy(), e) for g, e in zip(got, exp)])
