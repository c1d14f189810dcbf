"""Stand-alone replay time of every (stage, slot 0) CUDA graph of PipelinedFusionnet at config c2 -- shows which stage
bounds the pipelined throughput and how much of the gap to sum(stages) is contention.  GPU only."""
import os, sys, json
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import synth_data as synth
from dvmvs import pipeline, _ops as ops
from dvmvs.fusionnet.model import FeatureExtractor, FeatureShrinker, CostVolumeEncoder, LSTMFusion, CostVolumeDecoder

H = W = 256; D = 64; M = 2; dev = torch.device("cuda", 0)
ops.set_conv_backend("tc", terms=int(os.environ.get("DVMVS_TC_TERMS", "3")), stride2=True)
mods = {"fe": FeatureExtractor(), "fpn": FeatureShrinker(), "cve": CostVolumeEncoder(), "lstm": LSTMFusion(), "cvd": CostVolumeDecoder()}
for tag, m in mods.items():
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state_dict(shapes, seed=7).items()}, strict=True)
    m.to(dev).eval()
clip = synth.make_clip(0, 12, H, W, M)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
frames = [(T(clip["images"][r])[None], T(clip["poses"][r])[None], [T(clip["images"][j])[None] for j in m], [T(clip["poses"][j])[None] for j in m],
           T(clip["K"])[None]) for r, m in clip["frames"]]
res = {}
for ns in [int(a) for a in (sys.argv[1:] or ["3", "5"])]:
    eng = pipeline.PipelinedFusionnet(mods, batch=1, height=H, width=W, n_measurement_frames=M, n_depth_levels=D, n_stages=ns)
    with torch.no_grad():
        for f in frames[:2 * ns + 2]:
            eng.submit(*f)
        eng.synchronize(); torch.cuda.synchronize()
        slot = eng.slots[0]
        per = []
        for i in range(ns):
            g = slot["graph"][i][True if i == ns - 1 else False]
            s = eng.streams[i]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s):
                for _ in range(5): g.replay()
                a.record(s)
                for _ in range(50): g.replay()
                b.record(s)
            s.synchronize()
            per.append(a.elapsed_time(b) / 50 * 1e3)
        res[ns] = {"stage_us": [round(x, 1) for x in per], "sum_us": round(sum(per), 1), "kernels": eng._kernels}
print(json.dumps(res))
