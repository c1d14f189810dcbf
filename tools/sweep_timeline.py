"""In-kernel timeline of plane_sweep_tc_kernel: clock64 stamps of the phases of the first CTAs (development aid).

    python tools/sweep_timeline.py [--clips 8] [--terms 1]"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "deep-video-mvs_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402

NAMES = {0: "start", 1: "planner: start of first tile", 2: "consumers: waiting for the first tile's plan", 3: "matrices", 4: "geometry", 5: "planner: first tile planned", 6: "consumers: plan received", 7: "reference tile ready", 56: "plan: scratch ready (round 1)", 57: "plan: boxes merged", 58: "plan: rows scanned", 59: "[rounds * 1000 + chunks]", 62: "write-out done", 63: "exit barrier"}
for k in range(9):
    for j, n in enumerate(("wait mma", "mma done", "next loads issued", "C1 done", "published", "look-ups done")):
        NAMES[8 + 6 * k + j] = "chunk %d: %s" % (k, n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--terms", type=int, default=1)
    args = ap.parse_args()
    import synth_data as synth
    from dvmvs import _native as N
    from dvmvs import _ops as ops
    dev = torch.device("cuda", 0)
    B, h, w, D, M = args.clips, 128, 128, 64, 2
    g = torch.Generator().manual_seed(1)
    f1 = (torch.randn(B, h, w, 32, generator=g) * 4).to(dev)
    f2 = [(torch.randn(B, h, w, 32, generator=g) * 4).to(dev) for _ in range(M)]
    K = torch.from_numpy(synth.intrinsics(2 * h, 2 * w))[None].repeat(B, 1, 1).to(dev)
    K[:, 0:2, :] /= 2.0
    pose1 = torch.from_numpy(synth.camera_pose(M))[None].repeat(B, 1, 1).to(dev)
    pose2 = [torch.from_numpy(synth.camera_pose(M - k))[None].repeat(B, 1, 1).to(dev) for k in range(1, M + 1)]
    p1, p2 = ops.split_planes(f1), [ops.split_planes(t) for t in f2]
    buf = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.plane_sweep_tc(p1, p2, pose1, pose2, K, 0.25, 20.0, D, terms=args.terms)
    torch.cuda.synchronize()
    N.lib().dvmvs_plane_sweep_tc_set_timeline(buf.data_ptr())
    ops.plane_sweep_tc(p1, p2, pose1, pose2, K, 0.25, 20.0, D, terms=args.terms)
    torch.cuda.synchronize()
    N.lib().dvmvs_plane_sweep_tc_set_timeline(None)
    t = buf.cpu().view(8, 64)
    for cta in (0, 5):
        row = t[cta]
        t0 = int(row[0])
        print("CTA %d (cycles since its start; SM clock)" % cta)
        prev = 0
        order = [0, 1, 2, 5, 6, 7] + list(range(8, 56)) + [62]
        for slot in order:
            v = int(row[slot])
            if v == 0:
                continue
            print("  %-32s %8d  (+%d)" % (NAMES[slot], v - t0, v - t0 - prev))
            prev = v - t0


if __name__ == "__main__":
    main()
