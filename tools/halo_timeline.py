"""Phase timeline of conv_halo_kernel CTAs (globaltimer marks) for the refine.1-shaped layer."""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import numpy as np, torch
import synth_data as synth
from dvmvs import _native as N, _ops as ops
DEV = "cuda"
for (name, B, H, W, cin, Cout, k, kc) in [("refine.1", 1, 256, 256, 32, 32, 5, 16), ("fpn.layer0 B=3", 3, 128, 128, 32, 32, 3, 32)]:
    x = torch.from_numpy(synth.tensor("tl/x", (B, H, W, cin), seed=1)).to(DEV)
    w = torch.from_numpy(synth.tensor("tl/w", (Cout, cin, k, k), seed=2, scale=0.05))
    pc = ops.PackedConv(w, None, None, stride=1, act=N.ACT_RELU)
    ph = ops.PackedConvHalo(pc, [cin], DEV, kc=kc, concat_padded=True)
    blk = ops.split_blocked([(x, False)])
    n_cta = ((W + 7) // 8) * ((H + 15) // 16) * B
    for _ in range(3):
        ops.conv2d_halo([blk], ph, want_f32=True, want_blk=True, want_nhwc=False)
    dbg = torch.zeros(n_cta * 8, dtype=torch.int64, device=DEV)
    N.lib().dvmvs_debug_set_halo_timeline(ctypes.c_void_p(dbg.data_ptr()))
    ops.conv2d_halo([blk], ph, want_f32=True, want_blk=True, want_nhwc=False)
    torch.cuda.synchronize()
    N.lib().dvmvs_debug_set_halo_timeline(None)
    t = dbg.cpu().numpy().reshape(n_cta, 8).astype(np.float64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    print("== %s: %d CTAs; kernel span %.1f us" % (name, n_cta, rel[:, 7].max()))
    d = np.diff(t, axis=1) / 1e3
    labels = ["setup(bar init,TMEM alloc)", "pdl wait", "first halo load", "mainloop (MMA issue)", "MMA drain -> acc ready", "epilogue", "dealloc+exit"]
    for i, l in enumerate(labels):
        print("   %-28s median %6.2f us   p90 %6.2f   max %6.2f" % (l, np.median(d[:, i]), np.percentile(d[:, i], 90), d[:, i].max()))
    print("   CTA lifetime median %.2f us; start times: median %.1f p90 %.1f max %.1f" % (np.median(rel[:, 7] - rel[:, 0]), np.median(rel[:, 0]), np.percentile(rel[:, 0], 90), rel[:, 0].max()))
