"""The per-keyframe call sequence of the reference's test drivers (fusionnet/run-testing.py:145-202,
pairnet/run-testing.py:139-164) as a function over the drop-in modules, for callers that do not want to copy
the script's loop body (bench.py, smoke test, examples).  It makes exactly the module / utils calls the script
makes, with the same keyword arguments."""
import torch

from . import _ops as ops
from ._base import no_auto_graph
from .utils import (cost_volume_fusion, get_non_differentiable_rectangle_depth_estimation,
                    get_warp_grid_for_cost_volume_calculation)


class KeyframeState:
    """Recurrent state the caller carries between keyframes of one clip (run-testing.py:86-88)."""

    def __init__(self):
        self.lstm_state = None
        self.previous_depth = None
        self.previous_pose = None

    def reset(self):          # "TRACKING LOST" (run-testing.py:97-101)
        self.__init__()


def build_modules(weights, device="cuda", n_depth_levels=64, pairnet=False):
    """Constructs the modules, strict-loads `weights` (dict 'fe','fpn','cve'[,'lstm'],'cvd' -> state dict), eval()."""
    from .config import Config
    old = Config.train_n_depth_levels
    Config.train_n_depth_levels = n_depth_levels          # aggregator0 has D+32 inputs (fusionnet/model.py:170)
    try:
        if pairnet:
            from .pairnet import model as mm
        else:
            from .fusionnet import model as mm
        mods = {"fe": mm.FeatureExtractor(), "fpn": mm.FeatureShrinker(), "cve": mm.CostVolumeEncoder(), "cvd": mm.CostVolumeDecoder()}
        if not pairnet:
            mods["lstm"] = mm.LSTMFusion()
    finally:
        Config.train_n_depth_levels = old
    for tag, m in mods.items():
        m.load_state_dict(weights[tag], strict=True)
        m.to(device).eval()
    return mods


class FeatureCache:
    """half-resolution FPN features of past reference frames, keyed by caller-supplied frame ids.

    The reference recomputes FeatureExtractor + FeatureShrinker for every measurement frame of every keyframe
    (run-testing.py:153-156, run-testing-online.py:159-162) although each measurement frame was the reference frame of an
    earlier keyframe (the keyframe buffer only holds past reference frames).  The modules are in eval mode, so those
    features are the same numbers; this cache keeps them in a fixed ring of device buffers (FIFO eviction, capacity =
    the keyframe buffer size, Config.test_keyframe_buffer_size) and hands them back.  Ids are explicit because the scripts
    re-upload the images as fresh tensors each keyframe: nothing on the device identifies a frame without a host sync.
    A miss is not an error -- the caller computes the features from the image and stores them."""

    def __init__(self, capacity=30):
        if capacity < 1:
            raise ValueError("FeatureCache capacity must be >= 1")
        self.capacity = int(capacity)
        self._ring = [None] * self.capacity      # (B, h, w, 32) channel-last buffers, allocated on first use
        self._index = {}                          # frame id -> ring index
        self._owner = [None] * self.capacity
        self._next = 0
        self.hits = 0
        self.misses = 0

    def clear(self):
        self._index.clear()
        self._owner = [None] * self.capacity

    def __contains__(self, frame_id):
        return frame_id in self._index

    def lookup(self, frame_id):
        """Cached (B,32,h,w) API tensor (channels_last view of the ring buffer) or None."""
        idx = self._index.get(frame_id)
        if idx is None:
            self.misses += 1
            return None
        self.hits += 1
        return self._ring[idx].permute(0, 3, 1, 2)

    def store(self, frame_id, half_features):
        """Copies (B,32,h,w) features into the ring on the current stream (a D2D copy; the ring outlives the caller's tensor)."""
        idx = self._index.get(frame_id)
        if idx is None:
            idx = self._next
            self._next = (self._next + 1) % self.capacity
            if self._owner[idx] is not None:
                del self._index[self._owner[idx]]
            self._owner[idx] = frame_id
            self._index[frame_id] = idx
        src = half_features.permute(0, 2, 3, 1)
        if self._ring[idx] is None or self._ring[idx].shape != src.shape or self._ring[idx].device != src.device:
            self._ring[idx] = torch.empty(tuple(src.shape), dtype=torch.float32, device=src.device)
        self._ring[idx].copy_(src)
        return idx


def feature_stage(mods, reference_image, reference_pose, measurement_images, measurement_poses, full_K,
                  min_depth=0.25, max_depth=20.0, n_depth_levels=64, batch_features=True, cache=None, reference_id=None,
                  measurement_ids=None):
    """First half of a keyframe -- everything that does not depend on the recurrent state: FeatureExtractor +
    FeatureShrinker on the reference and measurement images and the fused plane-sweep cost volume
    (run-testing.py:153-171).  Returns (f2, f4, f8, f16, cost_volume, half_K).

    With `cache` (FeatureCache) and frame ids, measurement frames whose half-resolution features are cached skip
    FeatureExtractor + FeatureShrinker (row f1); the reference frame's features are stored under `reference_id`."""
    B = reference_image.shape[0]
    half_K = full_K.clone()
    half_K[:, 0:2, :] = half_K[:, 0:2, :] / 2.0
    if cache is not None:
        if measurement_ids is None or len(measurement_ids) != len(measurement_images):
            raise ValueError("feature cache: need one frame id per measurement image")
        meas_half = [cache.lookup(i) for i in measurement_ids]
        todo = [m for m, t in enumerate(meas_half) if t is None]
        stacked = torch.cat([reference_image] + [measurement_images[m] for m in todo], dim=0) if todo else reference_image
        a2, a4, a8, a16 = mods["fpn"](*mods["fe"](stacked))
        if todo:
            f2, f4, f8, f16 = ops.batch_slice(a2, 0, B), a4[:B], a8[:B], a16[:B]
        else:
            f2, f4, f8, f16 = a2, a4, a8, a16
        for n, m in enumerate(todo):
            meas_half[m] = ops.batch_slice(a2, (n + 1) * B, (n + 2) * B)
    elif batch_features and len(measurement_images) > 0:
        stacked = torch.cat([reference_image] + list(measurement_images), dim=0)
        a2, a4, a8, a16 = mods["fpn"](*mods["fe"](stacked))
        f2, f4, f8, f16 = ops.batch_slice(a2, 0, B), a4[:B], a8[:B], a16[:B]
        meas_half = [ops.batch_slice(a2, (m + 1) * B, (m + 2) * B) for m in range(len(measurement_images))]
    else:
        meas_half = []
        for im in measurement_images:
            half, _, _, _ = mods["fpn"](*mods["fe"](im))
            meas_half.append(half)
        f2, f4, f8, f16 = mods["fpn"](*mods["fe"](reference_image))
    cv = cost_volume_fusion(image1=f2, image2s=meas_half, pose1=reference_pose, pose2s=measurement_poses, K=half_K,
                            warp_grid=None, min_depth=min_depth, max_depth=max_depth, n_depth_levels=n_depth_levels,
                            device=reference_image.device, dot_product=True)
    if cache is not None:       # after the sweep: hits are views of ring entries that a store may evict and overwrite
        for m in todo:
            cache.store(measurement_ids[m], meas_half[m])
        if reference_id is not None:
            cache.store(reference_id, f2)
    return f2, f4, f8, f16, cv, half_K


def recurrent_stage(mods, state, features, reference_image, reference_pose, full_K):
    """Second half of a keyframe: cost-volume encoder, depth re-projection + ConvLSTM fusion (fusionnet only), decoder
    (run-testing.py:173-202).  `features` is feature_stage()'s return value.  Returns (depth (B,H,W), state)."""
    f2, f4, f8, f16, cv, half_K = features
    B, _, H, W = reference_image.shape
    device = reference_image.device
    s0, s1, s2, s3, bottom = mods["cve"](features_half=f2, features_quarter=f4, features_one_eight=f8,
                                         features_one_sixteen=f16, cost_volume=cv)
    if "lstm" in mods:
        lstm_K = full_K.clone()
        lstm_K[:, 0:2, :] = lstm_K[:, 0:2, :] / 32.0
        if state.previous_depth is not None:
            de = get_non_differentiable_rectangle_depth_estimation(reference_pose_torch=reference_pose,
                                                                   measurement_pose_torch=state.previous_pose,
                                                                   previous_depth_torch=state.previous_depth,
                                                                   full_K_torch=full_K, half_K_torch=half_K,
                                                                   original_height=H, original_width=W)
            de = de[:, :, ::16, ::16].contiguous()      # == F.interpolate(scale_factor=1/16, mode='nearest') (run-testing.py:187-189)
        else:
            de = torch.zeros(size=(B, 1, H // 32, W // 32), device=device)
        state.lstm_state = mods["lstm"](current_encoding=bottom, current_state=state.lstm_state,
                                        previous_pose=state.previous_pose, current_pose=reference_pose,
                                        estimated_current_depth=de, camera_matrix=lstm_K)
        bottom = state.lstm_state[0]
    pred = mods["cvd"](reference_image, s0, s1, s2, s3, bottom)[0]
    state.previous_depth = pred.view(B, 1, H, W)
    state.previous_pose = reference_pose
    return pred, state


def keyframe(mods, state, reference_image, reference_pose, measurement_images, measurement_poses, full_K,
             min_depth=0.25, max_depth=20.0, n_depth_levels=64, batch_features=True, cache=None, reference_id=None,
             measurement_ids=None):
    """One keyframe for B independent clips (tensors batched on dim 0, all CUDA).  With 'lstm' in mods this is the
    fusionnet loop body, without it the pairnet one.  Returns (depth (B,H,W), state).

    batch_features=True runs FeatureExtractor + FeatureShrinker ONCE over the reference and the M measurement images
    stacked on the batch axis (eval-mode BatchNorm: identical results, 1/(M+1) of the launches); False reproduces the
    script's M+1 separate passes (run-testing.py:153-159).  cache / reference_id / measurement_ids: see FeatureCache."""
    feats = feature_stage(mods, reference_image, reference_pose, measurement_images, measurement_poses, full_K, min_depth,
                          max_depth, n_depth_levels, batch_features, cache, reference_id, measurement_ids)
    return recurrent_stage(mods, state, feats, reference_image, reference_pose, full_K)


class GraphedFusionnet:
    """The keyframe loop body captured once into CUDA graphs (static shapes) and replayed: removes the ~300 host-side
    launches per keyframe.  Built from the same drop-in modules; two graphs are captured lazily -- keyframe without
    recurrent state (first frame / after reset) and steady state (hidden-state warp + depth re-projection on).

        eng = GraphedFusionnet(mods, batch=B, height=H, width=W, n_measurement_frames=M)
        depth = eng.step(ref_img, ref_pose, [meas_imgs], [meas_poses], full_K)      # CPU (pinned) or CUDA tensors

    step() copies the inputs into static device buffers on the current stream (H2D when they are host tensors),
    replays the graph and returns the static (B,H,W) depth buffer (valid until the next step)."""

    def __init__(self, mods, batch, height, width, n_measurement_frames, min_depth=0.25, max_depth=20.0, n_depth_levels=64,
                 device=None):
        self.mods, self.B, self.H, self.W, self.M = mods, batch, height, width, n_measurement_frames
        self.min_depth, self.max_depth, self.D = min_depth, max_depth, n_depth_levels
        dev = device or next(mods["fe"].parameters()).device
        self.device = dev
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.ref_image, self.ref_pose, self.full_K = z(batch, 3, height, width), z(batch, 4, 4), z(batch, 3, 3)
        self.meas_images = [z(batch, 3, height, width) for _ in range(n_measurement_frames)]
        self.meas_poses = [z(batch, 4, 4) for _ in range(n_measurement_frames)]
        self.state = KeyframeState()
        self._graphs = {}
        self._capture_stream = None
        self.kernels_per_replay = {}
        self._static_state = None     # (h, c, prev_depth, prev_pose) buffers the steady-state graph reads and rewrites
        self._out = None
        self._has_state = False

    def reset(self):
        self._has_state = False

    def _load_inputs(self, reference_image, reference_pose, measurement_images, measurement_poses, full_K):
        self.ref_image.copy_(reference_image, non_blocking=True)
        self.ref_pose.copy_(reference_pose, non_blocking=True)
        self.full_K.copy_(full_K, non_blocking=True)
        for dst, src in zip(self.meas_images, measurement_images):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(self.meas_poses, measurement_poses):
            dst.copy_(src, non_blocking=True)

    def _body(self, with_state):
        st = KeyframeState()
        if with_state:
            h, c, pd, pp = self._static_state
            st.lstm_state, st.previous_depth, st.previous_pose = (h, c), pd, pp
        pred, st = keyframe(self.mods, st, self.ref_image, self.ref_pose, self.meas_images, self.meas_poses, self.full_K,
                            self.min_depth, self.max_depth, self.D)
        return pred, st

    def _capture(self, with_state):
        # warm-up on a side stream (allocations, weight packing, function attributes), then capture
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(device=self.device)
        s = self._capture_stream        # same stream for warm-up and capture: per-stream scratch is allocated outside the graph
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s), torch.no_grad(), no_auto_graph():
            for _ in range(2):
                pred, st = self._body(with_state)
        torch.cuda.current_stream(self.device).wait_stream(s)
        s.synchronize()
        if self._static_state is None:
            self._static_state = (st.lstm_state[0].clone(), st.lstm_state[1].clone(), st.previous_depth.clone(), self.ref_pose.clone())
            self._out = torch.empty_like(pred)
        from . import _native
        g = torch.cuda.CUDAGraph()
        n0 = _native.launch_count()
        with torch.no_grad(), torch.cuda.graph(g, stream=s):
            pred, st = self._body(with_state)
            self.kernels_per_replay[with_state] = _native.launch_count() - n0   # our kernels captured in this graph
            h, c, pd, pp = self._static_state
            self._out.copy_(pred)
            # new recurrent state -> static buffers (read by the next replay)
            h.copy_(st.lstm_state[0])
            c.copy_(st.lstm_state[1])
            pd.copy_(st.previous_depth)
            pp.copy_(self.ref_pose)
        self._graphs[with_state] = g

    def step(self, reference_image, reference_pose, measurement_images, measurement_poses, full_K):
        if len(measurement_images) != self.M:
            raise ValueError("GraphedFusionnet was built for %d measurement frames, got %d" % (self.M, len(measurement_images)))
        with_state = self._has_state
        if with_state not in self._graphs:
            if with_state and self._static_state is None:
                raise RuntimeError("steady-state graph requested before any keyframe ran")
            saved = None
            if self._static_state is not None:
                saved = [t.clone() for t in self._static_state]
            self._load_inputs(reference_image, reference_pose, measurement_images, measurement_poses, full_K)
            self._capture(with_state)
            if saved is not None:                      # capture warm-ups must not advance the recurrent state
                for dst, src in zip(self._static_state, saved):
                    dst.copy_(src)
        self._load_inputs(reference_image, reference_pose, measurement_images, measurement_poses, full_K)
        self._graphs[with_state].replay()
        self._has_state = True
        return self._out


def _stage_side_inputs(slot):
    """Small state-independent preparations the last stage would otherwise do on the loop-carried critical path: the
    channel-last copy of the reference image the decoder's refinement convolutions read, and the 1/32 intrinsics."""
    from . import _ops as ops
    slot["ref_cl"] = ops.to_api(ops.to_nhwc(slot["ref_image"], "image"))      # (B,3,H,W) view with channels_last strides
    lstm_K = slot["full_K"].clone()
    lstm_K[:, 0:2, :] = lstm_K[:, 0:2, :] / 32.0
    slot["lstm_K"] = lstm_K


def _stacked_images(slot):
    """Reference + measurement images on the batch axis; the reference image alone when the measurement features come
    from the feature cache (slot["meas_half"], row f1)."""
    if slot.get("meas_half") is not None:
        return slot["ref_image"]
    return torch.cat([slot["ref_image"]] + list(slot["meas_images"]), dim=0)


def _stage_fe(mods, slot):
    """Stage 1 of 3: MnasNet trunk on the reference + measurement images stacked on the batch axis."""
    _stage_side_inputs(slot)
    return mods["fe"](_stacked_images(slot))


def _stage_fe_head(mods, slot):
    """MnasNet trunk up to layer3 (1/8 resolution) on the stacked images."""
    _stage_side_inputs(slot)
    return mods["fe"].forward_head(_stacked_images(slot))


def _stage_fe_tail(mods, slot, head):
    """MnasNet layers 4-5 (1/16, 1/32 resolution) from the head's outputs."""
    return mods["fe"].forward_tail(head)


def _sweep_from_pyramid(slot, pyramid, base, min_depth, max_depth, n_depth_levels):
    """Fused plane sweep of ONE keyframe from a feature pyramid whose batch axis stacks [reference, measurement 1..M] x B clips
    starting at row `base` (the lookahead engine's pyramid holds several keyframes)."""
    B = slot["ref_image"].shape[0]
    M = len(slot["meas_images"])
    a2, a4, a8, a16 = pyramid
    if slot.get("meas_half") is not None:         # feature cache: batch = the reference frames only
        f2, f4, f8, f16 = a2, a4, a8, a16
        meas_half = [t.permute(0, 3, 1, 2) for t in slot["meas_half"]]
        slot["ref_half"] = f2                     # the engine copies it into the cache ring after the stage's graph
    else:
        f2, f4, f8, f16 = ops.batch_slice(a2, base, base + B), a4[base:base + B], a8[base:base + B], a16[base:base + B]
        meas_half = [ops.batch_slice(a2, base + (m + 1) * B, base + (m + 2) * B) for m in range(M)]
    half_K = slot["full_K"].clone()
    half_K[:, 0:2, :] = half_K[:, 0:2, :] / 2.0
    cv = cost_volume_fusion(image1=f2, image2s=meas_half, pose1=slot["ref_pose"], pose2s=slot["meas_poses"], K=half_K, warp_grid=None,
                            min_depth=min_depth, max_depth=max_depth, n_depth_levels=n_depth_levels, device=f2.device, dot_product=True)
    return (f2, f4, f8, f16, cv), half_K


def _stage_sweep(mods, slot, fe_out, min_depth, max_depth, n_depth_levels):
    """Feature pyramid + fused plane sweep."""
    return _sweep_from_pyramid(slot, mods["fpn"](*fe_out), 0, min_depth, max_depth, n_depth_levels)


def _stage_enc(mods, slot, swept):
    """Cost-volume encoder (still independent of the recurrent state)."""
    (f2, f4, f8, f16, cv), half_K = swept
    enc = mods["cve"](features_half=f2, features_quarter=f4, features_one_eight=f8, features_one_sixteen=f16, cost_volume=cv)
    if "lstm" in mods:      # the input half of the ConvLSTM gate convolution does not depend on the recurrent state either
        slot["input_gates"] = mods["lstm"].lstm_cell.input_gates(enc[4])
    return enc, half_K


def _stage_mid(mods, slot, fe_out, min_depth, max_depth, n_depth_levels):
    """Feature pyramid, fused plane sweep, cost-volume encoder."""
    return _stage_enc(mods, slot, _stage_sweep(mods, slot, fe_out, min_depth, max_depth, n_depth_levels))


def _stage_rec(mods, state, slot, enc, half_K):
    """Last stage: depth re-projection + ConvLSTM fusion + decoder -- the only part with a loop-carried dependence."""
    s0, s1, s2, s3, bottom = enc
    reference_image, reference_pose, full_K = slot.get("ref_cl", slot["ref_image"]), slot["ref_pose"], slot["full_K"]
    B, _, H, W = reference_image.shape
    lstm_K = slot.get("lstm_K")
    if lstm_K is None:
        lstm_K = full_K.clone()
        lstm_K[:, 0:2, :] = lstm_K[:, 0:2, :] / 32.0
    if state.previous_depth is not None:
        de = get_non_differentiable_rectangle_depth_estimation(reference_pose_torch=reference_pose, measurement_pose_torch=state.previous_pose,
                                                               previous_depth_torch=state.previous_depth, full_K_torch=full_K,
                                                               half_K_torch=half_K, original_height=H, original_width=W)
        de = de[:, :, ::16, ::16].contiguous()
    else:
        de = torch.zeros(size=(B, 1, H // 32, W // 32), device=reference_image.device)
    state.lstm_state = mods["lstm"](current_encoding=bottom, current_state=state.lstm_state, previous_pose=state.previous_pose,
                                    current_pose=reference_pose, estimated_current_depth=de, camera_matrix=lstm_K,
                                    input_gates=slot.get("input_gates"))
    pred = mods["cvd"](reference_image, s0, s1, s2, s3, state.lstm_state[0])[0]
    state.previous_depth = pred.view(B, 1, H, W)
    state.previous_pose = reference_pose
    return pred, state


class PipelinedFusionnet:
    """Throughput engine for ONE clip (or B clips batched): consecutive keyframes are software-pipelined over CUDA streams.
    Only the last stage (depth re-projection, ConvLSTM, decoder) depends on the previous keyframe; everything before it
    -- n_stages=2: [FE + FPN + plane sweep | encoder + ConvLSTM + decoder]; n_stages=3 (default):
    [FE | FPN + plane sweep + encoder | ConvLSTM + decoder]; 4 and 5 split the MnasNet trunk at layer3 and (5) the
    encoder off the plane sweep -- runs ahead for the following keyframes on its own stream.
    Each (stage, slot) is a captured CUDA graph over n_stages-buffered static tensors.  Results are identical to
    GraphedFusionnet / keyframe(): the per-keyframe dataflow is unchanged, only independent work of neighbouring
    keyframes overlaps (tests/test_gpu_parity.py).

        eng = PipelinedFusionnet(mods, batch=B, height=H, width=W, n_measurement_frames=M)
        for frame in stream:  eng.submit(*frame, out=pinned_host_tensor_or_None)
        eng.synchronize()
    """

    def __init__(self, mods, batch, height, width, n_measurement_frames, min_depth=0.25, max_depth=20.0, n_depth_levels=64,
                 device=None, n_stages=3, feature_cache=0):
        if n_stages not in (2, 3, 4, 5):
            raise ValueError("n_stages must be 2, 3, 4 or 5")
        if feature_cache and n_stages < 3:
            raise ValueError("the feature cache needs n_stages >= 3 (feature pyramid + plane sweep in one stage)")
        self.mods, self.B, self.H, self.W, self.M = mods, batch, height, width, n_measurement_frames
        self.min_depth, self.max_depth, self.D = min_depth, max_depth, n_depth_levels
        dev = device or next(mods["fe"].parameters()).device
        self.device = dev
        self.n_stages = n_stages
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.slots = []
        for _ in range(n_stages):
            self.slots.append({"ref_image": z(batch, 3, height, width), "ref_pose": z(batch, 4, 4), "full_K": z(batch, 3, 3),
                               "meas_images": [z(batch, 3, height, width) for _ in range(n_measurement_frames)],
                               "meas_poses": [z(batch, 4, 4) for _ in range(n_measurement_frames)],
                               "out": [None] * n_stages, "depth": z(batch, height, width),
                               "graph": [dict() for _ in range(n_stages)],
                               "done": [torch.cuda.Event() for _ in range(n_stages)]})
        # row f1: measurement features from a ring of past reference-frame features instead of M extra FE + FPN passes
        if feature_cache and feature_cache < n_measurement_frames + 1:
            raise ValueError("feature_cache capacity must be at least n_measurement_frames + 1")
        self.cache = FeatureCache(feature_cache) if feature_cache else None
        if self.cache is not None:
            for slot in self.slots:
                slot["meas_half"] = [z(batch, height // 2, width // 2, 32) for _ in range(n_measurement_frames)]
        self._sweep_stage = {2: 0, 3: 1, 4: 2, 5: 2}[n_stages]      # the stage that holds FPN + plane sweep
        import os as _os
        # DVMVS_PIPE_PRIO=1 gives the last stage (the one carrying the loop dependence) a high-priority stream; measured
        # slower on B200 (922 vs 1067 keyframes/s at 3 stages), so it is off by default
        prio = _os.environ.get("DVMVS_PIPE_PRIO", "0") == "1"
        self.streams = [torch.cuda.Stream(device=dev, priority=(-1 if (prio and i == n_stages - 1) else 0)) for i in range(n_stages)]
        # DVMVS_PIPE_REC_PDL=0: capture the last stage without programmatic dependent launch (its early-launched CTAs then do
        # not sit on SMs waiting for their predecessor while other stages could use them) -- experiment switch
        self._rec_pdl = _os.environ.get("DVMVS_PIPE_REC_PDL", "1") == "1"
        # DVMVS_PIPE_PDL=1: capture the stages WITH programmatic dependent launch.  Off by default: with several stage graphs in
        # flight, early-launched CTAs that sit in griddepcontrol.wait hold shared memory / TMEM other stages' kernels could
        # use (B200, 5 stages, 256 x 256: 508 us per keyframe without, 546 us with; a stage replayed alone gains only ~2 % from it)
        self._pdl = _os.environ.get("DVMVS_PIPE_PDL", "0") == "1"
        if _os.environ.get("DVMVS_PIPE_SERIAL") == "1":        # debugging aid: all stages on one stream (no overlap)
            self.streams = [self.streams[0]] * n_stages
        self.stream_a, self.stream_b = self.streams[0], self.streams[-1]      # first / last stage (timing hooks)
        self._static_state = None
        self._has_state = False
        self.t = 0
        self._kernels = [0] * n_stages
        self.kernels_per_keyframe = 0

    def reset(self):
        """TRACKING LOST / new clip: drops the recurrent state (the feature cache is keyed by frame id and stays valid)."""
        self._has_state = False

    def load_state(self, lstm_state, previous_depth, previous_pose):
        """Continue a clip whose first keyframes ran elsewhere (e.g. through the module calls with fewer measurement frames):
        installs (h, c), the previous depth (B,1,H,W) and the previous pose as the recurrent state of the next submit().
        Needs the static state buffers, i.e. prime() or one earlier keyframe."""
        if self._static_state is None:
            raise RuntimeError("load_state: call prime() (or submit one keyframe) first")
        self.synchronize()
        h, c, pd, pp = self._static_state
        with torch.no_grad():
            h.copy_(lstm_state[0])
            c.copy_(lstm_state[1])
            pd.copy_(previous_depth.reshape(pd.shape))
            pp.copy_(previous_pose)
        torch.cuda.current_stream(self.device).synchronize()
        self._has_state = True

    # -- stage bodies -----------------------------------------------------------------------------------------------
    def _run_stage(self, i, slot, with_state):
        last = self.n_stages - 1
        prev = slot["out"][i - 1] if i > 0 else None
        depth_args = (self.min_depth, self.max_depth, self.D)
        if self.n_stages == 2:
            if i == 0:
                return feature_stage(self.mods, slot["ref_image"], slot["ref_pose"], slot["meas_images"], slot["meas_poses"],
                                     slot["full_K"], self.min_depth, self.max_depth, self.D)
        elif i < last:
            # stage plans (all stages before the last are independent of the recurrent state):
            #   3: FE | FPN + sweep + encoder | rec        4: FE head | FE tail | FPN + sweep + encoder | rec
            #   5: FE head | FE tail | FPN + sweep | encoder | rec
            plan = {3: ("fe", "mid"), 4: ("head", "tail", "mid"), 5: ("head", "tail", "sweep", "enc")}[self.n_stages]
            kind = plan[i]
            if kind == "fe":
                return _stage_fe(self.mods, slot)
            if kind == "head":
                return _stage_fe_head(self.mods, slot)
            if kind == "tail":
                return _stage_fe_tail(self.mods, slot, prev)
            if kind == "mid":
                return _stage_mid(self.mods, slot, prev, *depth_args)
            if kind == "sweep":
                return _stage_sweep(self.mods, slot, prev, *depth_args)
            return _stage_enc(self.mods, slot, prev)
        st = KeyframeState()
        if with_state:
            h, c, pd, pp = self._static_state
            st.lstm_state, st.previous_depth, st.previous_pose = (h, c), pd, pp
        if self.n_stages == 2:
            return recurrent_stage(self.mods, st, slot["out"][0], slot["ref_image"], slot["ref_pose"], slot["full_K"])
        enc, half_K = slot["out"][last - 1]
        return _stage_rec(self.mods, st, slot, enc, half_K)

    def _capture(self, i, slot, with_state):
        from . import _native
        last = self.n_stages - 1
        stream = self.streams[i]
        torch.cuda.synchronize(self.device)
        saved = [t.clone() for t in self._static_state] if (i == last and self._static_state is not None) else None
        pdl_off = (not self._pdl) or (i == last and not self._rec_pdl)
        if pdl_off:
            _native.lib().dvmvs_set_programmatic_launch(0)
        with torch.cuda.stream(stream), torch.no_grad(), no_auto_graph():
            for _ in range(2):
                res = self._run_stage(i, slot, with_state)
        stream.synchronize()
        if i == last and self._static_state is None:
            pred, st = res
            self._static_state = (st.lstm_state[0].clone(), st.lstm_state[1].clone(), st.previous_depth.clone(), slot["ref_pose"].clone())
        g = torch.cuda.CUDAGraph()
        n0 = _native.launch_count()
        with torch.no_grad(), torch.cuda.graph(g, stream=stream):
            res = self._run_stage(i, slot, with_state)
            if i == last:
                pred, st = res
                h, c, pd, pp = self._static_state
                slot["depth"].copy_(pred)
                h.copy_(st.lstm_state[0])
                c.copy_(st.lstm_state[1])
                pd.copy_(st.previous_depth)
                pp.copy_(slot["ref_pose"])
            else:
                slot["out"][i] = res
        self._kernels[i] = _native.launch_count() - n0
        if pdl_off:
            _native.lib().dvmvs_set_programmatic_launch(-1)
        slot["graph"][i][with_state if i == last else False] = g
        if saved is not None:
            for dst, src in zip(self._static_state, saved):
                dst.copy_(src)
        torch.cuda.synchronize(self.device)

    # -- steady state ----------------------------------------------------------------------------------------------
    def submit(self, reference_image, reference_pose, measurement_images, measurement_poses, full_K, out=None,
               reference_id=None, measurement_ids=None):
        """Enqueue keyframe t (inputs CPU-pinned or CUDA).  If `out` (pinned host or CUDA tensor (B,H,W)) is given the
        depth is copied into it on the last stage's stream; otherwise read eng.depth_of(t) after synchronisation.

        Engines built with feature_cache=N take frame ids: `measurement_ids[m]` names measurement frame m, `reference_id`
        the reference frame.  A measurement frame whose id is in the cache needs no image (pass None): its features are
        copied from the ring; on a miss the features are computed from the image (eagerly, on the sweep stage's stream)
        and cached.  The reference frame's features enter the ring under `reference_id`."""
        n, last = self.n_stages, self.n_stages - 1
        slot = self.slots[self.t % n]
        with_state = self._has_state
        hits = None
        if self.cache is not None:
            if measurement_ids is None or len(measurement_ids) != self.M:
                raise ValueError("feature cache: need %d measurement frame ids" % self.M)
            hits = [i in self.cache for i in measurement_ids]
            measurement_images = list(measurement_images)
            for m, hit in enumerate(hits):
                if not hit and measurement_images[m] is None:
                    raise ValueError("feature cache miss for frame id %r and no image given" % (measurement_ids[m],))
        elif reference_id is not None or measurement_ids is not None:
            raise ValueError("frame ids given but the engine was built without feature_cache")
        # the inputs were produced on the caller's stream: order the first stage after it and keep CUDA inputs alive
        # (caching-allocator wise) until our copies on the stage stream have run
        caller = torch.cuda.current_stream(self.device)
        self.streams[0].wait_stream(caller)
        for t_in in [reference_image, reference_pose, full_K] + list(measurement_images) + list(measurement_poses):
            if t_in is not None and t_in.is_cuda:
                t_in.record_stream(self.streams[0])
        if out is not None and out.is_cuda:
            out.record_stream(self.streams[last])
        for i in range(n):
            stream = self.streams[i]
            key = with_state if i == last else False
            with torch.cuda.stream(stream):
                if i == 0:
                    stream.wait_event(slot["done"][last])         # slot reuse: keyframe t-n has left the pipeline
                    slot["ref_image"].copy_(reference_image, non_blocking=True)
                    slot["ref_pose"].copy_(reference_pose, non_blocking=True)
                    slot["full_K"].copy_(full_K, non_blocking=True)
                    for m, (dst, src) in enumerate(zip(slot["meas_images"], measurement_images)):
                        if hits is None or not hits[m]:            # cached measurement frames need no image upload
                            dst.copy_(src, non_blocking=True)
                    for dst, src in zip(slot["meas_poses"], measurement_poses):
                        dst.copy_(src, non_blocking=True)
                else:
                    stream.wait_event(slot["done"][i - 1])
                if hits is not None and i == self._sweep_stage:
                    self._fill_measurement_features(slot, measurement_ids, hits)
                if key not in slot["graph"][i]:
                    self._capture(i, slot, with_state)
                slot["graph"][i][key].replay()
                if hits is not None and i == self._sweep_stage and reference_id is not None:
                    self.cache.store(reference_id, slot["ref_half"])
                if i == last and out is not None:
                    out.copy_(slot["depth"], non_blocking=True)
                slot["done"][i].record(stream)
        self._has_state = True
        self.kernels_per_keyframe = sum(self._kernels)
        self.t += 1
        return self.t - 1

    def _fill_measurement_features(self, slot, measurement_ids, hits):
        """Feature-cache engines, on the sweep stage's stream before its graph: slot["meas_half"][m] <- ring entry (hit) or
        <- FeatureShrinker(FeatureExtractor(image)) computed here and stored in the ring (miss)."""
        for m, hit in enumerate(hits):            # hits first: a miss's store may evict the oldest ring entry
            if hit:
                slot["meas_half"][m].copy_(self.cache.lookup(measurement_ids[m]).permute(0, 2, 3, 1))
        for m, hit in enumerate(hits):
            if not hit:
                self.cache.misses += 1
                with torch.no_grad():
                    half, _, _, _ = self.mods["fpn"](*self.mods["fe"](slot["meas_images"][m]))
                self.cache.store(measurement_ids[m], half)
                slot["meas_half"][m].copy_(half.permute(0, 2, 3, 1))

    def prime(self, reference_image, reference_pose, measurement_images, measurement_poses, full_K):
        """Captures every (stage, slot) graph -- including both variants of the last stage -- by running 2*n_stages throw-away
        keyframes, then resets the clip state.  Optional: submit() captures lazily; call this to keep the one-off
        captures out of a timed or latency-sensitive region."""
        for k in range(2 * self.n_stages):
            if self.cache is not None:
                self.submit(reference_image, reference_pose, measurement_images, measurement_poses, full_K,
                            reference_id=("prime", k), measurement_ids=[("prime-m", k, m) for m in range(self.M)])
            else:
                self.submit(reference_image, reference_pose, measurement_images, measurement_poses, full_K)
        self.synchronize()
        self.reset()
        if self.cache is not None:
            self.cache.clear()
            self.cache.hits = self.cache.misses = 0

    def depth_of(self, t):
        return self.slots[t % self.n_stages]["depth"]

    def flush(self):
        """Nothing is ever held back by this engine (LookaheadFusionnet buffers keyframes; same call there launches them)."""

    def synchronize(self):
        for s in self.streams:
            s.synchronize()


class LookaheadFusionnet:
    """Throughput engine with everything that does NOT depend on the recurrent state batched over TIME: FeatureExtractor,
    FeatureShrinker, the plane sweep and the cost-volume encoder run once per group of `lookahead` consecutive keyframes
    (lookahead x B "clips", each with its own poses and its own M measurement frames) instead of once per keyframe; only the
    loop-carried stage (depth re-projection, ConvLSTM, decoder) runs keyframe by keyframe, on batch slices of the group's
    encoder outputs.  Every keyframe still gets all of its M + 1 feature passes, its own cost volume and its own encoder pass
    -- nothing is cached or skipped (this is NOT the feature cache of row f1); the state-independent work is merely issued in
    batches the GPU runs far more efficiently than batches of one keyframe (at 8 x 8 .. 64 x 64 maps a single keyframe's
    kernels are one-tile CTAs and launch-bound).  Five streams as in PipelinedFusionnet: trunk head | trunk tail + pyramid |
    plane sweep | encoder | recurrent stage.

    Price: a keyframe's depth is available only after its group is complete (up to `lookahead` - 1 further submits) -- an
    offline / throughput engine, like the reference's run-testing.py loop over a recorded sequence.  submit() buffers; flush()
    (also called by synchronize()) launches an incomplete group.  Same per-sample arithmetic as the other engines; the only
    numerical difference is the split-K decision of a few convolutions, which depends on the batch (as with any batched run:
    <= 1 ulp of fp32 with 3-term operands, rounding flips of the fp16 operands in 1-term mode); the parity tests hold this
    engine to the same bounds against the oracle.

        eng = LookaheadFusionnet(mods, batch=B, height=H, width=W, n_measurement_frames=M, lookahead=4)
        for frame in stream:  eng.submit(*frame, out=pinned_host_tensor_or_None)
        eng.synchronize()
    """

    n_stages = 5

    def __init__(self, mods, batch, height, width, n_measurement_frames, min_depth=0.25, max_depth=20.0, n_depth_levels=64,
                 device=None, lookahead=4, n_groups=3):
        if lookahead < 1 or n_groups < 2:
            raise ValueError("lookahead >= 1 and n_groups >= 2 required")
        self.mods, self.B, self.H, self.W, self.M = mods, batch, height, width, n_measurement_frames
        self.min_depth, self.max_depth, self.D = min_depth, max_depth, n_depth_levels
        dev = device or next(mods["fe"].parameters()).device
        self.device = dev
        self.T, self.G = int(lookahead), int(n_groups)
        TB = self.T * batch
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        eye = lambda: torch.eye(4, dtype=torch.float32, device=dev).repeat(TB, 1, 1)
        self.groups, self.kslots = [], []
        for g in range(self.G):
            images = z((n_measurement_frames + 1) * TB, 3, height, width)      # [reference block | measurement 1 block | ...], TB rows each
            K0 = torch.tensor([[float(width), 0.0, width / 2.0], [0.0, float(width), height / 2.0], [0.0, 0.0, 1.0]], device=dev).repeat(TB, 1, 1)
            grp = {"images": images, "ref_image": images[:TB],
                   "meas_images": [images[(m + 1) * TB:(m + 2) * TB] for m in range(n_measurement_frames)],
                   "ref_pose": eye(), "full_K": K0, "meas_poses": [eye() for _ in range(n_measurement_frames)],   # sane until overwritten
                   "head": None, "pyramid": None, "swept": None, "enc": None, "graph": [None] * 4,
                   "done": [torch.cuda.Event() for _ in range(4)], "rec_done": torch.cuda.Event()}
            self.groups.append(grp)
            for j in range(self.T):
                lo, hi = j * batch, (j + 1) * batch
                self.kslots.append({"group": g, "lo": lo, "hi": hi, "ref_image": grp["ref_image"][lo:hi],
                                    "meas_images": [mi[lo:hi] for mi in grp["meas_images"]], "ref_pose": grp["ref_pose"][lo:hi],
                                    "full_K": grp["full_K"][lo:hi], "meas_poses": [mp[lo:hi] for mp in grp["meas_poses"]],
                                    "depth": z(batch, height, width), "graph": dict(), "done": torch.cuda.Event()})
        # the recurrent stage's small kernels carry the loop dependence; on a high-priority stream their CTAs are dispatched
        # ahead of the queued CTAs of the batched stages' big grids (DVMVS_LA_PRIO=0: all streams equal)
        import os as _os
        prio = _os.environ.get("DVMVS_LA_PRIO", "1") == "1"
        # measured switches, both off (B200, c2, lookahead 4: 2 396 keyframes/s as is): DVMVS_LA_FOREACH=1 writes the recurrent state
        # back with one multi-tensor copy instead of five copy kernels (2 410: inside run-to-run noise); DVMVS_LA_REC_PDL=1 captures
        # the recurrent stage's graph with programmatic dependent launch (2 321: slower, as for the other stages)
        self._foreach = _os.environ.get("DVMVS_LA_FOREACH", "0") == "1" and hasattr(torch, "_foreach_copy_")
        self._rec_pdl = _os.environ.get("DVMVS_LA_REC_PDL", "0") == "1"
        self.streams = [torch.cuda.Stream(device=dev, priority=(-1 if (prio and i == 4) else 0)) for i in range(5)]
        self.stream_a, self.stream_b = self.streams[0], self.streams[-1]
        self._static_state = None
        self._has_state = False
        self._gi, self._fill = 0, 0          # group counter, keyframes buffered in the open group
        self._pending = []                   # (kslot index, with_state, out) of the open group
        self._kslot_of = {}
        self.t = 0
        self._kernels = [0] * 5
        self.kernels_per_keyframe = 0

    def reset(self):
        """New clip / tracking lost: the next submitted keyframe starts without recurrent state (buffered keyframes keep theirs)."""
        self._has_state = False

    def load_state(self, lstm_state, previous_depth, previous_pose):
        """As PipelinedFusionnet.load_state: continue a clip whose first keyframes ran elsewhere."""
        if self._static_state is None:
            raise RuntimeError("load_state: call prime() (or submit one keyframe) first")
        self.synchronize()
        h, c, pd, pp = self._static_state
        with torch.no_grad():
            h.copy_(lstm_state[0])
            c.copy_(lstm_state[1])
            pd.copy_(previous_depth.reshape(pd.shape))
            pp.copy_(previous_pose)
        torch.cuda.current_stream(self.device).synchronize()
        self._has_state = True

    # -- capture helpers ------------------------------------------------------------------------------------------------
    def _graph_of(self, fn, stream, rec=False):
        """Warm up `fn` twice on `stream`, capture it; returns (graph, result of the captured run, kernels launched)."""
        from . import _native
        torch.cuda.synchronize(self.device)
        saved = [t.clone() for t in self._static_state] if (rec and self._static_state is not None) else None
        _native.lib().dvmvs_set_programmatic_launch(1 if (rec and self._rec_pdl) else 0)   # see PipelinedFusionnet: PDL costs throughput with stages in flight
        try:
            with torch.cuda.stream(stream), torch.no_grad(), no_auto_graph():
                for _ in range(2):
                    res = fn()
            stream.synchronize()
            if rec and self._static_state is None:
                pred, st = res
                self._static_state = (st.lstm_state[0].clone(), st.lstm_state[1].clone(), st.previous_depth.clone(), st.previous_pose.clone())
            g = torch.cuda.CUDAGraph()
            n0 = _native.launch_count()
            with torch.no_grad(), torch.cuda.graph(g, stream=stream):
                res = fn(capturing=True) if rec else fn()
            n = _native.launch_count() - n0
        finally:
            _native.lib().dvmvs_set_programmatic_launch(-1)
        if saved is not None:
            for dst, src in zip(self._static_state, saved):
                dst.copy_(src)
        torch.cuda.synchronize(self.device)
        return g, res, n

    def _rec_fn(self, ks, grp, with_state):
        lo, hi = ks["lo"], ks["hi"]

        def fn(capturing=False):
            st = KeyframeState()
            if with_state:
                h, c, pd, pp = self._static_state
                st.lstm_state, st.previous_depth, st.previous_pose = (h, c), pd, pp
            enc, half_K = grp["enc"]
            view = {"ref_image": ks["ref_image"], "ref_pose": ks["ref_pose"], "full_K": ks["full_K"], "ref_cl": grp["ref_cl"][lo:hi],
                    "lstm_K": grp["lstm_K"][lo:hi]}
            if grp.get("input_gates") is not None:
                view["input_gates"] = grp["input_gates"][lo:hi]
            pred, st = _stage_rec(self.mods, st, view, tuple(ops.batch_slice(e, lo, hi) for e in enc), half_K[lo:hi])
            if capturing:
                h, c, pd, pp = self._static_state
                dsts = [ks["depth"], h, c, pd, pp]
                srcs = [pred, st.lstm_state[0], st.lstm_state[1], st.previous_depth.reshape(pd.shape), ks["ref_pose"]]
                if self._foreach:        # one multi-tensor launch instead of five copy kernels at the end of the loop-carried chain
                    torch._foreach_copy_(dsts, [s_.reshape(d_.shape) for d_, s_ in zip(dsts, srcs)])
                else:
                    for d_, s_ in zip(dsts, srcs):
                        d_.copy_(s_)
            return pred, st
        return fn

    # -- steady state ---------------------------------------------------------------------------------------------------
    def submit(self, reference_image, reference_pose, measurement_images, measurement_poses, full_K, out=None):
        """Buffer keyframe t (inputs CPU-pinned or CUDA): its inputs are copied now, its stages are launched when its group of
        `lookahead` keyframes is complete (or at flush() / synchronize()).  `out` as in PipelinedFusionnet.submit."""
        g = self._gi % self.G
        grp = self.groups[g]
        ki = g * self.T + self._fill
        ks = self.kslots[ki]
        s0 = self.streams[0]
        caller = torch.cuda.current_stream(self.device)
        s0.wait_stream(caller)
        for t_in in [reference_image, reference_pose, full_K] + list(measurement_images) + list(measurement_poses):
            if t_in.is_cuda:
                t_in.record_stream(s0)
        if out is not None and out.is_cuda:
            out.record_stream(self.streams[4])
        with torch.cuda.stream(s0):
            if self._fill == 0:
                s0.wait_event(grp["rec_done"])           # every stage of this group's previous use has finished reading its buffers
            ks["ref_image"].copy_(reference_image, non_blocking=True)
            ks["ref_pose"].copy_(reference_pose, non_blocking=True)
            ks["full_K"].copy_(full_K, non_blocking=True)
            for dst, src in zip(ks["meas_images"], measurement_images):
                dst.copy_(src, non_blocking=True)
            for dst, src in zip(ks["meas_poses"], measurement_poses):
                dst.copy_(src, non_blocking=True)
        self._pending.append((ki, self._has_state, out))
        self._has_state = True
        self._kslot_of[self.t] = ki
        self._kslot_of.pop(self.t - 4 * self.T * self.G, None)
        self._fill += 1
        self.t += 1
        if self._fill == self.T:
            self.flush()
        return self.t - 1

    def flush(self):
        """Launch the open group (complete or not): trunk, pyramid, plane sweep and encoder over the group's buffers, then the
        recurrent stage for each buffered keyframe in order."""
        if not self._pending:
            return
        grp = self.groups[self._gi % self.G]
        s0, s1, s2, s3, s4 = self.streams
        depth_args = (self.min_depth, self.max_depth, self.D)

        def run(i, stream, fn, key, after):
            with torch.cuda.stream(stream):
                if after is not None:
                    stream.wait_event(after)
                if grp["graph"][i] is None:
                    grp["graph"][i], grp[key], self._kernels[i] = self._graph_of(fn, stream)
                grp["graph"][i].replay()
                grp["done"][i].record(stream)

        def head_fn():
            _stage_side_inputs(grp)                  # channel-last reference images + 1/32 intrinsics for the recurrent stage
            return self.mods["fe"].forward_head(grp["images"])

        run(0, s0, head_fn, "head", None)
        run(1, s1, lambda: self.mods["fpn"](*self.mods["fe"].forward_tail(grp["head"])), "pyramid", grp["done"][0])
        run(2, s2, lambda: _sweep_from_pyramid(grp, grp["pyramid"], 0, *depth_args), "swept", grp["done"][1])
        run(3, s3, lambda: _stage_enc(self.mods, grp, grp["swept"]), "enc", grp["done"][2])
        for ki, with_state, out in self._pending:
            ks = self.kslots[ki]
            with torch.cuda.stream(s4):
                s4.wait_event(grp["done"][3])
                if with_state not in ks["graph"]:
                    ks["graph"][with_state], _, self._kernels[4] = self._graph_of(self._rec_fn(ks, grp, with_state), s4, rec=True)
                ks["graph"][with_state].replay()
                if out is not None:
                    out.copy_(ks["depth"], non_blocking=True)
                ks["done"].record(s4)
        grp["rec_done"].record(s4)
        self.kernels_per_keyframe = sum(self._kernels[:4]) / float(self.T) + self._kernels[4]      # a full group's share + the recurrent stage
        self._pending = []
        self._fill = 0
        self._gi += 1

    def prime(self, reference_image, reference_pose, measurement_images, measurement_poses, full_K):
        """Captures every graph (both variants of the recurrent stage for the slot the next clip starts in) with
        2 x n_groups x lookahead throw-away keyframes, then resets the clip state."""
        for _ in range(2 * self.G * self.T):
            self.submit(reference_image, reference_pose, measurement_images, measurement_poses, full_K)
        self.synchronize()
        self.reset()

    def depth_of(self, t):
        return self.kslots[self._kslot_of[t]]["depth"]

    def synchronize(self):
        self.flush()
        for s in self.streams:
            s.synchronize()


class OnlineFusionnet:
    """The loop body of fusionnet/run-testing-online.py:103-215 as an object: feed every incoming (pose, image) to push();
    it polls the keyframe buffer (dvmvs.keyframe_buffer.KeyframeBuffer, same selection as the script), and for a new
    keyframe runs one fusionnet keyframe with the selected measurement frames, carrying the recurrent state and resetting
    it on tracking loss (run-testing-online.py:106-113).

    Unlike the script it keys a FeatureCache with the buffer's frame ids: a measurement frame
    that was a reference frame before is neither pre-processed nor pushed through FeatureExtractor + FeatureShrinker
    again.  Only the buffer's very first frame (stored without a prediction, response 0) ever misses.

        online = OnlineFusionnet(mods, full_K, preprocess=lambda raw: <(1,3,H,W) CUDA tensor>)
        depth = online.push(pose_4x4_numpy, raw_image)        # (1,H,W) CUDA tensor, or None when no keyframe was due

    `preprocess` maps whatever the caller stores in the buffer as "image" (a decoded frame, a file name, ...) to the
    network input; it is called for the reference frame and for cache misses only."""

    def __init__(self, mods, full_K, preprocess, n_measurement_frames=None, min_depth=0.25, max_depth=20.0, n_depth_levels=64,
                 buffer=None, cache_capacity=None):
        from .config import Config
        from .keyframe_buffer import KeyframeBuffer
        self.mods, self.preprocess = mods, preprocess
        self.device = next(mods["fe"].parameters()).device
        self.full_K = full_K.to(self.device).reshape(1, 3, 3).float()
        self.M = Config.test_n_measurement_frames if n_measurement_frames is None else int(n_measurement_frames)
        self.depth_args = (min_depth, max_depth, n_depth_levels)
        self.buffer = buffer if buffer is not None else KeyframeBuffer(
            buffer_size=Config.test_keyframe_buffer_size, keyframe_pose_distance=Config.test_keyframe_pose_distance,
            optimal_t_score=Config.test_optimal_t_measure, optimal_R_score=Config.test_optimal_R_measure, store_return_indices=False)
        self.cache = FeatureCache(cache_capacity or self.buffer.buffer.maxlen)
        self.state = KeyframeState()
        self.responses = []

    def _pose(self, pose):
        import numpy as np
        return torch.from_numpy(np.ascontiguousarray(pose, dtype=np.float32)).reshape(1, 4, 4).to(self.device)

    def push(self, pose, image):
        response = self.buffer.try_new_keyframe(pose, image)
        self.responses.append(response)
        if response == 3:                              # tracking lost: forget the recurrent state (run-testing-online.py:109-113)
            self.state.reset()
        if response != 1:
            return None
        reference_id = self.buffer.last_frame_id
        frames, ids = self.buffer.get_best_measurement_frames(self.M, with_ids=True)
        measurement_images = [None if i in self.cache else self.preprocess(f[1]) for f, i in zip(frames, ids)]
        measurement_poses = [self._pose(f[0]) for f in frames]
        with torch.no_grad():
            pred, self.state = keyframe(self.mods, self.state, self.preprocess(image), self._pose(pose), measurement_images,
                                        measurement_poses, self.full_K, *self.depth_args, cache=self.cache,
                                        reference_id=reference_id, measurement_ids=ids)
        return pred
