"""Keyframe selection of the online drivers (reference dvmvs/keyframe_buffer.py; imported by
fusionnet/run-testing-online.py:10, pairnet/run-testing-online.py:9 and simulate_keyframe_buffer.py:4).  Host-side,
cold, numpy only -- here because the online scripts import it and because it is the natural owner of the FRAME IDS the
measurement-feature cache is keyed by (dvmvs.pipeline.FeatureCache): every frame that enters a
buffer gets a serial number, and `get_best_measurement_frames(..., with_ids=True)` hands it out with the frame.

Same constructor / method signatures, tuple layouts and response codes as the reference:

    try_new_keyframe(pose, image, index=None) ->
        0  first frame after start / tracking loss: stored, nothing to predict yet        (keyframe_buffer.py:31-36)
        1  stored as a new keyframe: predict                                                (:44-50)
        2  pose valid but closer than keyframe_pose_distance to the last keyframe: skipped  (:51-52)
        3  more than 30 consecutive invalid poses and the buffer held frames: cleared ("TRACKING LOST")   (:56-59)
        4  still lost (buffer already empty)                                                (:60-61)
        5  invalid pose, not yet considered lost                                            (:62-63)

tests/test_host_logic.py replays the fixture scene's poses through KeyframeBuffer and regenerates the reference's three
shipped index files (sample-data/indices/keyframe+hololens-dataset+000+nmeas+{1,2,3}) line for line."""
from collections import deque

import numpy as np

from .utils import is_pose_available, pose_distance

FIRST_FRAME, NEW_KEYFRAME, TOO_CLOSE, TRACKING_LOST, STILL_LOST, POSE_MISSING = 0, 1, 2, 3, 4, 5
_LOST_AFTER = 30          # consecutive missing poses ("over a second", keyframe_buffer.py:56)


class _FrameStore:
    """What both buffers share: a bounded FIFO of (pose, image[, index]) tuples -- the layout the scripts unpack
    (run-testing-online.py:148) -- a parallel FIFO of serial frame ids, and the missing-pose counter."""

    def __init__(self, maxlen, with_indices):
        self.buffer = deque([], maxlen=maxlen)
        self._ids = deque([], maxlen=maxlen)
        self._with_indices = bool(with_indices)
        self._missing = 0
        self._serial = 0

    def _check_index(self, index):
        if self._with_indices and index is None:
            raise ValueError("Storing and returning the frame indices is requested in the constructor, but index=None is passed to the function")

    def _push(self, pose, image, index):
        self.buffer.append((pose, image, index) if self._with_indices else (pose, image))
        self._ids.append(self._serial)
        self._serial += 1

    def _pose_missing(self):
        """Bookkeeping for a frame without a usable pose; returns the response code."""
        self._missing += 1
        if self._missing <= _LOST_AFTER:
            return POSE_MISSING
        if len(self.buffer) == 0:
            return STILL_LOST
        self.buffer.clear()
        self._ids.clear()
        return TRACKING_LOST

    @property
    def last_frame_id(self):
        """Serial id of the newest stored frame (the reference frame after a response of 1), or None."""
        return self._ids[-1] if self._ids else None


class KeyframeBuffer(_FrameStore):
    def __init__(self, buffer_size, keyframe_pose_distance, optimal_t_score, optimal_R_score, store_return_indices):
        super().__init__(buffer_size, store_return_indices)
        self.keyframe_pose_distance = keyframe_pose_distance
        self.optimal_t_score = optimal_t_score
        self.optimal_R_score = optimal_R_score

    def calculate_penalty(self, t_score, R_score):
        """keyframe_buffer.py:16-24: squared distance from the preferred baseline / rotation; a baseline SHORTER than
        the optimum costs five times as much as a longer one."""
        dt = t_score - self.optimal_t_score
        t_cost = (5.0 if dt < 0.0 else 1.0) * np.abs(dt) ** 2.0
        return np.abs(R_score - self.optimal_R_score) ** 2.0 + t_cost

    def try_new_keyframe(self, pose, image, index=None):
        self._check_index(index)
        if not is_pose_available(pose):
            return self._pose_missing()
        self._missing = 0
        if len(self.buffer) > 0:
            moved, _, _ = pose_distance(pose, self.buffer[-1][0])
            if moved < self.keyframe_pose_distance:
                return TOO_CLOSE
            self._push(pose, image, index)
            return NEW_KEYFRAME
        self._push(pose, image, index)
        return FIRST_FRAME

    def get_best_measurement_frames(self, n_requested_measurement_frames, with_ids=False):
        """The n stored frames (newest excluded -- it is the reference frame) with the lowest penalty w.r.t. the newest,
        in np.argpartition's order (keyframe_buffer.py:65-89; that order is what the shipped index files record).
        with_ids=True returns (frames, ids): the serial frame ids to key the measurement-feature cache with."""
        frames, ids = list(self.buffer), list(self._ids)
        candidates = len(frames) - 1
        n = min(n_requested_measurement_frames, candidates)
        reference_pose = frames[-1][0]
        penalties = []
        for k in range(candidates):
            _, R_measure, t_measure = pose_distance(reference_pose, frames[k][0])
            penalties.append(self.calculate_penalty(t_measure, R_measure))
        chosen = np.argpartition(penalties, n - 1)[:n]
        picked = [frames[k] for k in chosen]
        return (picked, [ids[k] for k in chosen]) if with_ids else picked


class SimpleBuffer(_FrameStore):
    """keyframe_buffer.py:92-129: every frame with a pose is a keyframe; measurement frames = the previous buffer_size
    frames.  Response codes: 0 first frame, 1 stored, 2 tracking lost (cleared), 3 still lost, 4 pose missing."""

    def __init__(self, buffer_size, store_return_indices):
        super().__init__(buffer_size + 1, store_return_indices)

    def try_new_keyframe(self, pose, image, index=None):
        self._check_index(index)
        if not is_pose_available(pose):
            return {POSE_MISSING: 4, STILL_LOST: 3, TRACKING_LOST: 2}[self._pose_missing()]
        self._missing = 0
        was_empty = len(self.buffer) == 0
        self._push(pose, image, index)
        return 0 if was_empty else 1

    def get_measurement_frames(self, with_ids=False):
        frames, ids = list(self.buffer)[:-1], list(self._ids)[:-1]
        return (frames, ids) if with_ids else frames
