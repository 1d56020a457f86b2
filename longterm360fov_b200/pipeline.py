"""Device-side input pipeline: the reference prepares its training tensors on the host with NumPy
(get_data -> get_gt_target_xyz -> get_whole_span, mycode/others_LSTM_span_whole.py:403-419,640-668) and feeds them to
``model.fit``.  Here the HOST side of a training step is one raw video chunk - (viewers, seconds, 90) unit-sphere xyz,
3.6 KB per training sequence instead of the 32 KB of featurised tensors, because every viewer is the target once on
the same raw data - and the featuriser, the windowing and the target / others split run on the GPU:

    builder = M3VideoBatches(num_user=34, limit=8880)
    model.fit_generator(raw_video_generator, steps_per_epoch=..., batch_builder=builder)

The generator yields host arrays / pinned tensors ``frames``; ``fit_generator`` copies chunk i+1 on a side stream while
step i computes, then calls ``builder(frames_on_device) -> (inputs, targets)``.
"""
from __future__ import annotations

import torch

from . import ops


class M3VideoBatches:
    """(U, S, 90) raw seconds of one video -> every (target viewer, 20-second window) sequence of the concat-state
    model: inputs [enc (B,10,6), others (B,20,1,num_user-1,6), dec0 (B,1,6)], targets [future (B,10,6), others
    (B,20,(num_user-1)*6), enc].  ``limit`` truncates to a fixed batch size (whole waves of the persistent kernels)."""

    def __init__(self, num_user=34, stride=10, limit=None, draw=None):
        self.num_user, self.stride, self.limit, self.draw = num_user, stride, limit, draw
        self._idx = {}

    def __call__(self, frames):
        U = frames.shape[0]
        key = (U, frames.device)
        idx = self._idx.get(key)
        if idx is None:
            idx = self._idx[key] = torch.as_tensor(ops.others_index(U, self.num_user, self.draw), device=frames.device)
        return ops.m3_batches_from_video(frames, self.num_user, self.stride, self.limit, idx=idx)
