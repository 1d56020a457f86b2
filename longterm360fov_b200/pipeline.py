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


class M4HeatmapBatches:
    """(U, S, 90) raw seconds of one video -> the samples of the heatmap ConvLSTM (``convlstm_seq2seq``,
    cfg.use_one_hot) in the order mycode/data_generator_for_heatmap.py:103-216 yields them (target viewers outermost,
    windows in time order): inputs [enc (B,10,36,18,30), dec0 = the last observed second (B,1,36,18,30)], target
    (B,10,36,18,30).  The 10-degree one-hot FoV-centre maps (fov_onehot_heatmaps, :32,65-67 / utility.py:533-556) and the
    windowing (fov_window_stacks, utility.py:264-305) run on the GPU: the host sends 360 B per viewer-second instead of
    77 760 B.  ``limit`` keeps the first ``limit`` samples (the reference trains this model at batch 32)."""

    def __init__(self, stride=10, running_length=10, bin_size=10, limit=None):
        self.stride, self.running_length, self.bin_size, self.limit = stride, running_length, bin_size, limit

    def __call__(self, frames):
        U, S = frames.shape[0], frames.shape[1]
        L = self.running_length
        heat = ops.one_hot_heatmaps(frames.reshape(U, S, -1, 3), self.bin_size)              # (U,S,36,18,F)
        past, fut, _ = ops.reshape2second_stacks(heat.view(U, S, -1), collapse_user=False, stride=self.stride,
                                                 running_length=L)                         # (U,n,L,C): viewer-major
        shp = (-1, L) + tuple(heat.shape[2:])
        enc, tgt = past.view(shp), fut.view(shp)
        if self.limit is not None:
            enc, tgt = enc[:self.limit], tgt[:self.limit]
        return [enc, enc[:, -1:]], [tgt]
