"""Time the sample-builder kernels alone (CUDA events), large enough to stream from HBM."""
import sys

import torch

sys.path.insert(0, ".")
from longterm360fov_b200 import ops  # noqa: E402


def t(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


dev = torch.device("cuda")
for U, S, C, stride in ((48, 3000, 90, 1), (48, 600, 90, 1), (48, 3000, 6, 1)):
    vid = torch.rand(U, S, C, device=dev)
    ms = t(lambda: ops.reshape2second_stacks(vid, collapse_user=False, stride=stride))
    n = S - 10 + 1 - 10
    gb = 3 * n * U * 10 * C * 4 / 1e9
    print("windows U=%d S=%d C=%d: %.3f ms, %.0f GB/s written" % (U, S, C, ms, gb / ms * 1e3))
for rows in (16384, 4096):
    fr = torch.nn.functional.normalize(torch.randn(rows, 30, 3, device=dev), dim=-1)
    ms = t(lambda: ops.one_hot_heatmaps(fr))
    print("one-hot rows=%d: %.3f ms, %.0f GB/s written" % (rows, ms, rows * 36 * 18 * 30 * 4 / 1e9 / ms * 1e3))
x = torch.rand(4096, 10, 33, 30, 3, device=dev)
ms = t(lambda: ops.get_whole_span(x))
print("whole span (4096,10,33,30,3): %.3f ms, %.0f GB/s moved" % (ms, 3 * x.numel() * 4 / 1e9 / ms * 1e3))
# Gaussian-FoV / head-direction tiles of a whole video: 48 viewers x 200 s x 30 fps (compute-bound on float64 exp:
# the peak pass evaluates every painted full-resolution pixel, the tile pass 648 per frame)
pt = torch.rand(48, 6000, 2, device=dev, dtype=torch.float64)
for kind in ("fov", "head"):
    ms = t(lambda: ops.gaussian_fov_tiles(pt, kind), reps=3, warm=1)
    print("gaussian tiles kind=%s 48x200 s: %.3f ms, %.2f M frames/s" % (kind, ms, 48 * 6000 / ms / 1e3))
tiles = ops.gaussian_fov_tiles(pt, "fov")
ms = t(lambda: ops.heatmap_sum(tiles))
print("heatmap_sum (48,200,18,36,30): %.3f ms, %.0f GB/s read" % (ms, tiles.numel() * 4 / 1e9 / ms * 1e3))
