"""Golden vectors for the Gaussian-FoV / head-direction tile builders, made by EXECUTING the reference's own functions.

Run here (the build container), never on the GPU box:

    python tests/golden/make_gaussian_fov_golden.py

``mycode/data_generator_gaussian_FoV.py`` cannot be imported (it pulls mycode.config -> easydict and a saliency
module), so the function definitions are taken out of the file with ``ast`` and executed unmodified:
``crop_FoV_from_equirect`` (:57-118), ``get_gaussian_FoV`` (:121-127), ``blur_head_direction_equirect`` (:163-224),
``get_head_direction`` (:226-232), ``get_theta_phi_array_per_user`` (:34-55), the two ``*_giventhetaphi`` wrappers
(:130-138, :235-243), ``heatmap_sum`` / ``normalize_to_distribution`` (:246-261) and ``xyz2thetaphi``
(dataIO.py:77-82).  Only outputs are stored (``reference_gaussian_fov_golden.npz``).

One shim: ``crop_FoV_from_equirect`` caps ``longitude`` with the FLOAT ``img_w/2-1`` and then slices with it, which
the NumPy of the reference's day (Python 2, NumPy <= 1.11) accepted by truncating to int; today's NumPy raises.  The
arrays handed to the reference are an ndarray subclass that truncates float slice bounds the same way.
"""
import ast
import math
import os

import numpy as np

REF = "/root/reference/mycode/data_generator_gaussian_FoV.py"
WANT = ["crop_FoV_from_equirect", "get_gaussian_FoV", "blur_head_direction_equirect", "get_head_direction",
        "get_theta_phi_array_per_user", "get_gaussianFoV_per_vid_per_target_giventhetaphi",
        "get_headdirection_per_vid_per_target_giventhetaphi", "heatmap_sum", "normalize_to_distribution"]


class _FloatSliceArray(np.ndarray):
    """ndarray whose slices accept float bounds (truncated), like NumPy <= 1.11."""

    @staticmethod
    def _fix(key):
        def one(k):
            if isinstance(k, slice):
                f = lambda v: v if v is None or isinstance(v, (int, np.integer)) else int(v)
                return slice(f(k.start), f(k.stop), f(k.step))
            return k
        return tuple(one(k) for k in key) if isinstance(key, tuple) else one(key)

    def __getitem__(self, key):
        return super().__getitem__(self._fix(key))

    def __setitem__(self, key, value):
        super().__setitem__(self._fix(key), value)


class _NpShim:
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def ones(shape, *a, **k):
        return np.ones(shape, *a, **k).view(_FloatSliceArray)


def load():
    ns = {"np": _NpShim(), "math": math, "fps": 30}
    for node in ast.parse(open(REF).read()).body:
        if isinstance(node, ast.FunctionDef) and node.name in WANT:
            exec(compile(ast.Module([node], type_ignores=[]), REF, "exec"), ns)
    for node in ast.parse(open("/root/reference/mycode/dataIO.py").read()).body:
        if isinstance(node, ast.FunctionDef) and node.name == "xyz2thetaphi":
            exec(compile(ast.Module([node], type_ignores=[]), "dataIO.py", "exec"), ns)
    return ns


def main():
    ns = load()
    rng = np.random.default_rng(2024)
    out = {}
    # frame centres [phi, theta] in [0, 1]: the equator, both poles, the theta seam on either side, exact tile
    # boundaries, then random ones; 2 viewers x 2 seconds x 30 frames
    pt = rng.uniform(0, 1, (2, 60, 2))
    pt[0, :12] = [[0.5, 0.5], [0.5, 0.0], [0.5, 0.999], [0.0, 0.3], [0.999, 0.7], [0.25, 0.3], [0.75, 0.29],
                  [0.1, 0.02], [0.9, 0.98], [0.5, 0.3], [0.5, 0.7], [0.36, 0.71]]
    out["pt"] = pt
    out["fov"] = np.asarray(ns["get_gaussianFoV_per_vid_per_target_giventhetaphi"](pt.copy()))
    out["head"] = np.asarray(ns["get_headdirection_per_vid_per_target_giventhetaphi"](pt.copy()))
    # only frames whose centre wraps (no 1.00001 peak in the batch): the normaliser is a wrapped frame's own peak
    ptw = np.stack([rng.uniform(0.3, 0.7, 30), rng.uniform(0.0, 0.25, 30)], axis=-1)[None]
    out["pt_wrap"] = ptw
    out["fov_wrap"] = np.asarray(ns["get_gaussianFoV_per_vid_per_target_giventhetaphi"](ptw.copy()))
    out["head_wrap"] = np.asarray(ns["get_headdirection_per_vid_per_target_giventhetaphi"](ptw.copy()))
    # from xyz: get_theta_phi_array's target-viewer branch (:30-32) on float32-valued unit vectors
    v = rng.normal(size=(3, 2, 30, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    v = v.astype(np.float32).astype(np.float64)
    db = v.reshape(3, 2, 90)                                       # (chunks, seconds, interleaved xyz)
    theta, phi = ns["xyz2thetaphi"](db[:, :, 0::3], db[:, :, 1::3], db[:, :, 2::3])
    lab = ns["get_theta_phi_array_per_user"](theta, phi)            # (chunks, seconds * fps, 2)
    out["xyz"], out["xyz_phi_theta"] = v, lab
    # per-second sum of the frame channels, normalised to a distribution
    out["fov_sum"] = np.asarray(ns["heatmap_sum"](np.asarray(out["fov"]).copy()))
    out["head_sum"] = np.asarray(ns["heatmap_sum"](np.asarray(out["head"]).copy()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_gaussian_fov_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (a.shape, a.dtype) for k, a in out.items()})


if __name__ == "__main__":
    main()
