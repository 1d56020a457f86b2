"""Generate golden vectors by EXECUTING the reference's own pure-NumPy functions.

Run here (the build container), never on the GPU box:

    python tests/golden/make_reference_golden.py

``mycode/utility.py`` cannot be imported (its module top imports tensorflow, keras,
h5py ...), so the function *definitions* we need are pulled out of the file with
``ast`` and executed unmodified in a namespace that supplies ``np`` and a stub
``cfg`` holding the defaults of ``mycode/config.py``.  Nothing is copied into the
repo: only the outputs are stored (``reference_numpy_golden.npz``).

Functions executed: ``get_gt_target_xyz`` (utility.py:483-500),
``get_gt_target_xyz_oth`` (:505-517), ``reshape2second_stacks`` (:264-305),
``generate_fake_batch_numpy`` (:73-80), ``_create_one_hot`` (:546-556, its Python-2
float shape ``(360/bin_size)`` is cast to int by an ``np.zeros`` shim),
``xyz2thetaphi`` (dataIO.py:77-82), ``get_whole_span``
(others_LSTM_span_whole.py:403-419) and the index statements of the nested
``_theta_phi_index_for_onehot`` (utility.py:520-539; its two ``pickle.dump`` lines
are dropped and ``return theta_index, phi_index`` is appended).
"""
import ast
import os
import types

import numpy as np

REF = "/root/reference/mycode/utility.py"
WANT = ["get_gt_target_xyz", "get_gt_target_xyz_oth", "reshape2second_stacks",
        "generate_fake_batch_numpy"]


def load_reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    cfg = types.SimpleNamespace(running_length=10, data_chunk_stride=10, purelly_testing=False,
                                fps=30)
    ns = {"np": np, "cfg": cfg, "fps": 30}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANT:
            code = compile(ast.Module([node], type_ignores=[]), REF, "exec")
            exec(code, ns)
    return ns


class _NpShim:
    """numpy with ``zeros`` accepting the Python-2 style float shapes of _create_one_hot."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def zeros(shape, *a, **k):
        return np.zeros(tuple(int(v) for v in shape) if isinstance(shape, tuple) else shape, *a, **k)


def load_heatmap_functions():
    """_create_one_hot + the nested _theta_phi_index_for_onehot (utility.py), xyz2thetaphi (dataIO.py),
    get_whole_span (others_LSTM_span_whole.py), all executed from the reference's own source."""
    ns = {"np": _NpShim()}
    dio = ast.parse(open("/root/reference/mycode/dataIO.py").read())
    for node in dio.body:
        if isinstance(node, ast.FunctionDef) and node.name == "xyz2thetaphi":
            exec(compile(ast.Module([node], type_ignores=[]), "dataIO.py", "exec"), ns)
    tree = ast.parse(open(REF).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "_create_one_hot":
            exec(compile(ast.Module([node], type_ignores=[]), REF, "exec"), ns)
        if isinstance(node, ast.FunctionDef) and node.name == "_save_theta_phi_index":
            inner = [n for n in node.body if isinstance(n, ast.FunctionDef)][0]
            body = [st for st in inner.body
                    if not (isinstance(st, ast.Expr) and isinstance(st.value, ast.Call) and
                            "pickle" in ast.unparse(st.value.func))]
            ret = ast.parse("return theta_index, phi_index").body[0]
            inner.body = body + [ret]
            ast.fix_missing_locations(inner)
            exec(compile(ast.Module([inner], type_ignores=[]), REF, "exec"), ns)
    span = ast.parse(open("/root/reference/mycode/others_LSTM_span_whole.py").read())
    for node in span.body:
        if isinstance(node, ast.FunctionDef) and node.name == "get_whole_span":
            exec(compile(ast.Module([node], type_ignores=[]), "others_LSTM_span_whole.py", "exec"), ns)
    return ns


def main():
    ns = load_reference_functions()
    rng = np.random.default_rng(1234)
    out = {}
    y90 = rng.uniform(-1, 1, (3, 10, 90))
    out["xyz90_in"] = y90
    out["xyz90_out"] = ns["get_gt_target_xyz"](y90)
    y4 = rng.uniform(-1, 1, (2, 5, 30, 3))
    out["xyz4_in"] = y4
    out["xyz4_out"] = ns["get_gt_target_xyz"](y4)
    yo = rng.uniform(-1, 1, (2, 3, 33, 30, 3))
    out["oth_in"] = yo
    out["oth_out"] = ns["get_gt_target_xyz_oth"](yo)
    vid = rng.uniform(-1, 1, (3, 33, 6))
    out["stack_in"] = vid
    for stride in (10, 1, 2):
        for collapse in (True, False):
            a, b, c = ns["reshape2second_stacks"](vid.copy(), collapse_user=collapse, stride=stride,
                                                  purelly_testing=False)
            tag = "stack_s%d_c%d" % (stride, int(collapse))
            out[tag + "_past"], out[tag + "_fut"], out[tag + "_futin"] = a, b, c
    # generate_fake_batch_numpy draws from np.random: seed the legacy generator, record
    # the draws by re-running it with the same seed on standard-normal parameters.
    mu = rng.uniform(-1, 1, (6,))
    var = rng.uniform(-0.02, 0.05, (6,))
    np.random.seed(7)
    samp = np.array(ns["generate_fake_batch_numpy"](mu.copy(), var.copy(), 6))   # (6,30)
    np.random.seed(7)
    noise = np.array([np.random.normal(0.0, 1.0, 30) for _ in range(6)])
    out["fake_mu"], out["fake_var"], out["fake_noise"], out["fake_out"] = mu, var, noise, samp
    # ---- added after the cases above (their random draws stay what they were) ----
    vid90 = rng.uniform(-1, 1, (2, 27, 90))
    out["stack90_in"] = vid90
    for stride, testing in ((10, True), (1, True), (5, False)):
        for collapse in (True, False):
            a, b, c = ns["reshape2second_stacks"](vid90.copy(), collapse_user=collapse, stride=stride,
                                                  purelly_testing=testing)
            tag = "stack90_s%d_t%d_c%d" % (stride, int(testing), int(collapse))
            out[tag + "_past"], out[tag + "_fut"], out[tag + "_futin"] = a, b, c
    hs = load_heatmap_functions()
    hs["cfg"] = types.SimpleNamespace(shuffle_data=False)
    hs["num_user"] = 5
    oth = rng.uniform(-1, 1, (7, 10, 4, 6))                       # (N, 10, num_user-1, 6)
    out["span_in"], out["span_out"] = oth, hs["get_whole_span"](oth.copy())
    oth5 = rng.uniform(-1, 1, (3, 10, 4, 30, 3))
    out["span5_in"], out["span5_out"] = oth5, hs["get_whole_span"](oth5.copy())
    # unit vectors incl. the poles, the theta wrap and exact bin edges
    v = rng.normal(size=(3, 4, 30, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    v[0, 0, :6] = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]]
    v[0, 1, 0] = [-1.0, -1e-12, 0.0]
    v[0, 1, 1] = [np.cos(np.deg2rad(40.0)), np.sin(np.deg2rad(40.0)), 0.0]
    v = v.astype(np.float32).astype(np.float64)                   # the device sees float32 inputs
    ti, pj = hs["_theta_phi_index_for_onehot"](v)
    out["onehot_in"], out["onehot_theta"], out["onehot_phi"] = v, ti, pj
    out["onehot_out"] = hs["_create_one_hot"](ti, pj).transpose(0, 1, 3, 4, 2)   # frames as channels
    # hit rate: boundary_cases + get_iou_or_hitrate + bbox_overlaps_hit_rate (baseline_knn_mean.py:48-93), frame mode
    # of _eval_for_seq2seq (:157-160), spans of :131-132
    kn = {"np": np}
    ksrc = "".join(l for l in open("/root/reference/mycode/baseline_knn_mean.py") if not l.startswith("%"))   # IPython magics
    ktree = ast.parse(ksrc)
    for node in ktree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("get_iou_or_hitrate", "bbox_overlaps_hit_rate",
                                                               "boundary_cases"):
            exec(compile(ast.Module([node], type_ignores=[]), "baseline_knn_mean.py", "exec"), kn)
    n = 400
    gt_c = np.stack([rng.uniform(-np.pi, np.pi, n), rng.uniform(0, np.pi, n)])            # (2, n): theta, phi
    pr_c = gt_c + rng.normal(size=(2, n)) * np.array([[0.9], [0.5]])
    pr_c[0] = np.mod(pr_c[0] + np.pi, 2 * np.pi) - np.pi                                 # wraps across +-pi
    pr_c[:, :4] = gt_c[:, :4]                                                             # exact hits
    gt_c = gt_c.astype(np.float32).astype(np.float64)
    pr_c = pr_c.astype(np.float32).astype(np.float64)
    out["hit_pred"], out["hit_gt"] = pr_c.T.copy(), gt_c.T.copy()
    for a in (1.0, 0.75):
        span = np.array([a * 120, a * 120]) / 180.0 * np.pi
        gt_span = np.array([120, 120]) / 180.0 * np.pi
        g2, c2 = kn["boundary_cases"](gt_c.copy(), pr_c.copy())
        res = [kn["get_iou_or_hitrate"](c2[:, i], span, g2[:, i], gt_span)[0] for i in range(n)]
        out["hit_out_a%d" % int(a * 100)] = np.array(res, np.float64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_numpy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
