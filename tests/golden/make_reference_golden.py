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
``generate_fake_batch_numpy`` (:73-80).
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
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_numpy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
