"""Golden vectors for the small host utilities, made by EXECUTING the reference's own source (run here, never on the
GPU box):  python tests/golden/make_host_utils_golden.py

``rand_sample_ind`` / ``rand_sample`` (mycode/utility.py:575-591) and ``clip_xyz`` (mycode/dataIO.py:16-26) are pulled
out of their files with ``ast`` (the module tops import tensorflow / keras / pickles) and run unmodified; only the
outputs are stored (reference_host_utils_golden.npz)."""
import ast
import os
import random

import numpy as np


def _load(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], type_ignores=[]), path, "exec"), ns)
    return ns


def main():
    ns = _load("/root/reference/mycode/utility.py", ["rand_sample_ind", "rand_sample"], {"np": np, "random": random})
    ns = _load("/root/reference/mycode/dataIO.py", ["clip_xyz"], ns)
    out = {}
    cases = [(11040, 0, 32, 0.1), (1000, 100, 64, 0.2), (517, 17, 10, 0.15)]
    out["cases"] = np.array(cases, np.float64)
    for i, (tot, ntest, bs, vr) in enumerate(cases):
        random.seed(100 + i)
        ind = ns["rand_sample_ind"](tot, ntest, bs, validation_ratio=vr)
        out["ind%d" % i] = np.array(ind, np.int64)
        data = np.random.default_rng(i).standard_normal((tot - ntest, 3)).astype(np.float32)
        out["picked%d" % i] = ns["rand_sample"](data, ind)
    rng = np.random.default_rng(7)
    vids = {k: {a: rng.normal(0, 0.8, (5, 40)) for a in "xyz"} for k in ("v0", 3)}
    out["clip_in"] = np.stack([vids[k][a] for k in ("v0", 3) for a in "xyz"])
    clipped = ns["clip_xyz"]({k: {a: v.copy() for a, v in d.items()} for k, d in vids.items()})
    out["clip_out"] = np.stack([clipped[k][a] for k in ("v0", 3) for a in "xyz"])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "reference_host_utils_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
