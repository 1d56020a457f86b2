"""Golden vectors of the reference's ``get_data`` (mycode/utility.py:359-446), produced by EXECUTING the reference's own
source: the function definitions are pulled out of utility.py with ``ast`` (the module itself imports tensorflow /
keras / h5py and cannot be imported) and run unmodified against a stub ``cfg`` with the defaults of mycode/config.py
(running_length 10, data_chunk_stride 10, cut_data_head False, time_shift False).  ``np.random.randint`` is wrapped so
that the duplicate-padding indices the reference draws (utility.py:407-412) are recorded next to the outputs.

    python tests/golden/make_get_data_golden.py     # here, never on the GPU box

Only inputs / outputs are stored (reference_get_data_golden.npz)."""
import ast
import os
import types

import numpy as np

REF = "/root/reference/mycode/utility.py"
WANT = ["get_data", "reshape2second_stacks", "cut_head_or_tail_less_than_1sec"]


class _Rand:
    def __init__(self):
        self.drawn = []

    def randint(self, n):
        v = int(np.random.randint(n))
        self.drawn.append(v)
        return v


class _Np:
    """numpy whose ``random.randint`` records its draws."""

    def __init__(self):
        self.random = _Rand()

    def __getattr__(self, name):
        return getattr(np, name)


def load():
    tree = ast.parse(open(REF).read())
    npx = _Np()
    cfg = types.SimpleNamespace(running_length=10, data_chunk_stride=10, purelly_testing=False, fps=30,
                                cut_data_head=False, time_shift=False)
    ns = {"np": npx, "cfg": cfg, "fps": 30, "print": lambda *a, **k: None, "pdb": types.SimpleNamespace(set_trace=lambda: None)}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANT:
            exec(compile(ast.Module([node], type_ignores=[]), REF, "exec"), ns)
    return ns, npx


def main():
    ns, npx = load()
    rng = np.random.default_rng(77)
    # three "videos": 5 viewers x 23.4 s, 3 viewers x 31 s, 7 viewers x 12 s (skipped: < 20 s); frames at 30 fps
    shapes = {"v0": (5, 702), "v1": (3, 930), "v2": (7, 360)}
    datadb = {}
    for k, (u, f) in shapes.items():
        datadb[k] = {c: rng.uniform(-1, 1, (u, f)) for c in "xyz"}
    out = {}
    for k in datadb:
        for c in "xyz":
            out["in_%s_%s" % (k, c)] = datadb[k][c]
    a, b, c = ns["get_data"](datadb, pick_user=False)
    out["all_past"], out["all_fut"], out["all_futin"] = a, b, c
    for num_user in (4, 6):                       # 4: v0 keeps 3 of its 4 others (truncation), v1 pads 2 -> 3 (duplicates)
        np.random.seed(5)
        npx.random.drawn = []
        r = ns["get_data"](datadb, pick_user=True, num_user=num_user)
        for name, arr in zip(("tar_past", "tar_fut", "tar_futin", "oth_past", "oth_fut", "oth_futin"), r):
            out["u%d_%s" % (num_user, name)] = arr
        out["u%d_dups" % num_user] = np.array(npx.random.drawn, np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_get_data_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
