#!/usr/bin/env python
"""Layer wavefront vs sequential layers on the bare ConvLSTM stack op: forward (inference / training mode) and gradients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import ops
from oracle import keras_numpy as kn

torch.manual_seed(0)
w = kn.init_others_lstm_span_whole(seed=1, num_user=34)
dev = torch.device("cuda")
ops.set_math("bf16x2")
for B in [int(a) for a in sys.argv[1:]] or [3, 40]:
    x = torch.rand(B, 20, 1, 33, 6, device=dev) * 2 - 1
    gcat = torch.randn(B, 20, 1, 33, 56, device=dev)
    res = {}
    for wave in (False, True):
        ops.set_layer_wavefront(wave)
        for training in (False, True):
            wl = [tuple(torch.tensor(w["oth_convlstm%d/%s" % (l, n)], device=dev) for n in ("kernel", "recurrent_kernel", "bias")) for l in range(3)]
            sl = [tuple(torch.zeros_like(t) for t in ws_) for ws_ in wl] if training else None
            xx = x.clone().requires_grad_(False)
            for _ in range(3):          # repeat: races show up as run-to-run differences
                cat, states = ops.convlstm_stack(xx, wl, None, sl, training=training)
                if training:
                    for g3 in sl:
                        for t in g3:
                            t.zero_()
                    cat_ = cat if cat.requires_grad else cat.requires_grad_(True)
            torch.cuda.synchronize()
            res[(wave, training)] = (cat.detach().clone(), [h.detach().clone() for st_ in states for h in st_])
    for training in (False, True):
        a, b = res[(False, training)], res[(True, training)]
        print("B=%d training=%d: cat equal %s (max diff %.3e), states equal %s" % (
            B, training, torch.equal(a[0], b[0]), (a[0] - b[0]).abs().max().item(),
            all(torch.equal(p, q) for p, q in zip(a[1], b[1]))))
        if not torch.equal(a[0], b[0]):
            d = (a[0] - b[0]).abs()
            bad = d.amax(dim=(0, 2, 3)).cpu().numpy()      # per (t, channel)
            print("  per-layer max diff by timestep:", [("L%d" % l, np.round(bad[:, lo:hi].max(axis=1), 4).tolist())
                                                         for l, (lo, hi) in enumerate(((0, 32), (32, 48), (48, 56)))])

    # ---- gradients: BPTT wavefront vs sequential
    gres = {}
    for wave in (False, True):
        ops.set_layer_wavefront(wave)
        outs = []
        for rep in range(3):
            wl = [tuple(torch.tensor(w["oth_convlstm%d/%s" % (l, n)], device=dev) for n in ("kernel", "recurrent_kernel", "bias")) for l in range(3)]
            sl = [tuple(torch.zeros_like(t) for t in ws_) for ws_ in wl]
            xx = x.clone().requires_grad_(True)
            cat, states = ops.convlstm_stack(xx, wl, None, sl, training=True)
            (cat * gcat).sum().backward()
            torch.cuda.synchronize()
            outs.append([xx.grad.clone()] + [t.clone() for g3 in sl for t in g3])
        gres[wave] = outs
    names = ["dx"] + ["L%d/%s" % (l, n) for l in range(3) for n in ("gK", "gR", "gb")]
    for i, n in enumerate(names):
        a, b = gres[False][0][i], gres[True][0][i]
        scale = a.abs().max().item() + 1e-12
        rr = max((gres[True][0][i] - gres[True][r][i]).abs().max().item() for r in (1, 2))
        print("  B=%d %-6s wave-vs-seq max diff %.3e (scale %.3e), wave run-to-run %.3e" % (B, n, (a - b).abs().max().item(), scale, rr))
