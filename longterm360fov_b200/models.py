"""Keras-shaped models over the fov360 kernels.

The reference scripts build a Keras functional graph and then call
``model.compile / fit / predict / save``; the builders here return objects with the
same call shape so a script swaps its ``Model(...)`` for one of these and nothing
else changes.  Graph topologies follow SURVEY.md section 8a':

* ``fov_seq2seq``            - mycode/FoV_seq2seq.py:82-103 (M1) and, with
                               ``num_encoder_tokens=6``, mycode/FoV_seq2seq_mu_var.py:219-248 (M2);
                               ``teacher_forcing=False`` is mycode/FoV_seq2seq_no_teac_forc.py:37-149.
* ``others_lstm_span_whole`` - mycode/others_LSTM_span_whole.py:77-353, concat-state branch (M3).
* ``convlstm_seq2seq``       - mycode/convlstm_seq2seq.py:73-287 (M4).
"""
from __future__ import annotations

import collections
import gc
import math
import os
import weakref

import numpy as np
import torch

from . import _lib, h5lite, ops, parallel

_H5_EXT = (".h5", ".hdf5")


def _is_hdf5(path):
    with open(path, "rb") as fh:
        return fh.read(8) == h5lite.SIG


def _init_weights(kind, **kw):
    """Keras-default initialisers (glorot_uniform / orthogonal / forget-bias 1).
    Pure NumPy on the host; kept here (not imported from oracle/)."""
    from . import initializers as ini
    return getattr(ini, kind)(**kw)


LOSS_ALIASES = {
    "mean_squared_error": "mse", "mse": "mse", "_mse": "mse",
    "likelihood_loss": "nll", "nll": "nll",
    "categorical_crossentropy": "cce", "cce": "cce",
}


class Adam:
    """keras.optimizers.Adam (Keras 2.2 update rule, epsilon outside the bias correction)."""

    def __init__(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=None):
        if clipnorm is not None:
            raise NotImplementedError("clipnorm is not used by the reference's compiled models "
                                      "(mycode/others_LSTM_span_whole.py:351-352 passes the string 'Adam')")
        self.lr, self.beta_1, self.beta_2, self.epsilon = lr, beta_1, beta_2, epsilon
        self.iterations = 0
        self.state = None
        self.slots = ("m", "v")

    def init(self, n, device):
        self.state = (torch.zeros(n, device=device), torch.zeros(n, device=device))

    def step(self, p, g, grad_scale=1.0, grad_div=None):
        self.iterations += 1
        ops.adam_step(p, g, self.state[0], self.state[1], self.iterations, self.lr, self.beta_1,
                      self.beta_2, self.epsilon, grad_scale, grad_div)


class RMSprop:
    """keras.optimizers.RMSprop (Keras 2.2 update rule)."""

    def __init__(self, lr=1e-3, rho=0.9, epsilon=1e-7):
        self.lr, self.rho, self.epsilon = lr, rho, epsilon
        self.iterations = 0
        self.state = None
        self.slots = ("accumulator",)

    def init(self, n, device):
        self.state = (torch.zeros(n, device=device),)

    def step(self, p, g, grad_scale=1.0, grad_div=None):
        self.iterations += 1
        ops.rmsprop_step(p, g, self.state[0], self.lr, self.rho, self.epsilon, grad_scale, grad_div)


def _make_optimizer(opt):
    if isinstance(opt, str):
        name = opt.lower()
        if name == "adam":
            return Adam()
        if name == "rmsprop":
            return RMSprop()
        raise ValueError("unsupported optimizer %r" % opt)
    return opt


class History:
    def __init__(self):
        self.history = collections.defaultdict(list)
        self.epoch = []


class _GraphedStep:
    """One training step (zero grads + forward + losses + BPTT) of fixed input shapes captured in a CUDA graph: a replay
    costs one launch from the host instead of ~50 kernel launches plus the autograd bookkeeping, which is what bounds
    the step at the reference's batch sizes (32-64).  The optimiser step stays outside (its step count and learning
    rate are host-side arguments) and so does the gradient allreduce."""

    def __init__(self, model, xs, ys, seed=None):
        # weak: model -> _graphs -> step -> model would be a cycle, and a model (with its graphs' private memory pools)
        # that only the cyclic collector can free may be freed in the middle of ANOTHER model's capture - a cudaFree
        # there invalidates the capture
        self._model = weakref.ref(model)
        self.seed = seed
        self.sx = [torch.empty_like(t) for t in xs]
        self.sy = [torch.empty_like(t) for t in ys]
        self._load(xs, ys)
        side = torch.cuda.Stream(device=model.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up off the capture: kernel attributes, allocator pool
            for _ in range(2):
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        gc.collect()                                      # nothing left for the collector to free during the capture
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()

    @property
    def model(self):
        return self._model()

    def _load(self, xs, ys):
        for dst, src in zip(self.sx + self.sy, list(xs) + list(ys)):
            dst.copy_(src, non_blocking=True)

    def _fwd_bwd(self):
        m = self.model
        m.gflat.zero_()
        ops.set_math(m.compute)
        outs = m._forward(self.sx, True)
        total = m._loss(outs, self.sy)
        m._backward(total, self.seed)
        return total.detach()

    def run(self, xs, ys):
        self._load(xs, ys)
        self.graph.replay()
        return self.loss


class Model:
    """Common Keras-shaped surface; subclasses provide ``_forward`` and ``weight_order``."""

    weight_order: list = []
    n_inputs = 2
    n_outputs = 1

    def __init__(self, weights, device=None):
        _lib.load()                                   # fail loudly without the CUDA library
        if device is None:
            if not torch.cuda.is_available():
                raise _lib.FovError("no CUDA device: the fov360 models have no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        sizes = [(k, tuple(np.shape(weights[k]))) for k in self.weight_order]
        # 64-float (256 B) alignment of every tensor inside the flat buckets
        self._offsets, off = {}, 0
        for k, shp in sizes:
            self._offsets[k] = (off, shp)
            off += (int(np.prod(shp)) + 63) // 64 * 64
        # + 64 reserved floats: the last element of the gradient bucket carries the sample count of a
        # data-parallel step through the allreduce (parallel.allreduce_gradients)
        off += 64
        self.n_flat = off
        self.flat = torch.zeros(off, device=self.device)
        self.gflat = torch.zeros(off, device=self.device)
        self.params, self.grads = collections.OrderedDict(), collections.OrderedDict()
        for k, (o, shp) in self._offsets.items():
            n = int(np.prod(shp))
            self.params[k] = self.flat[o:o + n].view(shp).requires_grad_(True)
            self.grads[k] = self.gflat[o:o + n].view(shp)
        self.set_weights([weights[k] for k in self.weight_order])
        self.optimizer = None
        self.loss_kinds = None
        self.loss_weights = None
        self.stop_training = False
        self.comm = None
        self.world_size = 1
        self._seeds = {}
        self.running_length = 10
        # arithmetic of the conv / dense / ConvLSTM kernels (ops.set_math): tensor cores with 2 bf16
        # terms per operand by default (~16 mantissa bits, forward error ~1e-5); "fp32" selects the CUDA-core kernels
        self.compute = os.environ.get("FOV_COMPUTE", "bf16x2")
        self._graphs, self._use_graphs = {}, False
        # weight gradients on a side stream (ops._WgFork): None = always when the step is replayed from a CUDA graph,
        # eagerly from wgrad_side_stream_min_eager_batch sequences; True / False = always / never
        # fit / fit_generator read the loss of a step this many steps late (FOV_READBACK_DEPTH overrides)
        self.loss_readback_depth = int(os.environ.get("FOV_READBACK_DEPTH", "2"))
        self.layer_wavefront = None       # training steps: None = with CUDA graphs only, True / False = always / never
        self.wgrad_side_stream = None
        self.wgrad_side_stream_min_eager_batch = 384

    def enable_cuda_graphs(self, on=True):
        """Replay the training step from a CUDA graph (captured per input-shape set on first use).  Worth it when the
        step is launch bound (small batches); not available with ConvLSTM input dropout (fresh masks every step)."""
        if on and getattr(self, "dropout", 0.0):
            raise NotImplementedError("CUDA-graph replay needs a mask-free step (dropout=0)")
        self._use_graphs = bool(on)
        if not on:
            self._graphs = {}
        return self

    def set_compute(self, mode):
        """'fp32' | 'bf16' | 'bf16x2' | 'bf16x3' (see _lib.MATH)."""
        if mode not in _lib.MATH:
            raise ValueError("compute must be one of %s" % sorted(_lib.MATH))
        self.compute = mode
        return self

    # ------------------------------------------------------------------ #
    # weights
    # ------------------------------------------------------------------ #
    def get_weights(self):
        return [self.params[k].detach().cpu().numpy().copy() for k in self.weight_order]

    def set_weights(self, arrays):
        assert len(arrays) == len(self.weight_order)
        with torch.no_grad():
            for k, a in zip(self.weight_order, arrays):
                t = torch.as_tensor(np.asarray(a, dtype=np.float32))
                assert tuple(t.shape) == self._offsets[k][1], (k, t.shape, self._offsets[k][1])
                self.params[k].copy_(t.to(self.device))

    def get_weights_dict(self):
        return {k: v for k, v in zip(self.weight_order, self.get_weights())}

    def _keras_layers(self):
        """[(layer, [(layer/weight:0, array), ...]), ...]: ``weight_order`` grouped the way Keras' saving.py stores a
        model (one group per layer, weights in ``layer.weights`` order, TensorFlow-style ``:0`` names)."""
        layers = collections.OrderedDict()
        for k, a in self.get_weights_dict().items():
            layer = k.split("/", 1)[0]
            for half in ("_fwd", "_bwd"):                 # Keras keeps both halves of a Bidirectional in ONE layer group
                if layer.endswith(half):                  # (forward weights first, then backward)
                    layer = layer[:-len(half)]
            layers.setdefault(layer, []).append((k + ":0", a))
        return list(layers.items())

    def save_weights(self, path):
        """Keras-layout weights.  ``*.h5`` / ``*.hdf5``: an HDF5 file in Keras 2.2's ``save_weights`` layout
        (``layer_names`` / ``weight_names`` attributes, one group per layer; the names the reference's checkpoints
        use, mycode/FoV_seq2seq.py:108) written by ``h5lite``; anything else: ``.npz`` keyed by layer / weight name."""
        if path.endswith(_H5_EXT):
            h5lite.write_keras_weights(path, self._keras_layers())
            return
        if not path.endswith(".npz"):
            path = path + ".npz"
        np.savez(path, **{k.replace("/", "__"): v for k, v in self.get_weights_dict().items()})

    def save(self, path):
        """Keras ``model.save`` (mycode/FoV_seq2seq_mu_var.py:256): for ``*.h5`` the weights under ``/model_weights`` plus -
        once compiled - the optimiser's iteration count and moment buffers under ``/optimizer_weights`` and the
        ``training_config`` attribute, so that ``load_weights`` + ``load_optimizer_weights`` resume a run exactly;
        other paths: ``save_weights``."""
        if not path.endswith(_H5_EXT):
            return self.save_weights(path)
        opt_w, cfg = None, None
        if self.optimizer is not None and self.optimizer.state is not None:
            o = self.optimizer
            name = type(o).__name__
            opt_w = [("%s/iterations:0" % name, np.asarray(o.iterations, np.int64))]
            for slot, flat in zip(o.slots, o.state):
                host = flat.detach().cpu().numpy()
                for k in self.weight_order:
                    off, shp = self._offsets[k]
                    opt_w.append(("training/%s/%s/%s:0" % (name, k, slot), host[off:off + int(np.prod(shp))].reshape(shp)))
            conf = {k: float(getattr(o, k)) for k in ("lr", "beta_1", "beta_2", "rho", "epsilon") if hasattr(o, k)}
            cfg = {"optimizer_config": {"class_name": name, "config": conf}, "loss": list(self.loss_kinds or []),
                   "loss_weights": [float(x) for x in (self.loss_weights or [])], "metrics": [],
                   "sample_weight_mode": None}
        h5lite.write_keras_model(path, self._keras_layers(), opt_w, cfg,
                                 {"class_name": type(self).__name__, "config": {"weight_order": list(self.weight_order)}})

    def load_optimizer_weights(self, path):
        """Restore the optimiser state a ``save('x.h5')`` of the same architecture wrote (iterations + moment buffers);
        the model must be compiled with the same optimiser class.  Returns the file's ``training_config``."""
        cfg, ws = h5lite.read_keras_optimizer(path)
        o = self.optimizer
        if o is None or o.state is None:
            raise ValueError("compile() the model before load_optimizer_weights")
        name = type(o).__name__
        d = dict(ws)
        if "%s/iterations:0" % name not in d:
            raise ValueError("%s holds no %s state (optimizer_weights: %s)" % (path, name, [n for n, _ in ws][:3]))
        for slot, flat in zip(o.slots, o.state):
            host = np.zeros(self.n_flat, np.float32)
            for k in self.weight_order:
                off, shp = self._offsets[k]
                a = d["training/%s/%s/%s:0" % (name, k, slot)]
                if tuple(a.shape) != tuple(shp):
                    raise ValueError("optimizer slot %s of %s: shape %s, expected %s" % (slot, k, a.shape, shp))
                host[off:off + a.size] = a.ravel()
            flat.copy_(torch.from_numpy(host).to(flat.device))
        o.iterations = int(d["%s/iterations:0" % name])
        if cfg:
            for k, v in cfg.get("optimizer_config", {}).get("config", {}).items():
                if hasattr(o, k):
                    setattr(o, k, v)
        return cfg

    def load_weights(self, path, layer_map=None):
        """``.npz`` written by ``save_weights``, or an HDF5 checkpoint (``model.save_weights`` / ``model.save`` /
        ``ModelCheckpoint`` of Keras 2.2, or this class's own ``.h5``).  HDF5 layers are matched BY NAME when the file
        holds this model's layer names (or ``layer_map = {file layer name: this model's layer name}`` names them);
        otherwise - a checkpoint of the reference, whose layers carry Keras' auto names (``lstm_1``, ``dense_2``) - by
        ORDER, as Keras' own ``load_weights`` does: the file's layers that hold weights against this model's layers
        in ``weight_order``, every shape checked; a mismatch raises with both lists."""
        if path.endswith(_H5_EXT) or (os.path.exists(path) and _is_hdf5(path)):
            self.set_weights(self._match_keras_layers(h5lite.read_keras_weights(path), layer_map))
            return
        if not path.endswith(".npz") and not os.path.exists(path):
            path = path + ".npz"
        z = np.load(path)
        self.set_weights([z[k.replace("/", "__")] for k in self.weight_order])

    def _match_keras_layers(self, file_layers, layer_map=None):
        mine = self._keras_layers()
        file_layers = [(n, ws) for n, ws in file_layers if ws]
        if layer_map:
            file_layers = [(layer_map.get(n, n), ws) for n, ws in file_layers]
        by_name = dict(file_layers)
        if all(l in by_name for l, _ in mine):
            pairs = [(l, ws, by_name[l]) for l, ws in mine]
        elif len(file_layers) == len(mine):
            pairs = [(l, ws, fws) for (l, ws), (_, fws) in zip(mine, file_layers)]
        else:
            raise ValueError("checkpoint holds %d layers with weights %s, the model %d %s; pass layer_map" % (
                len(file_layers), [n for n, _ in file_layers], len(mine), [l for l, _ in mine]))
        out = {}
        for l, ws, fws in pairs:
            if len(ws) != len(fws):
                raise ValueError("layer %s: %d weights in the model, %d in the checkpoint" % (l, len(ws), len(fws)))
            short = {n.split("/")[-1].split(":")[0]: a for n, a in fws}
            for i, (k, a) in enumerate(ws):
                w = k.split("/", 1)[1].split(":")[0]
                src = short[w] if w in short and len(short) == len(fws) else fws[i][1]
                if tuple(np.shape(src)) != tuple(np.shape(a)):
                    raise ValueError("layer %s weight %s: shape %s in the checkpoint, %s in the model" % (
                        l, w, np.shape(src), np.shape(a)))
                out[k.split(":")[0]] = np.asarray(src, np.float32)
        return [out[k] for k in self.weight_order]

    def count_params(self):
        return sum(int(np.prod(s)) for _, s in self._offsets.values())

    # ------------------------------------------------------------------ #
    # compile / distribute
    # ------------------------------------------------------------------ #
    def compile(self, optimizer="Adam", loss="mean_squared_error", loss_weights=None, metrics=None):
        self.optimizer = _make_optimizer(optimizer)
        self.optimizer.init(self.n_flat, self.device)
        losses = loss if isinstance(loss, (list, tuple)) else [loss] * self.n_outputs
        kinds = []
        for l in losses:
            name = l if isinstance(l, str) else getattr(l, "__name__", str(l))
            if name not in LOSS_ALIASES:
                raise ValueError("unsupported loss %r" % (l,))
            kinds.append(LOSS_ALIASES[name])
        self.loss_kinds = kinds
        self.loss_weights = [1.0] * len(kinds) if loss_weights is None else [float(w) for w in loss_weights]
        return self

    def distribute(self, comm=None):
        """Data-parallel training: batch sharded by rank by the caller, one count-weighted summed allreduce of
        the flat gradient bucket per step.  ``comm``: a parallel.FovComm / parallel.TorchComm; default = the C
        ABI's own NCCL communicator (fov_dp_init ... over NVLink), bootstrapped through the initialised
        torch.distributed default group."""
        if comm is None:
            import torch.distributed as dist
            comm = parallel.FovComm(dist.get_rank(), dist.get_world_size())
        self.comm = comm
        self.world_size = comm.world
        comm.broadcast(self.flat, 0)        # identical start: rank 0's parameters
        return self

    # ------------------------------------------------------------------ #
    # steps
    # ------------------------------------------------------------------ #
    def _to_dev(self, arrs):
        out = []
        for a in arrs:
            if isinstance(a, torch.Tensor):
                t = a
                if t.dtype != torch.float32:
                    t = t.float()
            else:
                t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
            out.append(t.to(self.device, non_blocking=True))
        return out

    @staticmethod
    def _as_list(x):
        return list(x) if isinstance(x, (list, tuple)) else [x]

    def _forward(self, inputs, training):
        raise NotImplementedError

    def _loss(self, outs, ys):
        total = None
        for o, y, kind, w in zip(outs, ys, self.loss_kinds, self.loss_weights):
            l = ops.loss(kind, y, o, w, self.running_length)
            total = l if total is None else total + l
        return total

    def _backward(self, total, seed=None):
        """BPTT of the summed losses with a unit gradient (the fused loss gradients pass through untouched,
        ops.set_unit_loss_grad), the side-stream weight gradients joined, then - data parallel - the small flat gradient
        bucket scaled by the rank's sample count ``seed`` (a (1,) device tensor): the same numbers as back-propagating
        with that seed, without a pass over the (B,T,out) loss gradients."""
        ops.set_unit_loss_grad(True)
        try:
            total.backward()
        finally:
            ops.set_unit_loss_grad(False)
        ops.join_wgrad_stream()
        if seed is not None:
            self.gflat.mul_(seed)

    def train_step_device(self, xs, ys, targets_ready=None):
        """One optimiser step on device tensors; returns the loss as a device tensor
        (no host sync).  DP: each rank's gradient bucket is scaled by its sample count, sum-allreduced with the count
        riding in the last element, and the optimiser divides by the global count (unequal shards weight correctly).
        ``targets_ready``: optional CUDA event after which ``ys`` may be read (their H2D copy
        runs on a side stream while the forward pass computes)."""
        n_local = int(xs[0].shape[0])
        seed = None
        if self.world_size > 1:       # gradients weighted by the local sample count (parallel.py)
            seed = self._seeds.get(n_local)
            if seed is None:
                seed = self._seeds[n_local] = torch.full((1,), float(n_local), device=self.device)
        # small batches: the weight-gradient launches leave the backward chain for a side stream (ops.set_wgrad_side_stream)
        side = self.wgrad_side_stream
        if side is None:
            # measured on B200 (scripts/small_batch_ab.py, scripts/m3_side_stream_ab.py, config 2): replayed from a CUDA
            # graph the fork always pays (B=32 1.36 -> 1.17 ms, B=1110 3.92 -> 3.56 ms); launched eagerly its host cost
            # (events, stream switches) outweighs it below ~400 sequences (B=256 2.10 -> 2.82 ms, B=512 2.64 -> 2.14 ms);
            # at large batches the overlap still returns 1-4 % (B=2220 6.04 -> 5.83, B=8880 20.25 -> 19.97 ms)
            lo = 0 if self._use_graphs else self.wgrad_side_stream_min_eager_batch
            side = n_local >= lo
        ops.set_wgrad_side_stream(side)
        # layer wavefront of stacked ConvLSTMs (ops._wave_groups: batches whose whole stack fits on the SMs at once):
        # replayed from a graph B=32 1.19 -> 0.94 ms; launched eagerly a training step of that size is bound by the host,
        # and the extra stream bookkeeping costs more than the overlap returns (1.49 -> 1.89 ms) - graphs only
        ops.set_layer_wavefront(self._use_graphs if self.layer_wavefront is None else self.layer_wavefront)
        try:
            return self._train_step_device(xs, ys, targets_ready, n_local, seed, bool(side))
        finally:
            ops.set_wgrad_side_stream(False)
            ops.set_layer_wavefront(True)

    def _train_step_device(self, xs, ys, targets_ready, n_local, seed, side):
        if self._use_graphs:
            if targets_ready is not None:
                torch.cuda.current_stream().wait_event(targets_ready)
            key = (self.compute, self.world_size > 1, side) + tuple(tuple(t.shape) for t in list(xs) + list(ys))
            step = self._graphs.get(key)
            if step is None:
                step = self._graphs[key] = _GraphedStep(self, xs, ys, seed)
            total = step.run(xs, ys)
        else:
            self.gflat.zero_()
            ops.set_math(self.compute)
            outs = self._forward(xs, True)
            if targets_ready is not None:
                torch.cuda.current_stream().wait_event(targets_ready)
            total = self._loss(outs, ys)
            self._backward(total, seed)
        div = None
        if self.world_size > 1:
            div = parallel.allreduce_gradients(self.gflat, self.comm, n_local)
        with torch.no_grad():
            self.optimizer.step(self.flat, self.gflat, 1.0, div)
        return total.detach() if not self._use_graphs else total.clone()

    def train_on_batch(self, x, y):
        """keras Model.train_on_batch.  The inputs are copied on the compute stream; the targets - only needed by the
        loss at the end of the forward pass - are copied on a side stream, so their H2D transfer overlaps the forward."""
        xs = self._to_dev(self._as_list(x))
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._copy_stream):
            ys = self._to_dev(self._as_list(y))
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        for t in ys:
            t.record_stream(main)
        return float(self.train_step_device(xs, ys, ready).item())

    def _prefetch(self, x, y):
        """Host -> device copy of one batch on the side stream; returns (xs, ys, (inputs event, targets event), batch size)."""
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._copy_stream):
            xs = self._to_dev(self._as_list(x))
            x_ready = torch.cuda.Event()
            x_ready.record(self._copy_stream)
            ys = self._to_dev(self._as_list(y))                # only the loss needs them: may land during the forward
            y_ready = torch.cuda.Event()
            y_ready.record(self._copy_stream)
        for t in xs + ys:
            t.record_stream(main)
        return xs, ys, (x_ready, y_ready), len(xs[0])

    def _train_batches(self, batches, builder=None):
        """Input pipeline of fit / fit_generator: while the kernels of step i run, the batch of step i+1 is assembled
        on the host and copied on the side stream (the copy engines are idle during the step); the loss of a step is
        read back once the next step has been enqueued.  Every batch still crosses PCIe inside the loop and every
        loss is read; yields (batch size, loss) in step order.
        ``builder``: the items are RAW host chunks (an array or a list of arrays); ``builder(*device_chunks)`` turns
        them into (inputs, targets) on the GPU (pipeline.py)."""
        it = iter(batches)

        def fetch():
            item = next(it)
            if builder is None:
                return self._prefetch(*item)
            raw = self._as_list(item)
            dev, _, (ready, _), _ = self._prefetch(raw, [])
            return dev, None, (ready, None), None
        try:
            cur = fetch()
        except StopIteration:
            return
        pending = collections.deque()                          # (batch size, loss tensor) of the steps in flight
        depth = max(1, int(self.loss_readback_depth))
        # (A read-back on its own stream behind a per-step event - so that item() does not wait for the tail of the compute
        # stream - was tried: no gain on the raw-chunk path, and the host running further ahead made the 288 MB
        # host-featurised path slower, 20.5 -> 23.9 ms per step.  item() it stays.)
        while cur is not None:
            xs, ys, (x_ready, y_ready), b = cur
            torch.cuda.current_stream().wait_event(x_ready)
            if builder is not None:
                xs, ys = builder(*xs)                          # featuriser + windowing + target/others split on the GPU
                b = len(xs[0])
            loss = self.train_step_device(xs, ys, y_ready)     # asynchronous launches
            try:
                cur = fetch()
            except StopIteration:
                cur = None
            # the loss of step i-depth is read while the steps after it run: the host never drains the GPU between
            # steps, and with depth > 1 a launch-latency spike on the host (8 ranks sharing the host's cores and the
            # driver) is absorbed by the queued steps instead of opening a bubble
            pending.append((b, loss))
            while len(pending) > depth:
                pb, pl = pending.popleft()
                yield pb, float(pl.item())
        while pending:
            pb, pl = pending.popleft()
            yield pb, float(pl.item())

    def test_on_batch(self, x, y):
        xs, ys = self._to_dev(self._as_list(x)), self._to_dev(self._as_list(y))
        with torch.no_grad():
            ops.set_math(self.compute)
            outs = self._forward(xs, False)
            return float(self._loss(outs, ys).item())

    def predict_on_batch(self, x):
        xs = self._to_dev(self._as_list(x))
        with torch.no_grad():
            ops.set_math(self.compute)
            outs = self._forward(xs, False)
        outs = [o.cpu().numpy() for o in outs]
        return outs if self.n_outputs > 1 else outs[0]

    def predict(self, x, batch_size=32, verbose=0):
        xs = self._as_list(x)
        n = len(xs[0])
        chunks = []
        for s in range(0, n, batch_size):
            o = self.predict_on_batch([a[s:s + batch_size] for a in xs])
            chunks.append(o if isinstance(o, list) else [o])
        outs = [np.concatenate([c[i] for c in chunks], axis=0) for i in range(self.n_outputs)]
        return outs if self.n_outputs > 1 else outs[0]

    def evaluate(self, x, y, batch_size=32, verbose=0):
        xs, ys = self._as_list(x), self._as_list(y)
        n = len(xs[0])
        tot = 0.0
        for s in range(0, n, batch_size):
            b = len(xs[0][s:s + batch_size])
            tot += b * self.test_on_batch([a[s:s + batch_size] for a in xs], [a[s:s + batch_size] for a in ys])
        return tot / n

    def fit(self, x, y, batch_size=32, epochs=1, validation_split=0.0, shuffle=True, initial_epoch=0,
            callbacks=None, validation_data=None, verbose=0):
        """keras Model.fit semantics used by the scripts (mycode/FoV_seq2seq.py:112-117):
        the validation set is the LAST ``validation_split`` fraction taken before any
        shuffling, training indices are reshuffled every epoch, the last partial batch
        is kept, the epoch loss is the sample-weighted mean of the batch losses."""
        if self.optimizer is None:
            raise RuntimeError("compile() the model first")
        xs = [np.asarray(a) for a in self._as_list(x)]
        ys = [np.asarray(a) for a in self._as_list(y)]
        n = len(xs[0])
        if validation_data is not None:
            vx, vy = self._as_list(validation_data[0]), self._as_list(validation_data[1])
        elif validation_split and validation_split > 0.0:
            split = int(n * (1.0 - validation_split))
            vx, vy = [a[split:] for a in xs], [a[split:] for a in ys]
            xs, ys = [a[:split] for a in xs], [a[:split] for a in ys]
            n = split
        else:
            vx = vy = None
        hist = History()
        callbacks = callbacks or []
        for cb in callbacks:
            cb.set_model(self)
            cb.on_train_begin()
        self.stop_training = False
        for epoch in range(initial_epoch, epochs):
            idx = np.random.permutation(n) if shuffle else np.arange(n)
            tot = 0.0
            host = (([a[idx[s:s + batch_size]] for a in xs], [a[idx[s:s + batch_size]] for a in ys])
                    for s in range(0, n, batch_size))
            for b, l in self._train_batches(host):
                tot += l * b
            logs = {"loss": tot / n}
            if vx is not None:
                logs["val_loss"] = self.evaluate(vx, vy, batch_size)
            logs["lr"] = self.optimizer.lr
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history[k].append(v)
            if verbose:
                print("Epoch %d/%d - " % (epoch + 1, epochs) + " - ".join("%s: %.6f" % kv for kv in logs.items()))
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        return hist

    def fit_generator(self, generator, steps_per_epoch, epochs=1, validation_data=None,
                      validation_steps=None, callbacks=None, initial_epoch=0, verbose=0, batch_builder=None, **_):
        """keras Model.fit_generator (mycode/convlstm_heatmap.py:415-418): the generator
        yields (inputs, targets) batches forever.  With ``batch_builder`` (pipeline.py) it yields RAW chunks instead
        and the builder makes the batches on the device."""
        hist = History()
        callbacks = callbacks or []
        for cb in callbacks:
            cb.set_model(self)
            cb.on_train_begin()
        self.stop_training = False
        for epoch in range(initial_epoch, epochs):
            tot, cnt = 0.0, 0
            for b, l in self._train_batches((next(generator) for _ in range(steps_per_epoch)), batch_builder):
                tot += b * l
                cnt += b
            logs = {"loss": tot / max(cnt, 1)}
            if validation_data is not None:
                if isinstance(validation_data, tuple):
                    logs["val_loss"] = self.evaluate(validation_data[0], validation_data[1])
                else:
                    vt, vc = 0.0, 0
                    for _ in range(validation_steps or 1):
                        bx, by = next(validation_data)
                        b = len(self._as_list(bx)[0])
                        vt += b * self.test_on_batch(bx, by)
                        vc += b
                    logs["val_loss"] = vt / max(vc, 1)
            logs["lr"] = self.optimizer.lr
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history[k].append(v)
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        return hist

    def _sinks(self, *names):
        return tuple(self.grads[n] for n in names)

    def _dropout_masks(self, rate, B, H, W, cins):
        """keras ConvLSTM2D(dropout=rate) input masks of one stack call: per layer a (4,B,H,W,Cin_l) tensor (one mask per
        gate, constant over the timesteps of the call), inverted-dropout scaling.  Drawn with a per-model torch
        generator (``self.dropout_seed``); Keras' own random stream cannot be reproduced."""
        if not rate:
            return None
        if getattr(self, "_drop_gen", None) is None:
            self._drop_gen = torch.Generator(device=self.device)
            self._drop_gen.manual_seed(int(getattr(self, "dropout_seed", 0)))
        keep = 1.0 - float(rate)
        return [(torch.rand(4, B, H, W, c, device=self.device, generator=self._drop_gen) < keep).float() / keep
                for c in cins]


# --------------------------------------------------------------------------- #
# M1 / M2: target-only fc-LSTM encoder-decoder
# --------------------------------------------------------------------------- #


class _SubModel:
    """encoder_model / decoder_model of mycode/FoV_seq2seq.py:137-148."""

    def __init__(self, fn):
        self._fn = fn

    def predict(self, x, batch_size=32, verbose=0):
        return self._fn(x)

    predict_on_batch = predict


class FovSeq2Seq(Model):
    weight_order = ["encoder/kernel", "encoder/recurrent_kernel", "encoder/bias",
                    "decoder/kernel", "decoder/recurrent_kernel", "decoder/bias",
                    "decoder_dense/kernel", "decoder_dense/bias"]

    def __init__(self, weights, teacher_forcing=True, max_decoder_seq_length=10,
                 decoder_no_init_state=False, recurrent_activation="hard_sigmoid", device=None):
        super().__init__(weights, device)
        self.teacher_forcing = teacher_forcing
        self.T_dec = max_decoder_seq_length
        self.decoder_no_init_state = decoder_no_init_state
        self.rec_act = recurrent_activation
        self.encoder_model = _SubModel(self._encode)
        self.decoder_model = _SubModel(self._decode_step)

    def _w(self):
        p = self.params
        return [p[k] for k in self.weight_order]

    def _lstm_sinks(self):
        g = self.grads
        return {"enc_kernel": g["encoder/kernel"], "enc_recurrent": g["encoder/recurrent_kernel"],
                "enc_bias": g["encoder/bias"], "dec_kernel": g["decoder/kernel"],
                "dec_recurrent": g["decoder/recurrent_kernel"], "dec_bias": g["decoder/bias"],
                "head_kernel": g["decoder_dense/kernel"], "head_bias": g["decoder_dense/bias"]}

    def _forward(self, inputs, training, teacher_forcing=None, steps=None):
        enc, dec = inputs
        tf = self.teacher_forcing if teacher_forcing is None else teacher_forcing
        T_dec = dec.shape[1] if tf else (steps or self.T_dec)
        opts = {"T_dec": T_dec, "teacher_forcing": tf, "head_act": "tanh", "rec_act": self.rec_act,
                "dec_zero_init": self.decoder_no_init_state, "training": training, "need_enc_hseq": False}
        y, _, _ = ops.LSTMSeq2SeqFn.apply(opts, self._lstm_sinks() if training else None, enc, dec, None,
                                          *self._w())
        return [y]

    # --- inference sub-models (mycode/FoV_seq2seq.py:137-178) --- #
    def _encode(self, x):
        (x,) = self._to_dev([x])
        p = self.params
        with torch.no_grad():
            h, c = ops.lstm_states(x, p["encoder/kernel"], p["encoder/recurrent_kernel"], p["encoder/bias"],
                                   self.rec_act)
        return [h.cpu().numpy(), c.cpu().numpy()]

    def _decode_step(self, inputs):
        x, h, c = self._to_dev(inputs)
        p = self.params
        with torch.no_grad():
            y, h, c = ops.lstm_decode_steps(x, h, c, p["decoder/kernel"], p["decoder/recurrent_kernel"],
                                            p["decoder/bias"], p["decoder_dense/kernel"],
                                            p["decoder_dense/bias"], x.shape[1], True, "tanh", self.rec_act)
        return [y.cpu().numpy(), h.cpu().numpy(), c.cpu().numpy()]

    def decode_sequence_fov(self, input_seq, last_mu_var=None, steps=None):
        """Batched form of decode_sequence_fov (mycode/FoV_seq2seq.py:154-178,
        mycode/FoV_seq2seq_mu_var.py:286-311): encoder once, then ``steps`` decoder
        steps feeding the output back, all inside ONE persistent kernel launch.
        ``last_mu_var`` (B,1,6) defaults to the mean/var of the last observed second."""
        xs = self._to_dev([input_seq])[0]
        if last_mu_var is None:
            if xs.shape[-1] == 90:
                last_mu_var = ops.mean_var_xyz(xs[:, -1:, :].contiguous())
            else:
                last_mu_var = xs[:, -1:, :].contiguous()
        else:
            last_mu_var = self._to_dev([last_mu_var])[0]
        with torch.no_grad():
            y = self._forward([xs, last_mu_var], False, teacher_forcing=False, steps=steps)[0]
        return y.cpu().numpy()


def fov_seq2seq(latent_dim=64, num_encoder_tokens=90, num_decoder_tokens=6, max_encoder_seq_length=10,
                max_decoder_seq_length=10, teacher_forcing=True, decoder_no_init_state=False,
                recurrent_activation="hard_sigmoid", weights=None, seed=1, device=None):
    """Builder for M1 (mycode/FoV_seq2seq.py:19-28,82-103).  Inputs
    ``[encoder_inputs (B,T,num_encoder_tokens), decoder_inputs (B,T,6) | (B,1,6)]`` ->
    ``decoder_outputs (B,T,6)``."""
    if weights is None:
        weights = _init_weights("init_fov_seq2seq", seed=seed, num_encoder_tokens=num_encoder_tokens,
                                num_decoder_tokens=num_decoder_tokens, latent_dim=latent_dim)
    return FovSeq2Seq(weights, teacher_forcing, max_decoder_seq_length, decoder_no_init_state,
                      recurrent_activation, device)


def fov_seq2seq_mu_var(latent_dim=64, num_decoder_tokens=6, **kw):
    """Builder for M2 (mycode/FoV_seq2seq_mu_var.py:40-49,219-248): encoder consumes the
    per-second mean/var (B,10,6)."""
    return fov_seq2seq(latent_dim=latent_dim, num_encoder_tokens=6, num_decoder_tokens=num_decoder_tokens, **kw)


# --------------------------------------------------------------------------- #
# sibling models: 2- / 3-layer target-only fc-LSTM encoder-decoders
# --------------------------------------------------------------------------- #

_KH = 64     # hidden width of the fc-LSTM kernels


def _pad_gates(a, units):
    """(..., 4*units) gate-blocked [i|f|c|o] -> (..., 4*64): every gate block padded to 64 columns with zeros."""
    a = np.asarray(a, np.float32)
    out = np.zeros(a.shape[:-1] + (4 * _KH,), np.float32)
    for g in range(4):
        out[..., g * _KH:g * _KH + units] = a[..., g * units:(g + 1) * units]
    return out


def _strip_gates(a, units):
    return np.concatenate([a[..., g * _KH:g * _KH + units] for g in range(4)], axis=-1)


def _pad_rows(a, rows):
    out = np.zeros((rows,) + a.shape[1:], np.float32)
    out[:a.shape[0]] = a
    return out


class StackedFovSeq2Seq(Model):
    """mycode/Fov_seq2seq_2layers.py:232-272 and mycode/3layers.py:223-275: n_layers encoder LSTMs and n_layers decoder
    LSTMs of latent_dim // 2 = 32 units (layer l reads the hidden sequence of layer l-1; decoder layer l starts from
    the state of encoder layer l), Dense(6, tanh) on the last decoder layer, teacher forcing.

    The fc-LSTM kernels are built for 64 hidden units.  A 32-unit LSTM is EXACTLY a 64-unit LSTM whose extra units have
    zero weights and zero bias: their gates sit at hard_sigmoid(0) / tanh(0), so c = h = 0 for ever, they feed
    nothing into the real units (zero recurrent rows, zero rows in the next layer's kernel and in the head) and every
    one of their gradients is exactly 0 - Adam / RMSprop leave them at 0.  The parameters therefore LIVE padded in the
    flat bucket; get_weights / set_weights / save / load speak the Keras shapes.
    One persistent launch per layer runs its encoder and decoder phases; layers above the first have 64-wide inputs
    and take the time-batched projection (xproj_tc.cu) on the tensor-core path."""

    def __init__(self, weights, n_layers=2, share_last_decoder=None, recurrent_activation="hard_sigmoid", device=None):
        self.n_layers = n_layers
        self.share_last_decoder = (n_layers == 3) if share_last_decoder is None else bool(share_last_decoder)
        self.units = int(np.shape(weights["encoder0/recurrent_kernel"])[0])
        if self.units > _KH:
            raise NotImplementedError("fc-LSTM kernels support up to %d units" % _KH)
        self.weight_order = (["%s%d/%s" % (s, l, n) for s in ("encoder", "decoder") for l in range(n_layers)
                              for n in ("kernel", "recurrent_kernel", "bias")] +
                             ["decoder_dense/kernel", "decoder_dense/bias"])
        self._keras_shapes = {k: tuple(np.shape(weights[k])) for k in self.weight_order}
        super().__init__({k: self._pad(k, weights[k]) for k in self.weight_order}, device)
        self.rec_act = recurrent_activation

    def _pad(self, name, a):
        a = np.asarray(a, np.float32)
        u = self.units
        layer, kind = name.split("/")
        if layer == "decoder_dense":
            return _pad_rows(a, _KH) if kind == "kernel" else a
        if kind == "bias":
            return _pad_gates(a, u)
        a = _pad_gates(a, u)
        if kind == "recurrent_kernel" or not layer.endswith("0"):     # rows = hidden units (of this / the layer below)
            a = _pad_rows(a, _KH)
        return a

    def _strip(self, name, a):
        u = self.units
        layer, kind = name.split("/")
        shp = self._keras_shapes[name]
        if layer == "decoder_dense":
            return a[:shp[0]] if kind == "kernel" else a
        a = _strip_gates(a, u)
        return a[:shp[0]] if a.ndim == 2 else a

    def get_weights(self):
        return [self._strip(k, self.params[k].detach().cpu().numpy()) for k in self.weight_order]

    def set_weights(self, arrays):
        arrays = list(arrays)
        if tuple(np.shape(arrays[0])) == self._keras_shapes[self.weight_order[0]] and \
                tuple(np.shape(arrays[1])) == self._keras_shapes[self.weight_order[1]]:
            arrays = [self._pad(k, a) for k, a in zip(self.weight_order, arrays)]
        super().set_weights(arrays)

    def count_params(self):
        return sum(int(np.prod(s)) for s in self._keras_shapes.values())

    def _forward(self, inputs, training):
        enc, dec = inputs
        p, g = self.params, self.grads
        L = self.n_layers
        xe, xd = enc, dec
        y = None
        for l in range(L):
            last = l == L - 1
            pe = "encoder%d" % l
            pd = "decoder%d" % (l - 1 if (self.share_last_decoder and last) else l)     # 3layers.py:266 reuses decoder_lstm2
            opts = {"T_dec": dec.shape[1], "teacher_forcing": True, "head_act": "tanh", "rec_act": self.rec_act,
                    "dec_zero_init": False, "training": training, "need_enc_hseq": not last, "need_dec_hseq": not last}
            sinks = None
            if training:
                sinks = {"enc_kernel": g[pe + "/kernel"], "enc_recurrent": g[pe + "/recurrent_kernel"],
                         "enc_bias": g[pe + "/bias"], "dec_kernel": g[pd + "/kernel"],
                         "dec_recurrent": g[pd + "/recurrent_kernel"], "dec_bias": g[pd + "/bias"]}
                if last:
                    sinks["head_kernel"], sinks["head_bias"] = g["decoder_dense/kernel"], g["decoder_dense/bias"]
            y, xe, xd = ops.LSTMSeq2SeqFn.apply(
                opts, sinks, xe, xd, None, p[pe + "/kernel"], p[pe + "/recurrent_kernel"], p[pe + "/bias"],
                p[pd + "/kernel"], p[pd + "/recurrent_kernel"], p[pd + "/bias"],
                p["decoder_dense/kernel"] if last else None, p["decoder_dense/bias"] if last else None)
        return [y]


def stacked_fov_seq2seq(n_layers=2, latent_dim=64, num_encoder_tokens=6, num_decoder_tokens=6, share_last_decoder=None,
                        recurrent_activation="hard_sigmoid", weights=None, seed=1, device=None):
    """Builder for the 2-layer (mycode/Fov_seq2seq_2layers.py:232-272) and 3-layer (mycode/3layers.py:223-275)
    target-only models: inputs ``[encoder_inputs (B,T,6), decoder_inputs (B,T,6)]`` -> ``(B,T,6)``; every LSTM has
    latent_dim // 2 units.  ``weights`` in Keras shapes (keys encoder{l}/..., decoder{l}/..., decoder_dense/...)."""
    if weights is None:
        weights = _init_weights("init_stacked_fov_seq2seq", seed=seed, n_layers=n_layers,
                                num_encoder_tokens=num_encoder_tokens, num_decoder_tokens=num_decoder_tokens,
                                latent_dim=latent_dim)
    return StackedFovSeq2Seq(weights, n_layers, share_last_decoder, recurrent_activation, device)


class GivenOthersSeq2Seq(Model):
    """mycode/given_others_gt_mean_var_seq2seq.py:97-308: two-layer fc-LSTM encoder-decoder (32 units) that is
    GIVEN the other viewers' ground-truth mean / variance of every future second and mixes them into its own
    prediction; every step's output is the next decoder input (cfg.teacher_forcing = False) or the decoder is
    teacher forced.  Variants (the script's flags): 'mlp_mixing' (default there: Dense(6, tanh) over [others ;
    own Dense(6, tanh) prediction], :277-281), 'others_mlp' (Dense 256 -> 32 relu on the others slice, concatenated
    with the decoder state before decoder_dense, :251-261), 'target_only' (:247-248).

    An fc-LSTM cell is a ConvLSTM2D cell on a 1 x 1 image with 1 x 1 kernels, so the graph runs on the ConvLSTM
    kernels that already carry states and their gradients across one-step calls (as M4's decoder does): the encoder
    is one 2-layer stack call over the 10 past seconds, the re-fed decoder one 2-layer step call per future second.
    LSTM weights live in their Keras shapes in the flat bucket and are handed over as (1,1,in,4H) views."""

    n_inputs, n_outputs = 3, 1

    def __init__(self, weights, variant="mlp_mixing", teacher_forcing=False, recurrent_activation="hard_sigmoid",
                 device=None):
        if variant not in ("mlp_mixing", "others_mlp", "others_lstm", "conv_mixing", "target_only"):
            raise ValueError("variant must be 'mlp_mixing', 'conv_mixing', 'others_mlp', 'others_lstm' or 'target_only'")
        self.variant = variant
        self.teacher_forcing = bool(teacher_forcing)
        order = ["%s%d/%s" % (s, l, n) for s in ("encoder", "decoder") for l in range(2)
                 for n in ("kernel", "recurrent_kernel", "bias")]
        if variant == "others_mlp":
            order += ["others_dense1/kernel", "others_dense1/bias", "others_dense2/kernel", "others_dense2/bias"]
        if variant == "others_lstm":
            order += ["others_bilstm%d_%s/%s" % (l, d, n) for l in range(2) for d in ("fwd", "bwd")
                      for n in ("kernel", "recurrent_kernel", "bias")]
        order += ["decoder_dense/kernel", "decoder_dense/bias"]
        if variant == "mlp_mixing":
            order += ["mixing/kernel", "mixing/bias"]
        if variant == "conv_mixing":
            order += ["mixing_conv%d/%s" % (l, n) for l in range(3) for n in ("kernel", "bias")]
        self.weight_order = order
        super().__init__(weights, device)
        self.rec_act = recurrent_activation
        if variant == "target_only":
            self.n_inputs = 2

    def _cells(self, side, src):
        out = []
        for l in range(2):
            K, R, b = (src["%s%d/%s" % (side, l, n)] for n in ("kernel", "recurrent_kernel", "bias"))
            out.append((K.view(1, 1, K.shape[0], K.shape[1]), R.view(1, 1, R.shape[0], R.shape[1]), b))
        return out

    def _stack(self, side, x, states, training):
        B, T = x.shape[0], x.shape[1]
        return ops.convlstm_stack(x.reshape(B, T, 1, 1, x.shape[-1]), self._cells(side, self.params), states,
                                  self._cells(side, self.grads) if training else None, (1, 1), self.rec_act, training)

    def _one_cell(self, name, x, state, training):
        """One fc-LSTM over (B,T,C) as a 1 x 1 ConvLSTM cell -> ((B,T,H), (hT, cT))."""
        cell = lambda src: [tuple(t.view(1, 1, *t.shape) if t.dim() == 2 else t
                                  for t in (src[name + "/" + n] for n in ("kernel", "recurrent_kernel", "bias")))]
        B, T = x.shape[0], x.shape[1]
        seq, st = ops.convlstm_stack(x.reshape(B, T, 1, 1, x.shape[-1]), cell(self.params),
                                     None if state is None else [state], cell(self.grads) if training else None,
                                     (1, 1), self.rec_act, training)
        return seq[:, :, 0, 0, :], st[0]

    def _others_bilstm(self, oth, training):
        """Two stacked Bidirectional(LSTM, merge_mode='concat') over the others' future
        (mycode/given_others_gt_mean_var_seq2seq.py:151-158): the backward LSTM reads the time-reversed sequence and its
        output is reversed back before the concat; the second pair starts from the first pair's final states (the
        script hands Keras the [seq, h, c, h, c] list of the first layer, which Bidirectional.__call__ splits into
        input + initial_state)."""
        y, st = oth.reshape(oth.shape[0], oth.shape[1], -1), {"fwd": None, "bwd": None}
        for l in range(2):
            f, sf = self._one_cell("others_bilstm%d_fwd" % l, y, st["fwd"], training)
            b, sb = self._one_cell("others_bilstm%d_bwd" % l, torch.flip(y, dims=[1]).contiguous(), st["bwd"], training)
            y, st = torch.cat([f, torch.flip(b, dims=[1])], dim=-1), {"fwd": sf, "bwd": sb}
        return y

    def _head(self, s2, oth_t, training):
        p = self.params
        d = lambda name, x, act: ops.dense(x, p[name + "/kernel"], p[name + "/bias"], act,
                                           self._sinks(name + "/kernel", name + "/bias"), training)
        if self.variant == "target_only":
            return d("decoder_dense", s2, "tanh")
        if self.variant == "others_lstm":                             # oth_t: the bi-LSTM output of this step (B,2H)
            return d("decoder_dense", torch.cat([oth_t, s2], dim=-1), "tanh")
        if self.variant == "conv_mixing":                             # (:190-199,289-295) image (1, 6, num_user)
            pred = d("decoder_dense", s2, "tanh")
            img = torch.cat([oth_t, pred.unsqueeze(1)], dim=1).permute(0, 2, 1).unsqueeze(1).contiguous()
            for l in range(3):
                n = "mixing_conv%d" % l
                img = ops.conv2d(img, p[n + "/kernel"], p[n + "/bias"], "relu", (1, 1),
                                 self._sinks(n + "/kernel", n + "/bias"), training)
            return img[:, 0, :, 0]
        flat = oth_t.reshape(oth_t.shape[0], -1)
        if self.variant == "others_mlp":
            o = d("others_dense2", d("others_dense1", flat, "relu"), "relu")
            return d("decoder_dense", torch.cat([o, s2], dim=-1), "tanh")
        pred = d("decoder_dense", s2, "tanh")
        return d("mixing", torch.cat([flat, pred], dim=-1), "tanh")

    def _forward(self, inputs, training):
        if self.variant == "target_only":
            enc, dec = inputs
            oth = None
            T = dec.shape[1] if self.teacher_forcing else self.running_length
        else:
            enc, oth, dec = inputs
            T = oth.shape[1]
        B, H = enc.shape[0], self.params["encoder0/recurrent_kernel"].shape[0]
        _, states = self._stack("encoder", enc, None, training)
        if self.variant == "others_lstm":
            oth = self._others_bilstm(oth, training)
        outs = []
        if self.teacher_forcing:
            cat, _ = self._stack("decoder", dec, states, training)
            d2 = cat[:, :, 0, 0, H:]                                  # hidden sequence of the second decoder layer
            for t in range(T):
                # the script's others_lstm branch reads get_dim1_layer(decoder2_outputs): the FIRST decoder step for
                # every output step when teacher forced (:238); kept as written
                td = 0 if self.variant == "others_lstm" else t
                outs.append(self._head(d2[:, td].contiguous(), None if oth is None else oth[:, t], training))
        else:
            x = dec[:, 0:1]
            for t in range(T):
                cat, states = self._stack("decoder", x, states, training)
                y = self._head(cat[:, 0, 0, 0, H:].contiguous(), None if oth is None else oth[:, t], training)
                outs.append(y)
                x = y.unsqueeze(1)
        return [torch.stack(outs, dim=1)]


def given_others_gt_mean_var_seq2seq(latent_dim=32, num_user=34, num_encoder_tokens=6, num_decoder_tokens=6,
                                     variant="mlp_mixing", teacher_forcing=False, target_user_only=False,
                                     recurrent_activation="hard_sigmoid", weights=None, seed=1, device=None):
    """Builder for mycode/given_others_gt_mean_var_seq2seq.py:97-308 (cfg.input_mean_var, cfg.predict_mean_var):
    inputs ``[encoder_inputs (B,10,6), others_fut_inputs (B,10,num_user-1,6), decoder_inputs (B,1,6)]`` (``(B,10,6)``
    decoder inputs when teacher forced; no others input with ``target_user_only``) -> ``(B,10,6)``.  ``variant``:
    'mlp_mixing' (the script's flags), 'conv_mixing' (:190-199), 'others_mlp' (:146-149), 'others_lstm' (two stacked
    Bidirectional LSTMs over the others' future, :151-158), 'target_only'."""
    if target_user_only:
        variant = "target_only"
    if weights is None:
        weights = _init_weights("init_given_others_seq2seq", seed=seed, num_user=num_user, latent_dim=latent_dim,
                                num_encoder_tokens=num_encoder_tokens, num_decoder_tokens=num_decoder_tokens,
                                variant=variant)
    return GivenOthersSeq2Seq(weights, variant, teacher_forcing, recurrent_activation, device)


# --------------------------------------------------------------------------- #
# M3: concat-state model with others' whole-span ConvLSTM
# --------------------------------------------------------------------------- #


class OthersLSTMSpanWhole(Model):
    n_inputs, n_outputs = 3, 3
    weight_order = (
        ["oth_convlstm%d/%s" % (l, n) for l in range(3) for n in ("kernel", "recurrent_kernel", "bias")] +
        ["oth_recon_dense/kernel", "oth_recon_dense/bias", "oth_flat_dense/kernel", "oth_flat_dense/bias",
         "encoder/kernel", "encoder/recurrent_kernel", "encoder/bias",
         "decoder/kernel", "decoder/recurrent_kernel", "decoder/bias",
         "encoder_dense/kernel", "encoder_dense/bias", "decoder_dense/kernel", "decoder_dense/bias"])

    def __init__(self, weights, max_encoder_seq_length=10, max_decoder_seq_length=10,
                 recurrent_activation="hard_sigmoid", dropout=0.0, device=None):
        super().__init__(weights, device)
        self.T_enc, self.T_dec = max_encoder_seq_length, max_decoder_seq_length
        self.rec_act = recurrent_activation
        self.dropout = float(dropout)
        self.latent = weights["encoder/recurrent_kernel"].shape[0]
        self._zero_bias = torch.zeros(6, device=self.device)
        self._zero_bias_grad = torch.zeros(6, device=self.device)

    def _forward(self, inputs, training):
        enc_in, oth_in, dec_in = inputs
        p, g = self.params, self.grads
        B, Tall = oth_in.shape[0], oth_in.shape[1]
        Hl = self.latent
        wl = [(p["oth_convlstm%d/kernel" % l], p["oth_convlstm%d/recurrent_kernel" % l],
               p["oth_convlstm%d/bias" % l]) for l in range(3)]
        sl = [(g["oth_convlstm%d/kernel" % l], g["oth_convlstm%d/recurrent_kernel" % l],
               g["oth_convlstm%d/bias" % l]) for l in range(3)] if training else None
        masks = None
        if training and self.dropout:
            masks = self._dropout_masks(self.dropout, B, oth_in.shape[2], oth_in.shape[3],
                                        [oth_in.shape[4]] + [w[1].shape[2] for w in wl[:-1]])
        cat, _ = ops.convlstm_stack(oth_in, wl, None, sl, rec_act=self.rec_act, training=training,
                                    dropout_masks=masks)
        flat = cat.view(B, Tall, -1)
        # reconstruction head on all 20 slices + concat-state Dense256 on the 10 future slices: one Function, the
        # two input gradients accumulate into one buffer and the strided future view is read in place
        r_oth, s = ops.dual_dense(flat, p["oth_recon_dense/kernel"], p["oth_recon_dense/bias"],
                                  p["oth_flat_dense/kernel"], p["oth_flat_dense/bias"], self.T_enc,
                                  self._sinks("oth_recon_dense/kernel", "oth_recon_dense/bias"),
                                  self._sinks("oth_flat_dense/kernel", "oth_flat_dense/bias"), training)
        # decoder_dense(Concat[h_dec, s]) = h_dec . Wd[:H] + (s . Wd[H:] + bd): the second term does
        # not depend on the recurrence, so it is computed for all steps at once and enters the
        # persistent kernel as the additive head term.
        Wd, bd = p["decoder_dense/kernel"], p["decoder_dense/bias"]
        gWd = g["decoder_dense/kernel"]
        e = ops.dense(s, Wd[Hl:], bd, None, (gWd[Hl:], g["decoder_dense/bias"]), training)
        opts = {"T_dec": self.T_dec, "teacher_forcing": False, "head_act": None, "rec_act": self.rec_act,
                "dec_zero_init": False, "training": training}
        sinks = None
        if training:
            sinks = {"enc_kernel": g["encoder/kernel"], "enc_recurrent": g["encoder/recurrent_kernel"],
                     "enc_bias": g["encoder/bias"], "dec_kernel": g["decoder/kernel"],
                     "dec_recurrent": g["decoder/recurrent_kernel"], "dec_bias": g["decoder/bias"],
                     "head_kernel": gWd[:Hl], "head_bias": self._zero_bias_grad}
        y, enc_seq, _ = ops.LSTMSeq2SeqFn.apply(
            opts, sinks, enc_in, dec_in, e, p["encoder/kernel"], p["encoder/recurrent_kernel"],
            p["encoder/bias"], p["decoder/kernel"], p["decoder/recurrent_kernel"], p["decoder/bias"],
            Wd[:Hl], self._zero_bias)
        r_tar = ops.dense(enc_seq, p["encoder_dense/kernel"], p["encoder_dense/bias"], "tanh",
                          self._sinks("encoder_dense/kernel", "encoder_dense/bias"), training)
        return [y, r_oth, r_tar]


def others_lstm_span_whole(latent_dim=64, num_user=34, kernel_size=5, max_encoder_seq_length=10,
                           max_decoder_seq_length=10, oth_filters=(32, 16, 8), flat_dense=256,
                           recurrent_activation="hard_sigmoid", dropout=0.0, weights=None, seed=1, device=None):
    """Builder for M3, the canonical concat-state model (SURVEY.md hazard 2):
    inputs ``[encoder_inputs (B,10,6), encoder_inputs_oth (B,20,1,num_user-1,6),
    decoder_inputs (B,1,6)]`` -> ``[decoder_outputs (B,10,6), decoder_outputs_oth
    (B,20,(num_user-1)*6), encoder_reconstruct_tar (B,10,6)]``
    (mycode/others_LSTM_span_whole.py:348-349)."""
    if weights is None:
        weights = _init_weights("init_others_lstm_span_whole", seed=seed, num_user=num_user,
                                kernel_size=kernel_size, latent_dim=latent_dim, oth_filters=oth_filters,
                                flat_dense=flat_dense)
    return OthersLSTMSpanWhole(weights, max_encoder_seq_length, max_decoder_seq_length,
                               recurrent_activation, dropout, device)


class OthersConvLSTMTarget(Model):
    """The all-ConvLSTM form of mycode/others_LSTM_span_whole.py (use_fclstm_tar = False, :133-199,273-317) in the raw
    xyz layout (cfg.input_mean_var = cfg.predict_mean_var = False, the defaults of mycode/config.py:69-75): the
    target viewer's past runs through its own ConvLSTM2D stack (latent_dim_target = 8 / 4 / 2 filters, kernel (1,5)),
    and every future second a three-layer one-step decoder stack - seeded by the encoder states - reads the channel
    concat [last output (fps,3) ; others' ConvLSTM state of that second (fps,56)] (:273-275); Dense(3) on the
    concatenated decoder states is the output and the next input (:296,317).  Aux heads as in M3:
    Dense((num_user-1)*3) on the others' state of all 20 seconds, Dense(3,tanh) on the target's past state.
    (As shipped the script stops at a NameError, SURVEY.md hazard 1; this is the graph it describes.)
    Runs on the same kernels as M3 / M4: the persistent ConvLSTM kernels for the two 10 / 20-step stacks, one-step calls
    with carried states for the decoder (filters 4 and 2 take the fp32 CUDA-core ConvLSTM kernels)."""

    n_inputs, n_outputs = 3, 3
    weight_order = (
        ["%s_convlstm%d/%s" % (s, l, n) for s in ("oth", "tar_enc", "tar_dec") for l in range(3)
         for n in ("kernel", "recurrent_kernel", "bias")] +
        ["oth_recon_dense/kernel", "oth_recon_dense/bias", "encoder_dense/kernel", "encoder_dense/bias",
         "decoder_dense/kernel", "decoder_dense/bias"])

    def __init__(self, weights, max_encoder_seq_length=10, recurrent_activation="hard_sigmoid", device=None):
        super().__init__(weights, device)
        self.T_enc = max_encoder_seq_length
        self.rec_act = recurrent_activation

    def _stack(self, prefix, x, states, training):
        p, g = self.params, self.grads
        names = ("kernel", "recurrent_kernel", "bias")
        wl = [tuple(p["%s_convlstm%d/%s" % (prefix, l, n)] for n in names) for l in range(3)]
        sl = [tuple(g["%s_convlstm%d/%s" % (prefix, l, n)] for n in names) for l in range(3)] if training else None
        return ops.convlstm_stack(x, wl, states, sl, rec_act=self.rec_act, training=training)

    def _dense(self, name, x, act, training):
        p = self.params
        return ops.dense(x, p[name + "/kernel"], p[name + "/bias"], act,
                         self._sinks(name + "/kernel", name + "/bias"), training)

    def _forward(self, inputs, training):
        enc_in, oth_in, dec_in = inputs
        Tdec = oth_in.shape[1] - self.T_enc
        oth_seq, _ = self._stack("oth", oth_in, None, training)
        r_oth = self._dense("oth_recon_dense", oth_seq, None, training)
        pst, states = self._stack("tar_enc", enc_in, None, training)
        r_tar = self._dense("encoder_dense", pst, "tanh", training)
        x = dec_in
        outs = []
        for t in range(Tdec):
            cat_in = torch.cat([x, oth_seq[:, self.T_enc + t:self.T_enc + t + 1]], dim=-1)
            dstate, states = self._stack("tar_dec", cat_in, states, training)
            y = self._dense("decoder_dense", dstate, None, training)
            outs.append(y)
            x = y
        return [torch.cat(outs, dim=1), r_oth, r_tar]


def others_convlstm_target(num_user=34, kernel_size=5, max_encoder_seq_length=10, oth_filters=(32, 16, 8),
                           tar_filters=(8, 4, 2), recurrent_activation="hard_sigmoid", weights=None, seed=1,
                           device=None):
    """Builder for the all-ConvLSTM form of mycode/others_LSTM_span_whole.py (use_fclstm_tar=False): inputs
    ``[encoder_inputs (B,10,1,fps,3), encoder_inputs_oth (B,20,1,fps,(num_user-1)*3), decoder_inputs (B,1,1,fps,3)]`` ->
    ``[decoder_outputs (B,10,1,fps,3), decoder_outputs_oth (B,20,1,fps,(num_user-1)*3), encoder_reconstruct_tar
    (B,10,1,fps,3)]``."""
    if weights is None:
        weights = _init_weights("init_others_convlstm_target", seed=seed, num_user=num_user, kernel_size=kernel_size,
                                oth_filters=oth_filters, tar_filters=tar_filters)
    return OthersConvLSTMTarget(weights, max_encoder_seq_length, recurrent_activation, device)


# --------------------------------------------------------------------------- #
# M4: ConvLSTM encoder-decoder (heatmap / trajectory forms)
# --------------------------------------------------------------------------- #


class ConvLSTMSeq2Seq(Model):
    def __init__(self, weights, head_kind="conv2d", max_decoder_seq_length=10, dilation_rate=1,
                 recurrent_activation="hard_sigmoid", dropout=0.0, device=None, sample_and_refeed=False,
                 resample_mode="var_as_std", sample_seed=0):
        order = ["%s_convlstm%d/%s" % (s, l, n) for s in ("enc", "dec") for l in range(3)
                 for n in ("kernel", "recurrent_kernel", "bias")]
        if head_kind in ("conv2d", "conv1d"):
            order += ["head_conv%d/%s" % (l, n) for l in range(3) for n in ("kernel", "bias")]
        else:
            order += ["head_dense/kernel", "head_dense/bias"]
        self.weight_order = order
        super().__init__(weights, device)
        self.head_kind = head_kind
        self.T_dec = max_decoder_seq_length
        self.dilation = (dilation_rate, dilation_rate)
        self.rec_act = recurrent_activation
        self.dropout = float(dropout)
        # cfg.sample_and_refeed (mycode/convlstm_seq2seq.py:259-272): the Dense head's (mu, var) is re-sampled into
        # fps frames per axis which become the next decoder input.  The graph passes var as the stddev (:57,
        # 'var_as_std'); decode_sequence_fov_sampling uses the NumPy twin (utility.py:73-80, 'sqrt_floor').
        self.sample_and_refeed = bool(sample_and_refeed)
        if self.sample_and_refeed and head_kind != "dense":
            raise ValueError("sample_and_refeed needs the mean/var Dense head (cfg.predict_mean_var)")
        self.resample_mode = resample_mode
        self.sample_seed = int(sample_seed)
        self.noise_fn = None          # optional callable (step, B) -> (B,fps,3) N(0,1) tensor: explicit noise (parity runs)
        self._draws = 0
        # the heads' weight gradients (42 % of the backward time) overlap the backward-data convolutions of the next
        # decoder step: measured 35.8 -> 33.6 ms per training step at B=32 (scripts/m4_side_stream_ab.py)
        self.wgrad_side_stream = True

    def _noise(self, step, B, fps):
        """Standard-normal draws of one decoder step: the caller's explicit noise, else the in-kernel Philox stream
        keyed by (sample_seed; draw counter, data-parallel rank) - independent of the batch split."""
        if self.noise_fn is not None:
            z = self.noise_fn(step, B)
            z = z if isinstance(z, torch.Tensor) else torch.as_tensor(np.asarray(z, dtype=np.float32))
            return z.to(self.device, torch.float32).reshape(B, fps, 3)
        rank = self.comm.rank if self.comm is not None else 0
        off = (self._draws << 40) + (rank << 32)
        self._draws += 1
        return ops.philox_normal((B, fps, 3), self.sample_seed, off, self.device)

    def _stack(self, side, x, states, training, packs=None):
        p, g = self.params, self.grads
        wl = [(p["%s_convlstm%d/kernel" % (side, l)], p["%s_convlstm%d/recurrent_kernel" % (side, l)],
               p["%s_convlstm%d/bias" % (side, l)]) for l in range(3)]
        sl = [(g["%s_convlstm%d/kernel" % (side, l)], g["%s_convlstm%d/recurrent_kernel" % (side, l)],
               g["%s_convlstm%d/bias" % (side, l)]) for l in range(3)] if training else None
        masks = None
        if training and self.dropout:       # a fresh set of masks per layer call, as Keras draws them
            masks = self._dropout_masks(self.dropout, x.shape[0], x.shape[2], x.shape[3],
                                        [x.shape[4]] + [w[1].shape[2] for w in wl[:-1]])
        return ops.convlstm_stack(x, wl, states, sl, self.dilation, self.rec_act, training, dropout_masks=masks,
                                  pack_cache=packs)

    def _forward(self, inputs, training, resample_mode=None):
        enc_in, dec_in = inputs
        p = self.params
        B = enc_in.shape[0]
        _, states = self._stack("enc", enc_in, None, training)
        x = dec_in[:, 0:1]
        outs = []
        packs = {}          # packed head weights, valid for this pass (forward + its backward): packed once, used T_dec times
        for step in range(self.T_dec):
            cat, states = self._stack("dec", x, states, training, packs)
            d = cat[:, 0]                                              # (B,H,W,56)
            if self.head_kind == "conv2d":
                y = d
                for l in range(3):
                    y = ops.conv2d(y, p["head_conv%d/kernel" % l], p["head_conv%d/bias" % l], "relu", (1, 1),
                                   self._sinks("head_conv%d/kernel" % l, "head_conv%d/bias" % l), training, packs)
                y = ops.SoftmaxFn.apply(y)
                x = y.unsqueeze(1)
            elif self.head_kind == "conv1d":
                y = d[:, 0]
                for l in range(3):
                    y = ops.conv1d(y, p["head_conv%d/kernel" % l], p["head_conv%d/bias" % l],
                                   "relu" if l < 2 else None,
                                   self._sinks("head_conv%d/kernel" % l, "head_conv%d/bias" % l), training)
                y = ops.SoftmaxFn.apply(y)                             # Conv1D(activation='softmax'), :188-189
                y = y.unsqueeze(1)
                x = y.unsqueeze(1)
            else:
                y = ops.dense(d[:, 0].reshape(B, -1), p["head_dense/kernel"], p["head_dense/bias"], None,
                              self._sinks("head_dense/kernel", "head_dense/bias"), training)
                if self.sample_and_refeed:
                    fps = enc_in.shape[3]
                    fr = ops.gauss_resample(y, self._noise(step, B, fps), resample_mode or self.resample_mode)
                    x = fr.view(B, 1, 1, fps, 3)
                else:
                    x = y.view(B, 1, 1, 1, -1)
            outs.append(y)
        return [torch.stack(outs, dim=1)]

    def decode_sequence_fov_sampling(self, input_seq, last_location=None):
        """mycode/convlstm_seq2seq.py:479-502: encoder once, then max_decoder_seq_length decoder steps, each
        re-sampling fps frames from the predicted (mu, var) with the NumPy rule (negative variances floored to 1e-3,
        std = sqrt(var)) as the next input.  One batched pass; returns (B, steps, 6)."""
        if not self.sample_and_refeed:
            raise ValueError("built without sample_and_refeed")
        xs = self._to_dev([input_seq])[0]
        last = xs[:, -1:].contiguous() if last_location is None else self._to_dev([last_location])[0]
        with torch.no_grad():
            ops.set_math(self.compute)
            y = self._forward([xs, last], False, resample_mode="sqrt_floor")[0]
        return y.cpu().numpy()


def convlstm_seq2seq(latent_dim=16, kernel_size=5, dilation_rate=1, use_one_hot=True, input_mean_var=False,
                     predict_mean_var=False, max_decoder_seq_length=10, fps=30, head=(512, 1024, None),
                     recurrent_activation="hard_sigmoid", dropout=0.0, weights=None, seed=1, device=None,
                     sample_and_refeed=None, sample_seed=0):
    """Builder for M4 (mycode/convlstm_seq2seq.py:32-45,73-287).
    heatmap form (``use_one_hot=True``): ``[encoder_inputs (B,10,36,18,fps), decoder_inputs
    (B,1,36,18,fps)]`` -> ``(B,10,36,18,fps)``, heads Conv2D 56->512->1024->fps + channel softmax.
    trajectory form: ``(B,10,1,fps,3)`` images with the Conv1D(k=7) head; with ``predict_mean_var`` a Dense(6)
    mean/var head whose output is either fed back as a ``(B,1,1,1,6)`` image (``input_mean_var``) or - raw xyz
    inputs - re-sampled into fps frames per axis (``sample_and_refeed``, cfg defaults of mycode/config.py:69-75;
    mycode/convlstm_seq2seq.py:259-272).  Raw inputs with a mean/var head NEED the re-sampling step (6 != 3
    channels), so it defaults to on there."""
    filters = (latent_dim * 2, latent_dim, latent_dim // 2)
    if sample_and_refeed is None:
        sample_and_refeed = bool(predict_mean_var and not input_mean_var and not use_one_hot)
    if predict_mean_var and not use_one_hot and not input_mean_var and not sample_and_refeed:
        raise ValueError("predict_mean_var on raw xyz inputs feeds a 6-wide output into 3-channel ConvLSTMs: "
                         "enable sample_and_refeed (as the reference's cfg does) or input_mean_var")
    if sample_and_refeed and (use_one_hot or not predict_mean_var or input_mean_var):
        raise ValueError("sample_and_refeed applies to the raw-xyz trajectory form with predict_mean_var")
    if use_one_hot:
        kind, in_ch, hd = "conv2d", fps, (head[0], head[1], fps)
    elif predict_mean_var:
        kind, in_ch, hd = "dense", 6 if input_mean_var else 3, None
    else:
        kind, in_ch, hd = "conv1d", 3, (head[0], head[1], 3)
    if weights is None:
        flat_dim = sum(filters) * (1 if input_mean_var else fps)
        weights = _init_weights("init_convlstm_seq2seq", seed=seed, in_ch=in_ch, filters=filters,
                                kernel_size=kernel_size, head=hd, head_kind=kind, flat_dim=flat_dim)
    return ConvLSTMSeq2Seq(weights, kind, max_decoder_seq_length, dilation_rate, recurrent_activation, dropout, device,
                           sample_and_refeed=sample_and_refeed, sample_seed=sample_seed)
