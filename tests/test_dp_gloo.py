"""world_size-2 gloo test (CPU) of the data-parallel host logic: contiguous batch sharding,
one count-weighted summed allreduce of the flat gradient bucket (unequal shards: 9 samples over 2 ranks), identical
optimiser step on every rank == the single-process full-batch step.  The per-shard gradients come from the CPU
oracle (the CUDA path cannot run here); the code under test is longterm360fov_b200.parallel."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from longterm360fov_b200 import parallel
from oracle import keras_numpy as kn
from oracle import keras_torch as kt


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    rng = np.random.default_rng(0)
    enc = rng.uniform(-1, 1, (9, 10, 6)); dec = rng.uniform(-1, 1, (9, 10, 6)); tgt = rng.uniform(-1, 1, (9, 10, 6))
    return enc, dec, tgt


def _flat(d, order):
    return torch.cat([d[k].reshape(-1) for k in order])


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    w = kt.to_torch(kn.init_fov_seq2seq(seed=1, num_encoder_tokens=6))
    order = sorted(w)
    enc, dec, tgt = parallel.shard_batch(list(_data()), rank, world)
    opt = kt.KerasAdam(w)
    comm = parallel.TorchComm()
    assert (comm.rank, comm.world) == (rank, world)
    n_local = len(enc)                                       # 5 on rank 0, 4 on rank 1
    losses = []
    for _ in range(3):
        loss, _, grads = kt.loss_and_grads(kt.fov_seq2seq_forward, w, [torch.tensor(enc), torch.tensor(dec)],
                                           [torch.tensor(tgt)], [kt.mse])
        # the model scales its gradient bucket by n_local after BPTT (Model._backward); the bucket has a reserved tail
        bucket = torch.cat([_flat(grads, order) * n_local, torch.zeros(64, dtype=torch.float64)])
        div = parallel.allreduce_gradients(bucket, comm, n_local)
        assert float(div) == 9.0
        bucket = bucket / div
        off = 0
        g2 = {}
        for k in order:
            n = w[k].numel()
            g2[k] = bucket[off:off + n].view_as(w[k]); off += n
        opt.step(g2)
        losses.append(float(loss) * n_local / 9.0)
    out_q.put((rank, _flat({k: v.detach() for k, v in w.items()}, order).numpy(), losses))
    dist.destroy_process_group()


def test_dp2_equals_full_batch_step():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process full batch
    w = kt.to_torch(kn.init_fov_seq2seq(seed=1, num_encoder_tokens=6))
    order = sorted(w)
    enc, dec, tgt = _data()
    opt = kt.KerasAdam(w)
    full_losses = []
    for _ in range(3):
        loss, _, grads = kt.loss_and_grads(kt.fov_seq2seq_forward, w, [torch.tensor(enc), torch.tensor(dec)],
                                           [torch.tensor(tgt)], [kt.mse])
        opt.step(grads)
        full_losses.append(float(loss))
    ref = _flat({k: v.detach() for k, v in w.items()}, order).numpy()
    np.testing.assert_allclose(res[0][1], res[1][1], atol=0)            # ranks stay identical
    np.testing.assert_allclose(res[0][1], ref, atol=1e-12)              # == full-batch training
    np.testing.assert_allclose(np.sum([res[0][2], res[1][2]], axis=0), full_losses, atol=1e-12)


def test_shard_bounds_cover_and_partition():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
