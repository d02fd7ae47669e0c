"""Data-parallel host logic on CPU with the gloo backend, world_size 2.

The N-rank semantics the GPU path implements -- every rank runs the step on its own contiguous batch shard with
per-rank BatchNorm statistics, gradients are summed with ONE allreduce of the flat buffer and scaled by 1/world
inside Adam -- are checked against a single-process emulation of the same sharding built from the oracle.
The product's own helpers (dp.shard_bounds, dp.broadcast_parameters, the flat parameter order) are the ones used.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ae_b200
from oracle import seeded, torch_port as tp

ALPHA, LR, WORLD, BATCH, SEED = 35.0, 5e-3, 2, 10, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat(grads, keys):
    return torch.cat([grads[k].reshape(-1) for k in keys])


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(1)
    # rank 1 starts from different weights: broadcast_parameters must make the replicas identical
    model = ae_b200.SupervisedAutoencoder(64, 10)
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), SEED + 100 * rank)
    model.load_state_dict(st)
    ae_b200.dp.broadcast_parameters(model, src=0)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    x, y = seeded.seeded_images(BATCH, SEED), seeded.seeded_labels(BATCH, SEED)
    lo, hi = ae_b200.dp.shard_bounds(BATCH, rank, WORLD)
    keys = [k for k, _ in model.named_parameters()]
    _, _, _, grads, _ = tp.ae_train_step({k: v.clone() for k, v in state.items()}, {}, x[lo:hi], y[lo:hi], ALPHA, LR)
    flat = _flat(grads, keys)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)          # the single exchange step
    flat /= WORLD                                        # grad_scale = 1/world (folded into ae_adam_step_flat on the GPU)
    if rank == 0:
        out.put((flat.numpy().copy(), {k: state[k].numpy().copy() for k in keys}))   # by value: the worker exits first
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch_exactly():
    for total in (1, 2, 7, 48, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            if total < world:
                continue
            b = [ae_b200.dp.shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gradient_average_matches_single_process_emulation():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    flat, params0 = out.get(timeout=240)
    flat = torch.from_numpy(flat)
    params0 = {k: torch.from_numpy(v) for k, v in params0.items()}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # emulation: the oracle on each shard with rank 0's weights, gradients averaged
    torch.set_num_threads(1)
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), SEED)
    keys = list(params0)
    for k in keys:                                        # broadcast made every replica equal to rank 0's weights
        assert torch.equal(params0[k], st[k]), k
    x, y = seeded.seeded_images(BATCH, SEED), seeded.seeded_labels(BATCH, SEED)
    acc = None
    for r in range(WORLD):
        lo, hi = ae_b200.dp.shard_bounds(BATCH, r, WORLD)
        _, _, _, g, _ = tp.ae_train_step({k: v.clone() for k, v in st.items()}, {}, x[lo:hi], y[lo:hi], ALPHA, LR)
        f = _flat(g, keys)
        acc = f if acc is None else acc + f
    acc /= WORLD
    assert float((flat - acc).abs().max()) <= 1e-6 * float(acc.abs().max())
    # and it is NOT the gradient of the unsharded batch (BatchNorm statistics are per rank, as under torch DDP)
    _, _, _, gfull, _ = tp.ae_train_step({k: v.clone() for k, v in st.items()}, {}, x, y, ALPHA, LR)
    assert float((_flat(gfull, keys) - acc).abs().max()) > 1e-4 * float(acc.abs().max())
