"""Data-parallel step on real GPUs against the CPU oracle: needs at least two CUDA devices (skipped on a
one-GPU box; the same comparison runs inside every multi-GPU `bench.py` as `dp_check`).

Two processes, one GPU each, run captured TrainSteps on different shards from identical weights.  The oracle runs the
reference arithmetic on each shard with the same weights, averages the gradients (per-rank BatchNorm statistics: torch-DDP
semantics) and applies torch's Adam.  Both exchange forms are checked: the fused reduce-scatter + Adam + all-gather kernel
over NVLink peer memory and the NCCL allreduce form (AE_B200_DP_FUSED=0)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ae_b200
from oracle import seeded, torch_port as tp

pytestmark = pytest.mark.gpu

ALPHA, LR, WORLD, BATCH, SEED = 35.0, 1e-3, 2, 32, 9


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, fused, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), AE_B200_DP_FUSED="1" if fused else "0")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    comm = ae_b200.dp.init_communicator()
    model = ae_b200.SupervisedAutoencoder(64, 10).to(dev).train()
    # rank 1 starts from different weights: broadcast_parameters must make the replicas identical (and re-pack them)
    model.load_state_dict(seeded.seeded_state(seeded.ae_state_shapes(64, 10), SEED + 100 * rank))
    model.engine().prepare(dev, BATCH)
    ae_b200.dp.broadcast_parameters(model, src=0)
    opt = ae_b200.Adam(model.parameters(), lr=LR)
    step = ae_b200.TrainStep(model, opt, ALPHA, BATCH, comm=comm)
    assert (getattr(model.engine().flat, "_dp_flags", None) is not None) == fused
    x, y = seeded.seeded_images(WORLD * BATCH, SEED), seeded.seeded_labels(WORLD * BATCH, SEED)
    lo, hi = ae_b200.dp.shard_bounds(WORLD * BATCH, rank, WORLD)
    loss = step(x[lo:hi].to(dev), y[lo:hi].to(dev)).clone()
    torch.cuda.synchronize()
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    gathered = [None] * WORLD
    dist.all_gather_object(gathered, {k: v for k, v in state.items() if "running" not in k and "num_batches" not in k})
    if rank == 0:
        out.put((float(loss[0]), {k: v.numpy().copy() for k, v in state.items()},
                 all(all(torch.equal(gathered[0][k], g[k]) for k in gathered[0]) for g in gathered)))
    step.close()
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("fused", [True, False])
def test_two_gpu_step_matches_the_oracle(fused):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, fused, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    loss0, after, replicas_equal = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert replicas_equal, "parameters differ between the ranks after the step"
    # oracle: per-shard gradients with rank 0's weights, averaged, torch's Adam (first step)
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), SEED)
    x, y = seeded.seeded_images(WORLD * BATCH, SEED), seeded.seeded_labels(WORLD * BATCH, SEED)
    keys = tp.param_keys(st)
    grads, losses = None, []
    for r in range(WORLD):
        lo, hi = ae_b200.dp.shard_bounds(WORLD * BATCH, r, WORLD)
        l, _, _, g, _ = tp.ae_train_step({k: v.clone() for k, v in st.items()}, {}, x[lo:hi], y[lo:hi], ALPHA, LR)
        losses.append(float(l))
        grads = {k: g[k] / WORLD for k in keys} if grads is None else {k: grads[k] + g[k] / WORLD for k in keys}
    assert abs(loss0 - losses[0]) <= 1e-4 * abs(losses[0])
    # the first Adam step moves every weight by lr * sign(g) (|g| >> eps): all of them within 2 lr of the oracle's update, and
    # all but the entries whose averaged gradient is ~0 in the same direction
    worst_frac = 0.0
    for k in keys:
        if grads[k].abs().max() < 1e-5:          # biases in front of a training-mode BatchNorm: exactly zero here
            continue
        m = grads[k] * (1 - 0.9)
        v = grads[k] * grads[k] * (1 - 0.999)
        pred = st[k] - (LR / (1 - 0.9)) * m / (v.sqrt() / (1 - 0.999) ** 0.5 + 1e-8)
        d = (torch.from_numpy(after[k]) - pred).abs()
        assert float(d.max()) <= 2.01 * LR, k
        worst_frac = max(worst_frac, float((d > 0.1 * LR).float().mean()))
    assert worst_frac <= 2e-2, worst_frac
