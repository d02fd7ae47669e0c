"""Grid-search fan-out on CPU with the gloo backend, world_size 2: the sharding of the configurations, the
gather of the per-configuration results and the reference's selection rule (first strict improvement in grid order,
NB:2732 / NB:3536).  The training itself needs a GPU; a deterministic stand-in plays its role here."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import ae_b200
from ae_b200 import search

# validation losses with a tie for the minimum at grid positions 2 and 5: the reference keeps the earlier one
LOSSES = [0.9, 0.7, 0.31, 0.8, 0.5, 0.31, 0.6]
CONFIGS = search.ae_grid([10.0, 35.0, 60.0], [1e-3, 5e-3, 1e-2])[:len(LOSSES)]


def _run_one(i, cfg):
    return {"best_val_loss": LOSSES[i], "epochs": 3 + i, "who": dist.get_rank() if dist.is_initialized() else 0,
            "state": {"tag": f"state-of-{i}", "cfg": cfg}}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    res = search.fan_out(CONFIGS, _run_one, "best_val_loss", "min")
    out.put((rank, res["best_index"], res["best_state"], [r["who"] for r in res["results"]], [r["best_val_loss"] for r in res["results"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_grid_order_and_selection_rule():
    assert search.ae_grid([1, 2], [0.1, 0.2, 0.3]) == [(1, 0.1), (1, 0.2), (1, 0.3), (2, 0.1), (2, 0.2), (2, 0.3)]   # NB:2642-2643
    assert search.assigned(7, 0, 2) == [0, 2, 4, 6] and search.assigned(7, 1, 2) == [1, 3, 5]
    assert sorted(search.assigned(45, 3, 8) + search.assigned(45, 0, 8)) == sorted(set(search.assigned(45, 3, 8) + search.assigned(45, 0, 8)))
    assert sorted(sum((search.assigned(45, r, 8) for r in range(8)), [])) == list(range(45))
    res = [{"v": v} for v in LOSSES]
    assert search.select_best(res, "v", "min") == 2                      # the earlier of the tied minima
    assert search.select_best([{"v": 0.5}, {"v": 0.9}, {"v": 0.9}], "v", "max") == 1
    assert search.select_best([{"v": 0.0}, {"v": 0.0}], "v", "max") == -1    # NB:3450: starts at 0, strict '>'
    assert search.select_best([None, {"v": 1.0}], "v", "min") == 1


def test_single_process_fan_out_runs_everything_in_order():
    res = search.fan_out(CONFIGS, _run_one, "best_val_loss", "min")
    assert res["best_index"] == 2 and res["best_config"] == CONFIGS[2] and res["best_state"]["tag"] == "state-of-2"
    assert [r["best_val_loss"] for r in res["results"]] == LOSSES and all("state" not in r for r in res["results"])


@pytest.mark.timeout(300)
def test_two_rank_fan_out_shards_gathers_and_broadcasts_the_winner():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, best, state, who, losses in got:
        assert best == 2 and losses == LOSSES
        assert who == [i % 2 for i in range(len(LOSSES))]                # round-robin ownership
        assert state == {"tag": "state-of-2", "cfg": CONFIGS[2]}         # the owner's (rank 0) state reached both ranks
