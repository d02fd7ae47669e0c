"""Grid-search fan-out: the reference's hyper-parameter searches are embarrassingly parallel.

NB:2629-2741 trains one supervised autoencoder per (alpha, learning rate) pair -- 45 configurations -- one after the
other on one device and keeps the one with the lowest validation loss; NB:3447-3540 does the same over 11 MLP learning
rates by validation accuracy.  The runs are independent, so with one process per GPU the configurations are dealt out
round-robin and only the per-configuration results (a few floats) are exchanged: no gradient traffic at all, which for
this workload beats data-parallel training of one configuration at a time.

`torch.distributed` (any backend) is plumbing: the results travel with ``all_gather_object``, the winning state dict
with ``broadcast_object_list``.  Without an initialised process group the same functions run every configuration
locally, in the reference's order.
"""
from __future__ import annotations

import itertools
from typing import Callable, Dict, List, Optional, Sequence

import torch.distributed as dist


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def assigned(n_configs: int, rank: int, world: int) -> List[int]:
    """Indices of the configurations rank `rank` runs: round-robin, so long and short runs (early stopping) mix."""
    return list(range(rank, n_configs, world))


def select_best(results: Sequence[Optional[Dict]], key: str, mode: str = "min") -> int:
    """The reference's selection rule: walk the configurations in grid order and replace the incumbent only on a STRICT
    improvement (``if best_val_loss < global_best_loss`` NB:2732, ``if best_val_acc > global_best_val`` NB:3536), so the
    earliest configuration wins ties.  The reference starts from +inf (losses) / 0 (accuracies): an accuracy of 0 never
    becomes the incumbent.  Returns the index, or -1 if none qualifies."""
    best_i, best_v = -1, (float("inf") if mode == "min" else 0.0)
    for i, r in enumerate(results):
        if r is None:
            continue
        v = float(r[key])
        if (v < best_v) if mode == "min" else (v > best_v):
            best_i, best_v = i, v
    return best_i


def fan_out(configs: Sequence, run_one: Callable[[int, object], Dict], key: str, mode: str = "min",
            state_key: Optional[str] = "state") -> Dict:
    """Run ``run_one(index, config)`` for this rank's share of `configs`, gather every rank's results and pick the
    winner by the reference's rule.  ``run_one`` returns a dict of picklable scalars / lists plus, optionally, the
    trained state dict under `state_key`; states are not gathered -- only the winner's is broadcast from its owner.
    Returns {"results": [...in grid order, without states...], "best_index", "best_config", "best", "best_state"}."""
    rank, world = _rank_world()
    mine = {}
    for i in assigned(len(configs), rank, world):
        mine[i] = run_one(i, configs[i])
    light = {i: {k: v for k, v in r.items() if k != state_key} for i, r in mine.items()}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, light)
    else:
        gathered = [light]
    results: List[Optional[Dict]] = [None] * len(configs)
    for part in gathered:
        for i, r in part.items():
            results[i] = r
    best = select_best(results, key, mode)
    state = None
    if best >= 0 and state_key is not None:
        owner = best % world
        if world > 1:
            box = [mine[best].get(state_key) if rank == owner else None]
            dist.broadcast_object_list(box, src=owner)
            state = box[0]
        else:
            state = mine[best].get(state_key)
    return {"results": results, "best_index": best, "best_config": configs[best] if best >= 0 else None,
            "best": results[best] if best >= 0 else None, "best_state": state}


def ae_grid(alpha_values: Sequence[float], lr_values: Sequence[float]):
    """(alpha, lr) pairs in the reference's nesting order (NB:2642-2643: alpha outer, learning rate inner)."""
    return list(itertools.product(alpha_values, lr_values))


def grid_search_autoencoder(alpha_values, lr_values, train_loader, val_loader, latent_dim: int = 64, num_classes: int = 10,
                            num_epochs: int = 80, patience: int = 15, precision=None, log=None) -> Dict:
    """NB:2629-2741 fanned out over the ranks; every rank holds its own copy of the (device-resident) loaders."""
    from . import fit
    from .modules import SupervisedAutoencoder
    from .optim import Adam

    dev = train_loader.dataset.images.device

    def run_one(i, cfg):
        alpha, lr = cfg
        model = SupervisedAutoencoder(latent_dim, num_classes, precision=precision).to(dev)        # NB:2650
        model.engine().prepare(dev, train_loader.batch_size)
        opt = Adam(model.parameters(), lr=lr)                                                     # NB:2654
        res = fit.fit_autoencoder(model, opt, train_loader, val_loader, alpha, num_epochs, patience, log=log)
        res["state"] = {k: v.detach().cpu() for k, v in model.state_dict().items()}              # NB:2735 (last state)
        return res

    return fan_out(ae_grid(alpha_values, lr_values), run_one, "best_val_loss", "min")


def grid_search_mlp(lr_values, train, val, input_dim: int = 64, num_classes: int = 10, num_epochs: int = 30,
                    batch_size: int = 64, log=None) -> Dict:
    """NB:3447-3540 fanned out over the ranks (latents already on each rank's device)."""
    from . import fit
    from .modules import MLP
    from .optim import Adam

    dev = train[0].device

    def run_one(i, lr):
        clf = MLP(input_dim, num_classes).to(dev)                                                 # NB:3460
        clf._state.prepare(dev, batch_size)
        opt = Adam(clf.parameters(), lr=lr, weight_decay=1e-4)                                    # NB:3461
        res = fit.fit_mlp(clf, opt, train, val, num_epochs, batch_size, log=log)
        res["state"] = {k: v.cpu() for k, v in (res.pop("best_state") or {}).items()}            # NB:3519
        return res

    return fan_out(list(lr_values), run_one, "best_val_acc", "max")
