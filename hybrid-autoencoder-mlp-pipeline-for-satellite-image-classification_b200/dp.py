"""Data-parallel plumbing: one process per GPU, torch.distributed for the rendezvous, one NCCL
communicator owned by libae_b200 for the single flat-gradient allreduce per step."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


class Communicator:
    def __init__(self, handle, rank, world):
        self.handle, self.rank, self.world = handle, rank, world

    def allreduce_(self, t: torch.Tensor):
        check(_lib.load().ae_dp_allreduce(self.handle, C.c_void_p(t.data_ptr()), t.numel(),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return t

    def close(self):
        if self.handle is not None:
            _lib.load().ae_dp_destroy(self.handle)
            self.handle = None


def init_communicator() -> Communicator:
    """Create the library's NCCL communicator; the unique id travels over the already-initialised
    torch.distributed process group (any backend)."""
    lib = _lib.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_uint8 * _lib.DP_UNIQUE_ID_BYTES)()
    if rank == 0:
        check(lib.ae_dp_get_unique_id(buf))
    obj = [bytes(buf)]
    dist.broadcast_object_list(obj, src=0)
    raw = (C.c_uint8 * _lib.DP_UNIQUE_ID_BYTES).from_buffer_copy(obj[0])
    h = C.c_void_p()
    check(lib.ae_dp_init(raw, rank, world, C.byref(h)))
    comm = Communicator(h, rank, world)
    # one eager collective before anything is captured into a CUDA graph: NCCL connects its channels and allocates its
    # buffers on first use, which is not allowed inside a stream capture
    warm = torch.ones(1024, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
    comm.allreduce_(warm)
    torch.cuda.synchronize()
    if abs(float(warm[0]) - world) > 1e-6:
        raise RuntimeError(f"ae_b200.dp: warm-up allreduce returned {float(warm[0])}, expected {world}")
    return comm


def attach_peers(comm: Communicator, model) -> bool:
    """Give the library peer-mapped views of every rank's flat parameter / gradient buffers of `model` (CUDA IPC, all ranks on
    one box): a TrainStep captured afterwards with `comm` runs reduce-scatter + Adam + all-gather as ONE kernel over NVLink
    instead of NCCL allreduces followed by Adam.  Adam's moments are then maintained for the own 1/world shard only.
    Collective: every rank must call it.  Returns False (and changes nothing) when AE_B200_DP_FUSED=0."""
    import os
    if os.environ.get("AE_B200_DP_FUSED", "1") == "0" or comm.world > 8:
        return False
    lib = _lib.load()
    eng = model.engine()
    flat = eng.flat
    if getattr(flat, "_dp_flags", None) is not None and flat._dp_flags[0] is comm:
        return True
    dev = flat.data.device
    flags = torch.zeros(_lib.DP_FLAG_BYTES // 4, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    mine = []
    try:
        for t in (flat.data, flat.grad, flags):
            h = (C.c_uint8 * _lib.DP_IPC_HANDLE_BYTES)()
            off = C.c_int64()
            check(lib.ae_dp_ipc_export(C.c_void_p(t.data_ptr()), h, C.byref(off)))
            mine.append((bytes(h), int(off.value)))
    except RuntimeError as ex:        # e.g. an allocator whose blocks cannot be exported (expandable segments)
        mine = str(ex)
    everyone = [None] * comm.world
    dist.all_gather_object(everyone, mine)
    failed = [(r, m) for r, m in enumerate(everyone) if isinstance(m, str)]
    if failed:                        # every rank sees the same list: all of them keep the NCCL form
        if comm.rank == 0:
            import warnings
            warnings.warn(f"ae_b200.dp: peer memory not available (rank {failed[0][0]}: {failed[0][1]}); the step exchanges "
                          "gradients with NCCL allreduces instead of the fused peer-memory kernel")
        return False
    handles = (C.c_uint8 * (comm.world * 3 * _lib.DP_IPC_HANDLE_BYTES))()
    offsets = (C.c_int64 * (comm.world * 3))()
    for r, bufs in enumerate(everyone):
        for b, (h, off) in enumerate(bufs):
            base = (r * 3 + b) * _lib.DP_IPC_HANDLE_BYTES
            handles[base:base + _lib.DP_IPC_HANDLE_BYTES] = list(h)
            offsets[r * 3 + b] = off
    check(lib.ae_dp_peers_attach(comm.handle, handles, offsets, C.c_void_p(flat.data.data_ptr()), C.c_void_p(flat.grad.data_ptr()),
                                 C.c_void_p(flags.data_ptr()), flat.len))
    flat._dp_flags = (comm, flags)       # keeps the flag block alive as long as the flat buffers
    dist.barrier()                       # every rank has mapped (and zeroed) everything before anyone's first step
    return True


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous batch shard [lo, hi) of `rank`: sizes differ by at most one, all shards non-empty when total >= world."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(model, src: int = 0):
    """Make every rank start from rank `src`'s parameters and buffers (torch.distributed plumbing).  The writes go
    through ``.data`` (no version bump), so the model's packed weights are invalidated explicitly: the next forward or
    TrainStep re-derives them from the broadcast values."""
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src)
    for m in model.modules():
        eng = getattr(m, "_engine", None)
        if eng is not None:
            eng.invalidate()
