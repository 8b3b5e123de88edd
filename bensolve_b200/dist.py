"""Multi-GPU plumbing: one process per GPU, `torch.distributed` for the rendezvous, NCCL (called
from the C++ engine on its own stream) for the per-cut exchange.

    import torch.distributed as dist
    dist.init_process_group("nccl")                 # torchrun sets RANK / WORLD_SIZE / MASTER_*
    lib = capi.load_product()
    init_comm(lib)                                  # before the first poly__initialise
    ...                                             # every rank issues the same poly__* calls
    finalize_comm(lib)

For the CPU test double (tests/, gloo backend) the all-gather is a host callback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_CALLBACKS = []   # keep ctypes callbacks alive


def init_comm(lib, emulate: bool = False) -> int:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 1
    rank, world = dist.get_rank(), dist.get_world_size()
    lib.b200_comm_unique_id.argtypes = [C.c_char_p]
    lib.b200_comm_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
    buf = C.create_string_buffer(128)
    if rank == 0:
        if lib.b200_comm_unique_id(buf):
            raise RuntimeError("b200_comm_unique_id failed")
    payload = [buf.raw]
    dist.broadcast_object_list(payload, src=0)
    if emulate:
        fn_t = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t)

        def allgather(send, recv, nbytes):
            src = torch.from_numpy(np.frombuffer(C.string_at(send, nbytes), dtype=np.uint8).copy())
            outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(outs, src)
            cat = torch.cat(outs).contiguous().numpy()       # keep the buffer alive across the copy
            C.memmove(recv, cat.ctypes.data, nbytes * world)

        cb = fn_t(allgather)
        _CALLBACKS.append(cb)
        lib.b200_comm_set_allgather_callback.argtypes = [fn_t]
        if lib.b200_comm_set_allgather_callback(cb):
            raise RuntimeError("this library does not take an all-gather callback")
    if lib.b200_comm_init(rank, world, payload[0]):
        raise RuntimeError("b200_comm_init failed: " + lib.b200_last_error().decode())
    return world


def finalize_comm(lib) -> None:
    lib.b200_comm_finalize()
