"""Feasibility probe: CUDA IPC memory handles + peer access between the ranks of a torchrun job on this pool."""
import os, sys
import torch, torch.distributed as dist
from cuda.bindings import runtime as rt

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
err, = rt.cudaSetDevice(local)
err, ptr = rt.cudaMalloc(1 << 20)
assert err == rt.cudaError_t.cudaSuccess, err
err, = rt.cudaMemset(ptr, 0x10 + rank, 1 << 20)
err, h = rt.cudaIpcGetMemHandle(ptr)
print(rank, "get handle", err, flush=True)
handles = [None] * world
dist.all_gather_object(handles, bytes(h.reserved))
peer = (rank + 1) % world
ph = rt.cudaIpcMemHandle_t()
ph.reserved = handles[peer]
err, pptr = rt.cudaIpcOpenMemHandle(ph, rt.cudaIpcMemLazyEnablePeerAccess)
print(rank, "open peer handle", err, flush=True)
if err == rt.cudaError_t.cudaSuccess:
    import ctypes
    buf = (ctypes.c_ubyte * 16)()
    err, = rt.cudaMemcpy(buf, pptr, 16, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    print(rank, "read peer memory", err, list(buf)[:4], "expected", 0x10 + peer, flush=True)
    err, can = rt.cudaDeviceCanAccessPeer(local, peer)
    print(rank, "canAccessPeer", can, flush=True)
dist.barrier()
