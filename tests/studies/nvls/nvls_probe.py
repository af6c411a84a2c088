"""Feasibility probe for DESIGN.md 8(1) (multi-GPU box, NOT run in round 1 -- no GPU minutes were left; written
against the CUDA 12.9 driver API through cuda-python): can the layout replicas of N processes be bound to one NVSwitch
multicast object, so that the owner's row stores reach every replica with ONE store (egress 1x instead of (N-1)x)?

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/studies/nvls/nvls_probe.py [rows]

Steps (each prints what it did; the first failing driver call is reported with its error name):
  1. every rank: device = LOCAL_RANK; CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
  2. rank 0: cuMulticastCreate(numDevices = world, POSIX fd handle) -> export fd -> send it to the other ranks over a
     Unix socket (SCM_RIGHTS); they cuMemImportFromShareableHandle it
  3. every rank: cuMulticastAddDevice; barrier; cuMemCreate (local physical memory, multicast granularity);
     cuMulticastBindMem; barrier
  4. every rank: map the local memory (plain mapping) and the multicast object (multicast mapping)
  5. every rank writes ITS slice of rows through the multicast mapping (k_multimem_store_rows: compiles to a plain
     STG -- the replication is a property of the mapping), barrier, then checks that its LOCAL replica holds every
     rank's slice; timing of the store kernel against the same rows written with plain stores into one replica.
The kernel cubin is built in place with nvcc (sm_100a).
"""
import ctypes
import os
import socket
import subprocess
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
from cuda.bindings import driver as drv

HERE = os.path.dirname(os.path.abspath(__file__))


def ck(res, what):
    err = res[0]
    if err != drv.CUresult.CUDA_SUCCESS:
        _, name = drv.cuGetErrorName(err)
        raise RuntimeError(f"{what}: {name.decode() if name else err}")
    return res[1] if len(res) == 2 else res[1:]


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 11_000_000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    torch.zeros(1, device="cuda")                                  # primary context current on this thread
    dev = ck(drv.cuDeviceGet(local), "cuDeviceGet")
    sup = ck(drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev), "attr")
    print(f"[{rank}] multicast supported: {sup}", flush=True)
    if not sup:
        return
    POSIX = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    mprop = drv.CUmulticastObjectProp()
    mprop.numDevices = world
    mprop.handleTypes = POSIX
    mprop.flags = 0
    mprop.size = 1 << 21
    gran = ck(drv.cuMulticastGetGranularity(mprop, drv.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED), "granularity")
    size = ((rows * 8 + gran - 1) // gran) * gran
    mprop.size = size
    print(f"[{rank}] granularity {gran}, size {size}", flush=True)

    # ---- 2. one multicast object, shared through a POSIX fd
    sock_path = f"/tmp/annembed_nvls_{os.environ.get('MASTER_PORT', '0')}.sock"
    if rank == 0:
        mc = ck(drv.cuMulticastCreate(mprop), "cuMulticastCreate")
        fd = int(ck(drv.cuMemExportToShareableHandle(mc, POSIX, 0), "export multicast handle"))
        if os.path.exists(sock_path):
            os.unlink(sock_path)
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(sock_path)
        srv.listen(world)
        dist.barrier()
        for _ in range(world - 1):
            conn, _ = srv.accept()
            socket.send_fds(conn, [b"mc"], [fd])
            conn.close()
        srv.close()
        os.unlink(sock_path)
    else:
        dist.barrier()
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(sock_path)
        _, fds, _, _ = socket.recv_fds(c, 16, 1)
        c.close()
        mc = ck(drv.cuMemImportFromShareableHandle(fds[0], POSIX), "import multicast handle")
    print(f"[{rank}] multicast object ready", flush=True)

    # ---- 3. add the device, bind local physical memory
    ck(drv.cuMulticastAddDevice(mc, dev), "cuMulticastAddDevice")
    dist.barrier()                                                 # every device added before any bind
    aprop = drv.CUmemAllocationProp()
    aprop.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    aprop.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    aprop.location.id = local
    aprop.requestedHandleTypes = POSIX
    mem = ck(drv.cuMemCreate(size, aprop, 0), "cuMemCreate")
    ck(drv.cuMulticastBindMem(mc, 0, mem, 0, size, 0), "cuMulticastBindMem")
    dist.barrier()

    # ---- 4. mappings
    acc = drv.CUmemAccessDesc()
    acc.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = local
    acc.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    uc = ck(drv.cuMemAddressReserve(size, gran, 0, 0), "reserve (local)")
    ck(drv.cuMemMap(uc, size, 0, mem, 0), "map (local)")
    ck(drv.cuMemSetAccess(uc, size, [acc], 1), "access (local)")
    mcp = ck(drv.cuMemAddressReserve(size, gran, 0, 0), "reserve (multicast)")
    ck(drv.cuMemMap(mcp, size, 0, mc, 0), "map (multicast)")
    ck(drv.cuMemSetAccess(mcp, size, [acc], 1), "access (multicast)")
    ck(drv.cuMemsetD8(uc, 0, size), "memset")
    torch.cuda.synchronize()
    dist.barrier()
    print(f"[{rank}] mapped: local {int(uc):#x}, multicast {int(mcp):#x}", flush=True)

    # ---- 5. kernel
    cubin = os.path.join(HERE, "nvls_probe_kernel.cubin")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-o", cubin,
                           os.path.join(HERE, "nvls_probe_kernel.cu")])
    mod = ck(drv.cuModuleLoadData(open(cubin, "rb").read()), "module")
    k_mc = ck(drv.cuModuleGetFunction(mod, b"k_multimem_store_rows"), "function")
    k_pl = ck(drv.cuModuleGetFunction(mod, b"k_plain_store_rows"), "function")
    per = (rows + world - 1) // world
    first, count = rank * per, max(0, min(per, rows - rank * per))
    stream = torch.cuda.current_stream().cuda_stream

    def launch(fn, base, tag):
        args = ((int(base), first, count, float(tag)), (ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_float))
        ck(drv.cuLaunchKernel(fn, (count + 255) // 256, 1, 1, 256, 1, 1, 0, stream, args, 0), "launch")

    for name, fn, base in (("multicast", k_mc, mcp), ("plain (one replica)", k_pl, uc)):
        for _ in range(3):
            launch(fn, base, rank + 1)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(20):
            launch(fn, base, rank + 1)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        print(f"[{rank}] {name}: {count * 8 / dt / 1e9:.1f} GB/s of rows stored ({dt * 1e3:.3f} ms for {count * 8 / 1e6:.1f} MB)", flush=True)
        dist.barrier()
    # every replica must now hold every rank's slice (written through the multicast mapping)
    for r in range(world):
        launch(k_mc, mcp, 100 + rank) if r == rank else None
    torch.cuda.synchronize(); dist.barrier()
    host = np.empty(rows * 2, np.float32)
    ck(drv.cuMemcpyDtoH(host.ctypes.data, uc, rows * 8), "copy back")
    got = host.reshape(rows, 2)
    ok = True
    for r in range(world):
        f, c = r * per, max(0, min(per, rows - r * per))
        ok &= bool(np.all(got[f:f + c, 0] == 100 + r)) and bool(np.all(got[f:f + c, 1] == np.arange(f, f + c, dtype=np.float32)))
    print(f"[{rank}] local replica holds every rank's slice: {ok}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
