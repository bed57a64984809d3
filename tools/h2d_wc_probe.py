"""Probe: aggregate host->device bandwidth of all ranks copying at once, from ordinary pinned memory
(cudaHostAllocDefault, what torch's pin_memory() gives) and from write-combined pinned memory
(cudaHostAllocWriteCombined).  torchrun --nproc-per-node N tools/h2d_wc_probe.py
"""
import ctypes
import json
import os
import pathlib
import sys

import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def host_alloc(nbytes, flags):
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), buf


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ncpu = bench.bind_to_gpu_numa_node(local)
    n = 204_603_392
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    res = {"rank": local, "cpus": ncpu}
    keep = []
    for name, flags in (("default", 0), ("write_combined", 4), ("torch_pin", None)):
        if flags is None:
            src = torch.empty(n, dtype=torch.uint8).pin_memory()
        else:
            src, buf = host_alloc(n, flags)
            keep.append(buf)
        src.fill_(3)
        res[name + "_is_pinned"] = bool(src.is_pinned())
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[name + "_agg_gbs"] = world * n / (float(ms.item()) * 1e-3) / 1e9
        assert int(dst[12345].item()) == 3
    if local == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
