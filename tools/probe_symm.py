"""probe: does torch symmetric memory (NVLink peer access, no NCCL on the data path)
work on this box?  torchrun --nproc-per-node 2 tools/probe_symm.py"""
import os
import time
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    import torch.distributed._symmetric_memory as symm
    t = symm.empty((1024,), dtype=torch.float64, device="cuda")
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(float(rank))
    hdl.barrier()
    peer = (rank + 1) % world
    pbuf = hdl.get_buffer(peer, (1024,), torch.float64)
    src = torch.full((16,), 100.0 + rank, dtype=torch.float64, device="cuda")
    pbuf[:16].copy_(src)            # peer store over NVLink
    hdl.barrier()
    torch.cuda.synchronize()
    below = (rank - 1) % world
    ok = bool((t[:16] == 100.0 + below).all()) and bool((t[16:] == rank).all())
    # latency of copy + barrier
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        pbuf[:16].copy_(src)
        hdl.barrier()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0)/200
    print("rank %d: symmetric memory OK=%s, copy+barrier %.1f us, ptr peer=%x" % (
        rank, ok, dt*1e6, hdl.buffer_ptrs[peer]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
