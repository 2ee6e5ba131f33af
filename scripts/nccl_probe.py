"""Raw NCCL timings on the box for payloads of the merge's size (torchrun, one process per GPU): what the library gives for
reduce / all_reduce / reduce_scatter / gather-by-send-recv at the size of a C1 union (6208 bricks x 96 KB), as a yardstick for gsb_tsdf_reduce's phases."""
import os, sys, json, torch, torch.distributed as dist

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    bricks = 6208  # a C1 union, multiple of 8
    n = bricks * 4096 * 6  # 16^3 voxels x (tsdf*w, w, rgb*w + pad) floats
    x = torch.ones(n, device="cuda")
    out = {}
    def timed(name, fn, it=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(it): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / it], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = round(t.item(), 4)
    timed("reduce", lambda: dist.reduce(x, 0))
    timed("all_reduce", lambda: dist.all_reduce(x))
    sl = n // world
    timed("reduce_scatter_inplace", lambda: dist.reduce_scatter_tensor(x[rank * sl:(rank + 1) * sl], x))
    def gather():
        ops = []
        if rank == 0:
            for r in range(1, world): ops.append(dist.P2POp(dist.irecv, x[r * sl:(r + 1) * sl], r))
        else:
            ops.append(dist.P2POp(dist.isend, x[rank * sl:(rank + 1) * sl], 0))
        for w in dist.batch_isend_irecv(ops): w.wait()
    timed("gather_send_recv", gather)
    if rank == 0:
        print(json.dumps({"world": world, "payload_MB": round(n * 4 / 1e6, 1), "ms": out}))
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
