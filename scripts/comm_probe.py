"""Microseconds per slab-communication operation on N GPUs (torchrun): reduction, one-way send + wait, handshake exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 scripts/comm_probe.py
"""
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfd = importlib.import_module("macos-cfd_b200")
    ids = [cfd.comm_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    out = {"n_gpus": world}
    for p2p in (1,):
        s = cfd.Solver(8192, 1024 * world, 1.0, 1.0, device=local, rank=rank, nranks=world, comm_id=ids[0])
        stream = torch.cuda.ExternalStream(s.stream, device=local)
        names = {0: "reduce", 1: "send+wait (2 rows of u, v)", 2: "handshake exchange (2 rows of u, v)", 3: "reduce + 1 row of p"}
        for kind, name in names.items():
            s.debug_comm(kind, 20)
            s.sync()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 200
            e0.record(stream)
            s.debug_comm(kind, reps)
            e1.record(stream)
            s.sync()
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[f"{s.comm_mode}: {name}"] = round(float(t.item()), 2)
        s.close()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
