"""Does this box's NVSwitch fabric expose NVLS multicast to torch symmetric memory?  (2+ GPUs)"""
import os
import sys
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def w(rank, world):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    import torch.distributed._symmetric_memory as symm
    t = symm.empty((1024,), dtype=torch.float32, device=f"cuda:{rank}")
    h = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "multicast_ptr", getattr(h, "multicast_ptr", None), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    mp.spawn(w, args=(n,), nprocs=n)
