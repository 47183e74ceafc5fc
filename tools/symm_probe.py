import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda",local))
g = dist.new_group(list(range(world)))
try:
    t = symm_mem.empty(1<<20, dtype=torch.uint8, device=torch.device("cuda",local))
    h = symm_mem.rendezvous(t, g)
    print(rank, "rendezvous ok", h.world_size, [hex(p) for p in h.buffer_ptrs], "multicast", h.has_multicast_support(torch.device("cuda").type if False else 0, local) if False else hex(h.multicast_ptr), "sigpad", h.signal_pad_size, flush=True)
    t.fill_(rank+1)
    h.barrier()
    peer = h.get_buffer((rank+1)%world, (16,), torch.uint8)
    print(rank, "peer read", peer[:4].tolist(), flush=True)
    h.barrier()
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "FAILED", repr(e), flush=True)
dist.destroy_process_group()
