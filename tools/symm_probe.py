"""torchrun probe: does torch's symmetric memory (peer pointers, NVLS multicast pointer, signal pads) come up on this box?"""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t0 = time.time()
n = 42 * 1024 * 1024
t = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
t.fill_(rank + 1)
torch.cuda.synchronize()
hdl.barrier()
print("rank %d: rendezvous %.2f s, world %d, multicast_ptr %#x, has_multicast %s, buffer_ptrs %s, signal pad %d B" %
      (rank, time.time() - t0, hdl.world_size, hdl.multicast_ptr, _ := None or symm_mem._SymmetricMemory.has_multicast_support(DeviceType := torch._C._autograd.DeviceType.CUDA, local) if hasattr(symm_mem._SymmetricMemory, "has_multicast_support") else "?",
       [hex(p) for p in hdl.buffer_ptrs], hdl.signal_pad_size), flush=True)
peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
print("rank %d reads peer: %s" % (rank, peer[:4].tolist()), flush=True)
dist.barrier()
dist.destroy_process_group()
