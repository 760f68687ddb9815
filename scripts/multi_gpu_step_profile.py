"""Where does a short multi-GPU timed region lose time?  Per-step CUDA events of the bench's region (20 forwards behind a
device-side rendezvous), rank 0 prints them.   torchrun --nproc-per-node N scripts/multi_gpu_step_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import koemorph_b200 as K
from oracle import koemorph_oracle as O

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
w = O.make_weights(1234, 30, style="init")
m = K.SequentialDualStreamModel().to(dev).eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
m.precision = "bf16"
audio = 0.1 * torch.randn(512, 136000, device=dev)
eg = torch.randn(512, 264, device=dev)
kept = torch.empty(20, 512, 1, 52, device=dev)
gathered = torch.empty(world * 20, 512, 1, 52, device=dev) if world > 1 else None
tiny = torch.zeros(1, device=dev)
for i in range(5):
    m(audio, egemaps=eg, out=kept[i])
if world > 1:
    dist.all_gather_into_tensor(gathered, kept); dist.all_reduce(tiny); dist.barrier()
torch.cuda.synchronize()
for rep in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(23)]
    ev[0].record()
    if world > 1:
        dist.all_reduce(tiny)
    ev[1].record()
    for i in range(20):
        m(audio, egemaps=eg, out=kept[i])
        ev[2 + i].record()
    if world > 1:
        dist.all_gather_into_tensor(gathered, kept)
    ev[22].record()
    torch.cuda.synchronize()
    if rank == 0:
        steps = [ev[1 + i].elapsed_time(ev[2 + i]) * 1e3 for i in range(20)]
        print(f"world {world} rep {rep}: gate {ev[0].elapsed_time(ev[1])*1e3:.0f} us; steps " + " ".join(f"{s:.0f}" for s in steps) +
              f"; gather {ev[21].elapsed_time(ev[22])*1e3:.0f} us; total {ev[1].elapsed_time(ev[22]):.3f} ms")
if world > 1:
    dist.destroy_process_group()
