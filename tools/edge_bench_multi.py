"""dev tool (multi-GPU): does the edge pass slow down when the GPUs of one box run it at the same time?  (SCALE_r01: the
edge-pass roofline fraction read 0.70 at N = 1, 2 and 0.53 at N = 4, 8.)  Under torchrun, every rank times the edge pass
alone on its own GPU (a) all ranks at once, after a barrier, (b) one rank at a time, the others idle.
usage: python -m torch.distributed.run --nproc-per-node N tools/edge_bench_multi.py [B]"""
import ctypes, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mergenet_b200 import _lib, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, W, C, K = 1024, 2048, 9, 10
offs = np.ascontiguousarray(np.array(synth.generate_offsets(40, K), np.int32))
L = _lib.lib()
gb = 4.0 * H * W * ((C + K) + (C + 2 * K)) * B / 1e9


def run():
    ms = ctypes.c_float(0)
    rc = L.mn_debug_edge_bench(H, W, C, K, offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), B, 20, 0, ctypes.byref(ms))
    return rc, max(ms.value, 1e-9)  # (the hook runs on the caller's current device: torch.cuda.set_device above)


def barrier():
    dist.barrier()
    torch.cuda.synchronize()


run()  # warm-up (allocations, module load)
barrier()
rc, ms = run()
print("rank %d ALL-AT-ONCE rc=%d ms/launch=%.3f GB/s=%.1f frac=%.3f" % (rank, rc, ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / 6538.9), flush=True)
for r in range(world):
    barrier()
    if r == rank:
        rc, ms = run()
        print("rank %d ALONE       rc=%d ms/launch=%.3f GB/s=%.1f frac=%.3f" % (rank, rc, ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / 6538.9), flush=True)
barrier()
dist.destroy_process_group()
