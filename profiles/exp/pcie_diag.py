"""Multi-rank PCIe / NUMA diagnostic: per-rank pinned H2D + D2H bandwidth alone and with all ranks
moving data at once; prints topology facts.  torchrun --nproc-per-node N profiles/exp/pcie_diag.py"""
import os, sys, time, subprocess
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch, torch.distributed as dist
import bench
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
bind = os.environ.get("DIAG_BIND", "1") == "1"
aff0 = sorted(os.sched_getaffinity(0))
msg = bench.bind_to_gpu_numa_node(lr) if bind else "unbound (by request)"
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; ls /sys/devices/system/node | head; nproc; grep -c processor /proc/cpuinfo; free -g | head -2"]):
        print(subprocess.run(cmd, capture_output=True, text=True).stdout, flush=True)
    print("initial affinity", len(aff0), aff0[:4], "...", aff0[-4:], flush=True)
N = 256 << 20
h_in = torch.empty(N, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(N, dtype=torch.uint8).pin_memory(); h_out.fill_(2)
d_a = torch.empty(N, dtype=torch.uint8, device="cuda"); d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both=True, reps=6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return N * reps / (time.perf_counter() - t0) / 1e9
run()
res = {}
for r in range(world):
    dist.barrier()
    if r == rank:
        res["alone_h2d"] = run(False); res["alone_bidir_each"] = run(True)
    dist.barrier()
dist.barrier()
res["all_h2d"] = run(False)
dist.barrier()
res["all_bidir_each"] = run(True)
dist.barrier()
node = "?"
try:
    bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(lr)], capture_output=True, text=True).stdout.strip().lower()
    if bus.startswith("00000000:"): bus = bus[4:]
    node = Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip()
except Exception as e: node = repr(e)
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank} gpu_numa_node {node} bind '{msg}' now_on_cpus {len(os.sched_getaffinity(0))}: " + ", ".join(f"{k} {v:.1f} GB/s" for k, v in res.items()), flush=True)
dist.barrier(); dist.destroy_process_group()
