"""Prints what the host-side placement of the per-GPU ranks can be based on: allowed cores, NUMA nodes, NVML's ideal
cores per GPU and nvidia-smi's topology matrix.  usage: python tools/topo_probe.py"""
import os, subprocess
print("allowed cpus:", sorted(os.sched_getaffinity(0)))
print(subprocess.run("lscpu | grep -i -E 'numa|model name|^CPU\\(s\\)|socket'", shell=True, capture_output=True, text=True).stdout)
try:
    import pynvml as N
    N.nvmlInit()
    for i in range(N.nvmlDeviceGetCount()):
        h = N.nvmlDeviceGetHandleByIndex(i)
        words = N.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [w * 64 + b for w, x in enumerate(words) for b in range(64) if (x >> b) & 1]
        print("gpu", i, "ideal cpus:", cpus[:4], "...", cpus[-4:], len(cpus))
except Exception as e:
    print("nvml:", e)
print(subprocess.run("nvidia-smi topo -m", shell=True, capture_output=True, text=True).stdout)
