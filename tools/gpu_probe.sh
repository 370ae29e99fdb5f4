#!/bin/bash
# host / GPU topology of the gpurun box (read once per round; decides thread and NUMA policy of the host paths)
{
echo "== nproc: $(nproc)  affinity: $(taskset -p $$ 2>/dev/null)"
lscpu 2>/dev/null | egrep "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core" 
echo "== numa nodes online: $(cat /sys/devices/system/node/online 2>/dev/null)"
for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist 2>/dev/null) $(grep MemTotal $n/meminfo 2>/dev/null)"; done
echo "== cpuset: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null) mems: $(cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null)"
echo "== cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"
grep -E "MemTotal|MemAvailable" /proc/meminfo
nvidia-smi --query-gpu=index,name,pci.bus_id,clocks.max.sm,power.limit --format=csv
nvidia-smi topo -m 2>/dev/null
which numactl ncu compute-sanitizer 2>/dev/null
} 2>&1
