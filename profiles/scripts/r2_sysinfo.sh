#!/bin/bash
# Host facts the end-to-end path depends on (round 2): NUMA balancing, THP, what backs /tmp and /dev/shm,
# cuFile availability, PCIe/NUMA topology.
echo "numa_balancing=$(cat /proc/sys/kernel/numa_balancing 2>/dev/null) thp=$(cat /sys/kernel/mm/transparent_hugepage/enabled) shmem_thp=$(cat /sys/kernel/mm/transparent_hugepage/shmem_enabled 2>/dev/null)"
uname -r
lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node[0-9]|L3"
for n in /sys/devices/system/node/node*; do echo "$n: $(grep -E 'MemTotal|MemFree' $n/meminfo | tr -s ' ' | tr '\n' ' ')"; done
echo "--- memory limits: cgroup memory.max=$(cat /sys/fs/cgroup/memory.max 2>/dev/null) current=$(cat /sys/fs/cgroup/memory.current 2>/dev/null) v1=$(cat /sys/fs/cgroup/memory/memory.limit_in_bytes 2>/dev/null)"; free -g | head -2; nproc
echo "--- mounts"; df -hT /tmp /dev/shm /root 2>/dev/null; mount | grep -E " /tmp | /dev/shm | / " | head
echo "--- block devices"; lsblk -o NAME,SIZE,TYPE,ROTA,MOUNTPOINT 2>/dev/null | head -20
echo "--- cufile"; ls /usr/local/cuda/lib64 2>/dev/null | grep -i cufile; ls /usr/local/cuda/gds 2>/dev/null | head; lsmod 2>/dev/null | grep -E "nvidia_fs|nvidia" | head
echo "--- ulimit -l: $(ulimit -l)   nofile: $(ulimit -n)"
nvidia-smi topo -m 2>/dev/null | head -20
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv 2>/dev/null
echo "--- dirty/writeback sysctls"; sysctl vm.dirty_ratio vm.dirty_background_ratio 2>/dev/null
