#!/bin/bash
# Round-2 first experiment: why is one vmig_migrate_tree call over 8 GPUs (17.6 GiB/s)
# slower than 8 processes with one GPU each (41.8 GiB/s)?  Part A needs no GPU (run it on a 1-GPU box: 1x charge);
# part B needs the 8-GPU box.   usage: bash profiles/scripts/r2_inprocess_diag.sh A|B
set -u
case "${1:-A}" in
A)  echo "numa_balancing=$(cat /proc/sys/kernel/numa_balancing 2>/dev/null) thp=$(cat /sys/kernel/mm/transparent_hugepage/enabled) shmem_thp=$(cat /sys/kernel/mm/transparent_hugepage/shmem_enabled)"
    lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node[0-9]"
    timeout 300 gpu-docker-api_b200/csrc/probe/hostcopy_probe /dev/shm/vmig_hcp 512 split ;;
B)  S=profiles/scripts/e2e_multigpu.py
    timeout 200 python $S 40 1 1,8 2
    VMIG_RING_MBIND=1 timeout 200 python $S 40 1 8 2
    VMIG_READERS=3 VMIG_WRITERS=5 timeout 200 python $S 40 1 8 2
    VMIG_BIND_IO=0 timeout 200 python $S 40 1 8 2 ;;
esac
