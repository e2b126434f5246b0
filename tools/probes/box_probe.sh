#!/usr/bin/env bash
# developer probe: what does the GPU box look like from the host side?
nvidia-smi topo -m 2>&1 | head -30
nvidia-smi --query-gpu=index,name,memory.total,pcie.link.gen.current,pcie.link.width.current --format=csv
lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|^CPU\(s\)|MHz"
free -g | head -3
ls /sys/devices/system/node/ 2>/dev/null | head
nproc
