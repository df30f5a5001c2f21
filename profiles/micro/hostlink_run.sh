#!/bin/bash
# Host-link survey of an 8-GPU box (round 2): topology + simultaneous pinned-copy bandwidth for GPU subsets.
# usage (from the repo root on the GPU box): bash profiles/micro/hostlink_run.sh > gpurun_out/r02_hostlink_probe.txt 2>&1
P=profiles/micro/hostlink_probe
echo "## nvidia-smi topo -m"; nvidia-smi topo -m 2>&1 | head -40
echo "## lscpu"; lscpu | egrep 'Model name|^CPU\(s\)|Socket|NUMA|Thread|Hypervisor'
echo "## numa"; (numactl -H 2>/dev/null || cat /sys/devices/system/node/node*/meminfo 2>/dev/null | egrep 'MemTotal') | head -20
echo "## per-GPU pci"; for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do s=$(echo ${b#0000} | tr A-Z a-z); d=/sys/bus/pci/devices/$s; echo "$b numa=$(cat $d/numa_node 2>/dev/null) cpus=$(cat $d/local_cpulist 2>/dev/null) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; done
nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current --format=csv
N=$(nvidia-smi -L | wc -l)
ALL=$(seq -s, 0 $((N-1)))
echo "## d2h, one process per GPU"
if [ "$N" -ge 8 ]; then
$P --dir d2h --set 0 --set 1 --set 4 --set 7 --set 0,1 --set 0,2 --set 0,4 --set 0,7 --set 2,3 --set 0,1,2,3 --set 4,5,6,7 --set 0,2,4,6 --set $ALL
else
$P --dir d2h --set 0 --set $ALL
fi
echo "## h2d / both directions"
$P --dir h2d --set 0 --set $ALL
$P --dir both --set 0 --set $ALL
echo "## one process, one thread per GPU"
$P --mode thread --dir d2h --set $ALL
echo "## bound to the GPU's local CPUs / write-combined host buffers"
$P --dir d2h --bind --set 0 --set $ALL
$P --dir d2h --wc --set $ALL
echo "## 64 MiB copies"
$P --dir d2h --mb 64 --iters 192 --set 0 --set $ALL
echo "## with one 'nvidia-smi -lms 50' poller per GPU running (what bench.py's clock sampler does)"
pids=""
for i in $(seq 0 $((N-1))); do nvidia-smi -i $i --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 50 > /dev/null 2>&1 & pids="$pids $!"; done
sleep 1
$P --dir d2h --set 0 --set $ALL
kill $pids
echo "## again, plain"
$P --dir d2h --set $ALL
