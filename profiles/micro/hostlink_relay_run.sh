#!/bin/bash
# Can the GPUs with the slow host path (0-3 on this pool's boxes) export through the copy engines of the fast ones (4-7) over NVLink?
# usage: bash profiles/micro/hostlink_relay_run.sh > gpurun_out/r02_hostlink_relay.txt 2>&1
P=profiles/micro/hostlink_probe
A="--dir d2h --mb 256 --secs 1.0"
t() { echo "# t=$(date +%s.%N | cut -c1-14)"; }
t; echo "## direct, all 8 / lower 4 / upper 4 (1 s windows: steady-state aggregate)"
$P $A --set 0,1,2,3,4,5,6,7 --set 0,1,2,3 --set 4,5,6,7
t; echo "## relay 1 (stream of GPU d+4 copies GPU d's memory to the host), all 8: 0-3 through 4-7, 4-7 direct"
$P $A --relay 1 --set 0,1,2,3,4,5,6,7
t; echo "## relay 2 (peer copy into a staging buffer on GPU d+4, then D2H from there), all 8"
$P $A --relay 2 --set 0,1,2,3,4,5,6,7
t; echo "## only GPUs 0-3 have data, exported through 4-7 (relay-base 4)"
$P $A --relay 1 --relay-base 4 --set 0,1,2,3
$P $A --relay 2 --relay-base 4 --set 0,1,2,3
t; echo "## GPUs 0,1 through 4,5"
$P $A --relay 2 --relay-base 4 --set 0,1
t; echo "## relay through 6,7 only (two links carry all eight)"
$P $A --relay 2 --relay-base 6 --set 0,1,2,3,4,5,6,7
t
