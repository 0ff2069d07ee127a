#!/bin/bash
# publisher queues at N=4 and N=8 (one warp per group), against DFD_NPUB=0 (one system-scope fence per activation) at N=8
cd "$(dirname "$0")/.."
NS="8 4" bash tools/gpu_scale.sh
DFD_NPUB=8 NS=8 bash tools/gpu_scale.sh | sed 's/^/[DFD_NPUB=8] /'
