#!/bin/bash
cd "$(dirname "$0")/.."
B=tools/bench_conv
export VTB_GRAPH=1
r() { echo "## $*  [NA=$VTB_CONV_NA]"; $B "$@" 2>&1 | grep -E "graph replay|waits" | head -4; }
for na in 2 4; do
export VTB_CONV_NA=$na
r 256 176 176 16 32 3 1 1
r 256 176 176 32 64 3 2 1
r 256 88 88 32 32 3 1 1
r 256 44 44 64 64 3 1 1
r 256 88 88 64 128 3 2 1
r 256 22 22 128 128 3 1 1
done
