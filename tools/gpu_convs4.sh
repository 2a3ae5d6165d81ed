#!/bin/bash
cd "$(dirname "$0")/.."
B=tools/bench_conv
r() { echo "## $*"; VTB_FLUSH=1 $B "$@" 2>&1 | grep -vE "^\s*$"; }
r 256 22 22 128 128 1 1 0
r 256 11 11 256 256 1 1 0
r 256 6 6 512 512 1 1 0
r 256 44 44 64 64 1 1 0
