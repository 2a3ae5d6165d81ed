#!/bin/bash
cd "$(dirname "$0")/.."
B=tools/bench_conv
export VTB_GRAPH=1 VTB_WG_NOREDUCE=1
sweep() {
  echo "## $*"
  for kp in 64 128; do for bx in 1 2 3 4; do
    echo -n "kpix=$kp boxes=$bx: "; VTB_WG_KPIX=$kp VTB_WG_BOXES=$bx $B "$@" | grep -E "wgrad  graph|ctas" | tail -2 | tr '\n' ' ' | sed 's/(host-free, back to back)//; s/setup.*producer0/producer0/'; echo
  done; done
}
sweep 256 22 22 128 128 3 1 1
sweep 256 11 11 256 256 3 1 1
sweep 256 6 6 512 512 3 1 1
