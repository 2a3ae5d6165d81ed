#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest "$@" -m gpu -x -q -s 2>&1 | tail -25
