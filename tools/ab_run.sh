#!/bin/bash
# Development aid: same-box A/B of library variants built by tools/ab_build.sh.
#   tools/ab_run.sh "<workloads>" name1 name2 ...   (two interleaved rounds per variant)
cd "$(dirname "$0")/.."
W=$1; shift
for round in 1 2; do
for v in "$@"; do
  echo "== $v (round $round)"; python tools/kbench.py --steps 40 --check --lib tools/_ab/$v.so $W 2>&1 | tail -n 8
done; done
