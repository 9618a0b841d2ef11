#!/bin/bash
# A/B on ONE box: alternate `bench.py --only-timed` between this tree and a second tree (default ab_old/, a git worktree of
# the previous commit with its own built .so), N rounds each; prints ms_per_step per run.
#   tools/ab_bench.sh [other_tree] [rounds] [extra bench args...]
OTHER=${1:-ab_old}; ROUNDS=${2:-3}; shift 2 || true
ROOT=$(pwd)
for i in $(seq 1 $ROUNDS); do
  for tree in "$OTHER" "."; do
    cd "$ROOT/$tree"
    python bench.py --only-timed --no-cpu-baseline --no-generate --no-rices --steps 50 --warmup 5 "$@" 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tree', round(d['ms_per_step'],4), d['gpu_launches'])"
  done
done
