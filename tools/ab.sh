#!/bin/bash
# A/B runs of bench.py over library variants / env settings: tools/ab.sh "<label>|<env assignments>" ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  label=${spec%%|*}; envs=${spec#*|}
  out=$(env $envs python bench.py --scenarios 2048 --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e 2>&1 | python tools/brief.py)
  echo "$label :: $out"
done
