#!/bin/bash
# A/B runs of bench.py over argument sets: tools/ab_args.sh "<label>|<bench args>" ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  label=${spec%%|*}; args=${spec#*|}
  out=$(python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e $args 2>&1 | python tools/brief.py)
  echo "$label :: $out"
done
