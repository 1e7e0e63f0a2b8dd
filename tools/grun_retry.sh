#!/bin/bash
# tools/grun_retry.sh <timeout> '<command>' : tools/grun.sh, retried while the pod answers "transient" (no box free)
cd "$(dirname "$0")/.."
for i in 1 2 3 4 5 6 7 8; do
  out=$(tools/grun.sh "$1" "$2" $3 $4 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"; exit 0
done
echo "$out"
