#!/bin/sh
# compute-sanitizer passes over the small sanity problem -- the reference's tools/debug/check_race.sh:3-4
# runs racecheck only; memcheck and synccheck (mbarrier misuse) are added for the TMA / tcgen05 kernel.
# Usage: tools/debug/check_race.sh [kernel index]      (KERNELS=tune to cover both machine mappings)
cd "$(dirname "$0")/../.."
K=${1:--1}
LOG=$(mktemp)
rc=0
for TOOL in racecheck memcheck synccheck; do
  echo "== compute-sanitizer --tool $TOOL"
  compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/debug/sanity_check.py --kernel=$K --small > "$LOG" 2>&1 || rc=1
  tail -6 "$LOG"
done
rm -f "$LOG"
exit $rc
