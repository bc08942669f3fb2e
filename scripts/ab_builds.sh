#!/bin/bash
# Same-box A/B of several builds of the library on the scan's phase profile: the in-tree build against the variants under
# memex_b200/_lib/<name>/ (an earlier commit built with its own build.py, or scan_tc.cu compiled with -DMX_TC_<name>).
#   gpurun --timeout 300 -- bash scripts/ab_builds.sh old NO_APPEND PREP4 NO_FINE
set -u
for rep in 1 2; do
  for v in "$@" new; do
    if [ "$v" = new ]; then unset MX_B200_LIB; else export MX_B200_LIB=$PWD/memex_b200/_lib/$v/libmemex_b200.so; fi
    python scripts/scan_tc_prof.py 2>&1 | awk -v v="$v" '/ rows /{r=$1} /remaining|total first/{printf "%-10s rows %-8s %s\n", v, r, $0}'
  done
done
