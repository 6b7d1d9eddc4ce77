#!/bin/bash
# A/B of builds / knobs on ONE box: per-role cycle counters of the 128^3 layers (DWMH_TC_DEBUG = 8, and 10 = MMAs off) and
# the wall-clock time of a 32-forward batch, alternating.
# usage: ab_probe.sh "ENV=.. ENV=.. lib.so" "ENV=.. lib2.so" ...    (each argument: optional env assignments, then the library)
run() { local spec="$1"; shift; local lib="${spec##* }"; local envs="${spec% *}"; [ "$envs" = "$spec" ] && envs=""; env $envs DWMH_LIB_PATH=$PWD/$lib "$@"; }
if [ -z "$AB_NO_PROF" ]; then
for S in "$@"; do
  for D in ${AB_DEBUGS:-8 10}; do
    echo "== $S DWMH_TC_DEBUG=$D"
    run "$S" DWMH_TC_DEBUG=$D timeout 300 python tests/tools/profile_forward.py 32 0 2>&1 | grep "C0=32 C1=0 Cout=32\|C0=32 C1=32 Cout=32" | tail -2
  done
done
fi
for i in 1 2 3; do
  for S in "$@"; do echo -n "$S: "; run "$S" timeout 300 python tests/tools/sa_probe.py 32 | tail -1; done
done
