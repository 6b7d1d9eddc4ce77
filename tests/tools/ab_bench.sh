#!/bin/bash
# Long-run A/B on ONE box: the main bench block only (20 volumes, the part reaches its power-capped steady state), alternating
# environment settings.  usage: ab_bench.sh "ENV=.." "ENV=.." ...   ("-" = no setting)
for round in 1 2; do
  for S in "$@"; do
    [ "$S" = "-" ] && envs="" || envs="$S"
    echo -n "$S: "
    env $envs timeout 600 python bench.py --steps 20 --warmup 3 --cpu-baseline 0 --incumbent 0 --ensemble 0 --big-volume 0 --alt-dtype 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f vol/s  %.2f ms  e2e %.3f  kernel %.1f TFLOP/s  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks']))"
  done
done
