#!/bin/bash
# A/B of library variants on the headline bench: tools/ab_bench.sh name1 name2 ... ("main" = the regular build)
for rep in 1 2; do for v in "$@"; do
  if [ "$v" = main ]; then unset SBD_LIB_PATH; else export SBD_LIB_PATH=tools/experiments/libsbd_$v.so; fi
  python bench.py --steps 10 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v', round(d['value']), round(d['e2e']['value']), round(d['e2e_host_buffers']['value']))"
done; done
