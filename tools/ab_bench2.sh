#!/bin/bash
# A/B of library variants and environment knobs on the headline bench:
#   tools/ab_bench2.sh "name[:ENV=VAL]" ...      ("main" = the regular build)
for rep in 1 2; do for spec in "$@"; do
  v="${spec%%:*}"; envs=""; [ "$spec" != "$v" ] && envs="${spec#*:}"
  if [ "$v" = main ]; then unset SBD_LIB_PATH; else export SBD_LIB_PATH=tools/experiments/libsbd_$v.so; fi
  env $envs SBD_SKIP_BUILD_ID_CHECK=1 python bench.py --steps 10 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$spec', round(d['value']), round(d['e2e']['value']), round(d['e2e_host_buffers']['value']), d['bad_bins'])"
done; done
