#!/bin/bash
# prints value, ms/step and per-kernel ms / frac from a short bench run
python $(dirname $0)/../bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -3 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('value %.4g  ms/step %.2f' % (d['value'], d['ms_per_step'])); print('  roofline', {k: d['roofline'][k] for k in ('kernel','frac','ms_per_launch','frac_best_isolated','step_frac')})
        for grp in ('kernels', 'alternative_kernels'):
            for k, v in d.get(grp, {}).items(): print('  %-12s %8.3f ms  frac %s  live %s ms frac %s' % (k, v['ms'], v.get('frac'), v.get('live_ms'), v.get('live_frac')))
    else: print(ln.rstrip())
"
