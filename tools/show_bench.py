#!/usr/bin/env python
"""Pretty-print the JSON line(s) bench.py wrote to a file."""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'n_gpus')}, 'e2e', round(d['e2e']['value'], 2),
          'it', d['config'].get('iterations'), 'levels', d['config'].get('levels'))
    print('  roofline', d['roofline'])
    for k, v in d['kernels'].items():
        print('  %-12s n %5d  %9.2f us  %8s GB/s  share %.3f' % (k, v['launches'], v['us_per_launch'], v.get('gbs', v.get('gbs_all_ranks')), v['share']))
    if d.get('cpu_baseline'): print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
