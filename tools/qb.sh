#!/bin/bash
# usage: tools/qb.sh "<quick_bench args>" ...   -> one labelled line per configuration
for a in "$@"; do
  python tools/quick_bench.py $a | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['streaming']; c=d['chained']
print('%-48s stream %.2f us %.3f | chained %.2f us %.3f' % (sys.argv[1], s['ms_per_step']*1e3, s['frac_of_6545'], c['ms_per_step']*1e3, c['frac_of_6545']))
" "$a"
done
