for c in C3 C4 C1; do
  echo "== $c"; python bench.py --config $c --steps 2 --warmup 3 --cpu-seconds 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.3e ms/step %.1f k2_ms %.1f k1_ms %.3f k1 GB/s %.0f frac %.3f e2e %s cpu %.3e (%d cores) waves %s proposal %.2f ms launches %d' % (d['value'], d['ms_per_step'], r['k2_ms_per_step'], r['k1_ms_per_step'], r['k1_hbm_gbs'], r['frac'], ('%.3e' % d['e2e']['value']) if d['e2e'] else None, d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['waves'], d['proposal_latency']['ms'], d['gpu_launches']))
    else: print(l.rstrip()[-300:])
"
done
