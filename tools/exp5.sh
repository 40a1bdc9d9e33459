for b in 32 64 128; do for v in 7 8; do
  echo "== batch $b variant $v"; MCT_K2_VARIANT=$v python bench.py --steps 2 --warmup 2 --no-cpu --batch $b 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.3e ms/step %.1f k2_ms %.1f layer_steps/s %.3e' % (d['value'], d['ms_per_step'], r['k2_ms_per_step'], r['layer_steps_per_s']))
"
done; done
