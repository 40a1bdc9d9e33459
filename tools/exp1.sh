mkdir -p gpurun_out
for v in 3 6 7 8; do
  echo "== variant $v"; MCT_K2_VARIANT=$v python bench.py --steps 2 --warmup 2 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.3e ms/step %.1f k2_ms %.1f k1_ms %.2f e2e %.3e layer_steps/s %.3e launches %d' % (d['value'], d['ms_per_step'], r['k2_ms_per_step'], r['k1_ms_per_step'], d['e2e']['value'], r['layer_steps_per_s'], d['gpu_launches']))
    else: print(l.rstrip())
"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
