timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in C1 C2 C3; do
  python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline --no-predict-leg 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'acc',d['e2e']['accepted'],'gibbs',round(d['device_ms_per_step']['gibbs'],4))"; done
python tools/timeline.py C1 24 1e-4 2>&1 | grep -v Warn | grep -A40 "rejected"
