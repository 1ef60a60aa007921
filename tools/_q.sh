TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/run_partition.py 3 20000 2>&1 | grep -v "^W\|warn" | tail -12
ST_PART_GRAPH=0 timeout 600 $TR tools/run_partition.py 2 12000 2>&1 | grep -v "^W\|warn" | tail -4
for g in 1 0; do ST_PART_GRAPH=$g timeout 600 $TR bench.py --gpus 2 --workload C3 --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3 n2 graph=$g value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'acc',d['e2e']['accepted'],d['parity_probe']['loglik_w'])"; done
