N=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q -k "nccl" 2>&1 | tail -4 | tee gpurun_out/r2_pytest_sharded_nccl_${N}gpu.log
for cfg in "64 400"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_p2p.py $cfg 2>&1 | grep -E "^\{|rror" | head -5 | tee -a gpurun_out/r2_bench_p2p_${N}gpu.jsonl
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_line_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_line_${N}gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('tau_sharded'), indent=1)); print('value', d['value'], 'e2e', d['e2e']['value'])
PY
