timeout 600 python -m pytest tests/test_gpu_cgpipe.py tests/test_sharded.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/bench_cgpipe.py > gpurun_out/cgpipe_bench.jsonl 2> gpurun_out/cgpipe_bench.err; tail -5 gpurun_out/cgpipe_bench.err
python - <<'PY'
import json
for ln in open('gpurun_out/cgpipe_bench.jsonl'):
    d=json.loads(ln)
    print(d['lattice'], d['model'], {k:(v['us_per_iter'], v['variant'], v.get('slices_per_cta')) for k,v in d.items() if isinstance(v,dict)})
PY
