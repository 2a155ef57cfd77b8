timeout 600 python -m pytest tests/test_gpu_cgpipe.py -x -q 2>&1 | tail -15
timeout 300 python scripts/bench_cgpipe.py quick > gpurun_out/cgpipe_bench.jsonl 2> gpurun_out/cgpipe_bench.err; tail -5 gpurun_out/cgpipe_bench.err; cat gpurun_out/cgpipe_bench.jsonl
