mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; tail -2 gpurun_out/r2_bench_i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_i.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["langevin_rk_kpm"]["steps_per_s"], d["pcg_kpm"]["one_persistent_kernel"], d["tau_sharded"]["pcg_kpm"]["us_per_iter"])
PY
