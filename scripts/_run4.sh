N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q -k "nccl" 2>&1 | tail -4 | tee gpurun_out/r2_pytest_sharded_kpm_nccl_${N}gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/run_sharded_langevin.py 32 200 4 rk 2>&1 | grep -E "^\{|rror" | tee gpurun_out/r2_run_sharded_langevin_${N}gpu.jsonl
timeout 300 python scripts/run_sharded_langevin.py 32 200 4 rk 2>&1 | grep -E "^\{|rror" | tee gpurun_out/r2_run_sharded_langevin_1gpu.jsonl
