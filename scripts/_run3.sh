# round-2 tau-sharded KPM-PCG / SSH slabs: tests + timings.  usage: scripts/_run3.sh N [nopytest]
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q -k "kpm" 2>&1 | tail -5 | tee gpurun_out/r2_pytest_sharded_kpm_1gpu.log
  for cfg in "32 200" "64 400" "64 400 --fused"; do
    timeout 300 python scripts/bench_sharded_pcg.py $cfg 2>&1 | grep -E "^\{|rror" | head -5 | tee -a gpurun_out/r2_bench_sharded_pcg_1gpu.jsonl
  done
else
  if [ "$2" != "nopytest" ]; then
    timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q -k "nccl" 2>&1 | tail -5 | tee gpurun_out/r2_pytest_sharded_kpm_nccl_${N}gpu.log
  fi
  rm -f gpurun_out/r2_bench_sharded_pcg_${N}gpu.jsonl
  for fl in "--p2p" "--p2p --fused"; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/bench_sharded_pcg.py 64 400 $fl 2>&1 | grep -E "^\{|rror" | head -5 | tee -a gpurun_out/r2_bench_sharded_pcg_${N}gpu.jsonl
  done
fi
