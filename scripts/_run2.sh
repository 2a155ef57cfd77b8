N=$1
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q -k "nccl" 2>&1 | tail -8
for cfg in "64 100" "64 200" "32 200" "64 400"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_p2p.py $cfg 2>&1 | grep -E "^\{|rror" | head -5
done
