mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_kpm_shard_launches.csv python scripts/prof_kpm_shard.py 64 400 > gpurun_out/prof_kpm_shard.log 2>&1
tail -3 gpurun_out/prof_kpm_shard.log
grep -c "gpu__time_duration" gpurun_out/r2_kpm_shard_launches.csv
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_targets.py shardkpm > gpurun_out/memcheck_shardkpm.log 2>&1; tail -4 gpurun_out/memcheck_shardkpm.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_targets.py shardkpm > gpurun_out/racecheck_shardkpm.log 2>&1; tail -3 gpurun_out/racecheck_shardkpm.log
