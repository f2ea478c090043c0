# 1-GPU: weak-scaling base point, headline bench with CPU variants, launch list + full capture of the EVP kernels on qu7.5
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
JOBS="weak_p2p qu60_p2p" bash tools/gpu_job_scale.sh 1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_qu7.5_r02b.json 2> gpurun_out/bench_qu7.5_r02b.err; echo "bench qu7.5 rc=$?"; cat gpurun_out/bench_qu7.5_r02b.json; tail -3 gpurun_out/bench_qu7.5_r02b.err
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-checksum"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_qu7.5_r02.csv $B > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:evp_ -s 40 -c 4 -o gpurun_out/prof_qu7.5_r02 -f $B > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
