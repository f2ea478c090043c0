export EVP_B200_MESH_CACHE=/tmp/evp_cache
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 python bench.py > gpurun_out/bench_qu7.5_final.json 2> gpurun_out/bench_qu7.5_final.err; echo "bench rc=$?"; cat gpurun_out/bench_qu7.5_final.json
timeout 300 $B > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_qu7.5_final.csv $B > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 $B > gpurun_out/plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:evp_ -s 40 -c 4 -o gpurun_out/prof_qu7.5_final -f $B > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
