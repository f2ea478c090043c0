mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 200 python -m pytest tests/test_ir_parity.py -m gpu -x -q > gpurun_out/ir_parity_quick.log 2>&1; echo "ir parity rc=$?"; tail -2 gpurun_out/ir_parity_quick.log
timeout 120 python tools/ir_bench.py --level 7 --steps 10 --warmup 3 > gpurun_out/ir_bench_qu60_r02f.json 2> gpurun_out/ir_bench_qu60_r02f.err; cut -c1-420 gpurun_out/ir_bench_qu60_r02f.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ir_launches_qu60_r02f.csv python tools/ir_bench.py --level 7 --steps 1 --warmup 1 > gpurun_out/ir_ncu_list.log 2>&1; echo "ncu list rc=$?"
