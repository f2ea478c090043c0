# First device run of the incremental-remap transport.  One gpurun call, ~6 min:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_job_ir.sh'
# 1. parity: every `cuda` case of tests/test_ir_parity.py + tests/test_ir_blocks.py (bit-exact against the oracle)
# 2. timing: tools/ir_bench.py on QU60 (level 7) and QU15 (level 9), the oracle on QU60 beside it
# 3. the launch list of one QU15 step (ncu, after the plain run has exited 0)
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
IR_B200_CUDA_LEG=1 timeout 600 python -m pytest -q -x -m gpu -k "cuda and not child_process" -p no:cacheprovider \
    tests/test_ir_parity.py tests/test_ir_blocks.py > gpurun_out/ir_parity.log 2>&1; echo "ir parity rc=$?"; tail -5 gpurun_out/ir_parity.log
timeout 300 python tools/ir_bench.py --level 7 --steps 5 --warmup 2 --check > gpurun_out/ir_bench_qu60.json 2> gpurun_out/ir_bench_qu60.err; echo "bench qu60 rc=$?"; cat gpurun_out/ir_bench_qu60.json
timeout 300 python tools/ir_bench.py --level 7 --steps 2 --warmup 1 --cpu > gpurun_out/ir_bench_qu60_cpu.json 2>&1; cat gpurun_out/ir_bench_qu60_cpu.json
timeout 600 python tools/ir_bench.py --level 9 --steps 3 --warmup 1 > gpurun_out/ir_bench_qu15.json 2> gpurun_out/ir_bench_qu15.err; echo "bench qu15 rc=$?"; cat gpurun_out/ir_bench_qu15.json
IR_B200_PIN_HOST=1 timeout 600 python tools/ir_bench.py --level 9 --steps 3 --warmup 1 > gpurun_out/ir_bench_qu15_pinned.json 2> gpurun_out/ir_bench_qu15_pinned.err; echo "bench qu15 pinned rc=$?"; cat gpurun_out/ir_bench_qu15_pinned.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ir_launches_qu15.csv \
    python tools/ir_bench.py --level 9 --steps 1 --warmup 1 > gpurun_out/ir_ncu_list.log 2>&1; echo "ncu list rc=$?"
