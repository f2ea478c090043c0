# Transport path on one B200: parity of every `cuda` case, timing on QU60 / QU15 (115 tracer rows), launch list and full
# ncu capture of the step kernels.  One gpurun call, ~5 min:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_job_ir.sh'
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
IR_B200_CUDA_LEG=1 timeout 600 python -m pytest -q -x -m gpu -k "cuda and not child_process" -p no:cacheprovider \
    tests/test_ir_parity.py tests/test_ir_blocks.py > gpurun_out/ir_parity.log 2>&1; echo "ir parity rc=$?"; tail -5 gpurun_out/ir_parity.log
timeout 600 python tools/ir_bench.py --level 7 --steps 10 --warmup 3 --check --cpu-baseline > gpurun_out/ir_bench_qu60.json 2> gpurun_out/ir_bench_qu60.err; echo "bench qu60 rc=$?"; cat gpurun_out/ir_bench_qu60.json
IR_B200_PIN_HOST=1 timeout 300 python tools/ir_bench.py --level 7 --steps 10 --warmup 3 > gpurun_out/ir_bench_qu60_pinned.json 2> gpurun_out/ir_bench_qu60_pinned.err; cat gpurun_out/ir_bench_qu60_pinned.json
[ -n "$SKIP_QU15" ] || timeout 600 python tools/ir_bench.py --level 9 --steps 3 --warmup 1 > gpurun_out/ir_bench_qu15.json 2> gpurun_out/ir_bench_qu15.err; cat gpurun_out/ir_bench_qu15.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ir_launches_qu60.csv \
    python tools/ir_bench.py --level 7 --steps 1 --warmup 1 > gpurun_out/ir_ncu_list.log 2>&1; echo "ncu list rc=$?"
[ -n "$SKIP_FULL" ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_reconstruct|k_triangles|k_fluxes|k_update|k_prepare" -c 5 \
    -o gpurun_out/ir_prof_qu60 -f python tools/ir_bench.py --level 7 --steps 1 --warmup 0 > gpurun_out/ir_ncu_full.log 2>&1; echo "ncu full rc=$?"
# summaries for profiles/: python tools/ncu_summary.py gpurun_out/ir_prof_qu60.ncu-rep > profiles/ir_rNN_ncu.txt
