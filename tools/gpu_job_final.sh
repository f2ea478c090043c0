# final 1-GPU check of the tree: GPU parity suite, smoke, one short bench through every leg, transport e2e with and
# without page-locked host arrays
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_final.log
[ -n "$SKIP_BENCH" ] || { timeout 200 python bench.py --workload qu60 --steps 5 --warmup 3 > gpurun_out/bench_qu60_final.json 2> gpurun_out/bench_qu60_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_qu60_final.json; }
timeout 120 python tools/ir_bench.py --level 7 --steps 10 --warmup 3 > gpurun_out/ir_bench_qu60_final.json 2> gpurun_out/ir_bench_qu60_final.err; cat gpurun_out/ir_bench_qu60_final.json
IR_B200_PIN_HOST=1 timeout 120 python tools/ir_bench.py --level 7 --steps 10 --warmup 3 > gpurun_out/ir_bench_qu60_final_pinned.json 2> gpurun_out/ir_bench_qu60_final_pinned.err; cat gpurun_out/ir_bench_qu60_final_pinned.json
