# final 1-GPU check of the tree: GPU parity suite, smoke, one short bench through every leg
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_final.log
timeout 200 python bench.py --workload qu60 --steps 5 --warmup 3 > gpurun_out/bench_qu60_final.json 2> gpurun_out/bench_qu60_final.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_qu60_final.json
