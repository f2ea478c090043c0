mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 600 python -m pytest tests/test_multirank.py tests/test_ir_multirank.py -m gpu -x -q > gpurun_out/pytest_multirank_n2c.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multirank_n2c.log
JOBS="qu15_p2p qu15_nccl qu60_p2p qu60_nccl weak_p2p" bash tools/gpu_job_scale.sh 2
timeout 300 python tools/ir_bench.py --level 7 --steps 10 --warmup 3 --check > gpurun_out/ir_bench_qu60_r02e.json 2> gpurun_out/ir_bench_qu60_r02e.err; cat gpurun_out/ir_bench_qu60_r02e.json
