# 2-GPU job: parity suite incl. the multi-rank tests, then the halo exchange variants side by side.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_job_n2.sh'
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for wl in qu60 qu15; do
  for halo in p2p nccl; do
    timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl --halo $halo --no-cpu-baseline \
        > gpurun_out/bench_${wl}_n2_${halo}.json 2> gpurun_out/bench_${wl}_n2_${halo}.err; echo "bench $wl $halo rc=$?"
    cat gpurun_out/bench_${wl}_n2_${halo}.json; tail -3 gpurun_out/bench_${wl}_n2_${halo}.err
  done
done
timeout 300 python bench.py --steps 5 --warmup 3 --workload qu60 --no-cpu-baseline > gpurun_out/bench_qu60_n1.json 2> gpurun_out/bench_qu60_n1.err; cat gpurun_out/bench_qu60_n1.json
