# 2-GPU job: multi-rank parity tests + the peer-to-peer exchange on QU60 and QU15.
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 600 python -m pytest tests/test_multirank.py -m gpu -x -q > gpurun_out/pytest_multirank_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_multirank_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for wl in qu60 qu15; do
  for halo in ${HALOS:-p2p}; do
    timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl --halo $halo --no-cpu-baseline --no-e2e \
        > gpurun_out/bench_${wl}_n2_${halo}${TAG}.json 2> gpurun_out/bench_${wl}_n2_${halo}${TAG}.err; echo "bench $wl $halo rc=$?"
    python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_n2_${halo}${TAG}.json"))
r=d["roofline"]
print("$wl $halo", "value", round(d["value"],1), "graph us/subcycle", round(1e3*r["graph_ms_per_subcycle"],2), "cell", round(1e3*r["kernel_ms"],2), "vertex", round(1e3*r["vertex_kernel_ms"],2), "checksum", d["checksum"]["value"], d["halo_exchange"])
PY
  done
done
