# Timing experiments for the fused vertex kernel + halo exchange (2 GPUs); EVP_B200_P2P_DEBUG variants give INVALID results.
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for dbg in 0 1 2 4 8; do
  EVP_B200_P2P_DEBUG=$dbg timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --workload ${WL:-qu15} --halo p2p --no-cpu-baseline --no-e2e --no-checksum \
      > gpurun_out/dbg_$dbg.json 2> gpurun_out/dbg_$dbg.err; echo "dbg $dbg rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/dbg_$dbg.json") if l.startswith("{")][-1]); r=d["roofline"]
print("dbg $dbg", "value", round(d["value"],1), "graph us/subcycle", round(1e3*r["graph_ms_per_subcycle"],2), "cell", round(1e3*r["kernel_ms"],2), "vertex", round(1e3*r["vertex_kernel_ms"],2))
PY
done
