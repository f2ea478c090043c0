# N-GPU scaling job (N = $1): strong scaling on qu7.5 and QU60 with both halo exchanges, weak scaling (planar hex,
# 1 310 720 cells per GPU).  One JSON line per run under gpurun_out/scale_*; a summary line each on stdout.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_job_scale.sh 8'
N=$1
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
run() {   # tag, extra args...
  tag=$1; shift
  timeout 900 $TR bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/scale_${tag}_n$N.json 2> gpurun_out/scale_${tag}_n$N.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_${tag}_n$N.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("${tag} N=$N rc=$rc", "value", round(d["value"],1), "us/subcycle", round(1e3*r["graph_ms_per_subcycle"],2), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "cells/s/gpu", "%.3e" % d["cell_updates_per_sec_per_gpu"], "checksum", d["checksum"] and d["checksum"]["value"], d["halo_exchange"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("${tag} N=$N rc=$rc: no result:", e); print(open("gpurun_out/scale_${tag}_n$N.err").read()[-1500:])
PY
}
for job in ${JOBS:-qu75_p2p qu75_nccl qu60_p2p qu60_nccl weak_p2p}; do
  case $job in
    qu75_p2p)  run qu7.5_p2p  --workload qu7.5 --halo p2p  --steps 10 --warmup 3 ;;
    qu75_nccl) run qu7.5_nccl --workload qu7.5 --halo nccl --steps 10 --warmup 3 --no-e2e ;;
    qu60_p2p)  run qu60_p2p   --workload qu60  --halo p2p  --steps 20 --warmup 5 ;;
    qu60_nccl) run qu60_nccl  --workload qu60  --halo nccl --steps 20 --warmup 5 --no-e2e ;;
    qu15_p2p)  run qu15_p2p   --workload qu15  --halo p2p  --steps 10 --warmup 3 ;;
    qu15_nccl) run qu15_nccl  --workload qu15  --halo nccl --steps 10 --warmup 3 --no-e2e ;;
    weak_p2p)  run weak_p2p   --scaling weak   --halo p2p  --steps 10 --warmup 3 ;;
    weak_nccl) run weak_nccl  --scaling weak   --halo nccl --steps 10 --warmup 3 --no-e2e ;;
  esac
done
