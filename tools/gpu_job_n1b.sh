# 1-GPU job: parity suite (EVP + transport), small-mesh persistent kernel, transport timing and profile.
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_n1b.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_n1b.log
for wl in square qu240; do
  for pers in 1 0; do
    EVP_B200_PERSISTENT=$pers timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_pers$pers.json 2> gpurun_out/bench_${wl}_pers$pers.err; echo "bench $wl persistent=$pers rc=$?"
    python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${wl}_pers$pers.json") if l.startswith("{")][-1]); r=d["roofline"]
print("$wl persistent=$pers", "value", round(d["value"],1), "us/subcycle", round(1e3*r["graph_ms_per_subcycle"],3), "launches", d["gpu_launches"], "checksum", d["checksum"]["value"], "e2e", round(d["e2e"]["value"],1))
PY
  done
done
timeout 600 python tools/ir_bench.py --level 7 --steps 10 --warmup 3 --check --cpu-baseline > gpurun_out/ir_bench_qu60_r02${TAG:-b}.json 2> gpurun_out/ir_bench_qu60_r02${TAG:-b}.err; echo "ir bench qu60 rc=$?"; cat gpurun_out/ir_bench_qu60_r02${TAG:-b}.json; tail -3 gpurun_out/ir_bench_qu60_r02${TAG:-b}.err
[ -n "$SKIP_QU15" ] || timeout 600 python tools/ir_bench.py --level 9 --steps 3 --warmup 1 > gpurun_out/ir_bench_qu15_r02${TAG:-b}.json 2> gpurun_out/ir_bench_qu15_r02${TAG:-b}.err; echo "ir bench qu15 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ir_launches_qu60_r02${TAG:-b}.csv \
    python tools/ir_bench.py --level 7 --steps 1 --warmup 1 > gpurun_out/ir_ncu_list.log 2>&1; echo "ir ncu list rc=$?"
[ -n "$SKIP_FULL" ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_reconstruct|k_triangles|k_fluxes|k_update|k_prepare" -c 5 -o gpurun_out/ir_prof_qu60_r02${TAG:-b} -f \
    python tools/ir_bench.py --level 7 --steps 1 --warmup 0 > gpurun_out/ir_ncu_full.log 2>&1; echo "ir ncu full rc=$?"
