# 1-GPU job: parity suite, small-mesh persistent kernel vs the two-kernel graph, the headline bench, IR profile.
mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n1.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_n1.log
for wl in square qu240; do
  for pers in 1 0; do
    EVP_B200_PERSISTENT=$pers timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_${wl}_pers$pers.json 2> gpurun_out/bench_${wl}_pers$pers.err; echo "bench $wl persistent=$pers rc=$?"
    python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${wl}_pers$pers.json") if l.startswith("{")][-1]); r=d["roofline"]
print("$wl persistent=$pers", "value", round(d["value"],1), "us/subcycle", round(1e3*r["graph_ms_per_subcycle"],3), "launches", d["gpu_launches"], "parity", d["parity"] and d["parity"]["bit_exact"], "checksum", d["checksum"]["value"], "e2e", round(d["e2e"]["value"],1))
PY
  done
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_qu7.5_r02a.json 2> gpurun_out/bench_qu7.5_r02a.err; echo "bench qu7.5 rc=$?"; cat gpurun_out/bench_qu7.5_r02a.json; tail -4 gpurun_out/bench_qu7.5_r02a.err
# IR transport: timing, launch list, full capture of the four step kernels (QU60, 115 rows)
timeout 300 python tools/ir_bench.py --level 7 --steps 5 --warmup 2 > gpurun_out/ir_bench_qu60_r02a.json 2> gpurun_out/ir_bench_qu60_r02a.err; echo "ir bench rc=$?"; cat gpurun_out/ir_bench_qu60_r02a.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ir_launches_qu60_r02a.csv \
    python tools/ir_bench.py --level 7 --steps 1 --warmup 1 > gpurun_out/ir_ncu_list.log 2>&1; echo "ir ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_reconstruct|k_triangles|k_fluxes|k_update|k_prepare" -c 5 -o gpurun_out/ir_prof_qu60_r02a -f \
    python tools/ir_bench.py --level 7 --steps 1 --warmup 0 > gpurun_out/ir_ncu_full.log 2>&1; echo "ir ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
