mkdir -p gpurun_out
export EVP_B200_MESH_CACHE=/tmp/evp_cache
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_reconstruct|k_triangles|k_fluxes|k_update|k_prepare" -c 5 -o gpurun_out/ir_prof_qu60_r02f -f python tools/ir_bench.py --level 7 --steps 1 --warmup 0 > gpurun_out/ir_ncu_full.log 2>&1; echo "ir ncu full rc=$?"
