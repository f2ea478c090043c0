# last GPU job of round 2: the transport options on the device, a few seconds of work (no torch, no pytest)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader > gpurun_out/options_gpu.txt 2>&1
timeout 150 python tools/gpu_check_options.py > gpurun_out/options_check.log 2>&1; echo "options rc=$?"; tail -5 gpurun_out/options_check.log
