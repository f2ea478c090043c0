# the reference-executed fixtures replayed on the device (no torch needed by these tests)
mkdir -p gpurun_out
timeout 160 python -m pytest tests/test_golden.py tests/test_refexec_init.py tests/test_refexec_step.py -m gpu -q -rs -p no:cacheprovider > gpurun_out/refexec_gpu.log 2>&1; echo "refexec rc=$?"; tail -15 gpurun_out/refexec_gpu.log
