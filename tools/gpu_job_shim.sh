# the Fortran shims executed by the interpreter against the CUDA libraries
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_fortran_shim_executed.py tests/test_refexec_step.py -m gpu -q -rs -p no:cacheprovider > gpurun_out/shim_gpu.log 2>&1; echo "shim rc=$?"; tail -15 gpurun_out/shim_gpu.log
