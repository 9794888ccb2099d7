cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
: > gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_c_abi.py -q -m gpu -x --timeout 300 > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 300 > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/rc.txt
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_tokrows.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -q -m gpu --timeout 900 > gpurun_out/t_train.log 2>&1; echo "train+model rc=$?" >> gpurun_out/rc.txt
tail -30 gpurun_out/t_kernels.log; tail -30 gpurun_out/t_gemm.log; tail -60 gpurun_out/t_train.log; cat gpurun_out/rc.txt
