set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_convergence_boundary.py tests/test_lane_kernel.py tests/test_q_limits.py tests/test_ward_impedance.py tests/test_security_constrained.py -m gpu -x -q 2>&1 | tail -5
echo "== default (shared slots, 3 envs)"; timeout 300 python tests/_time_hv_threads.py 128 2>&1 | tail -1
echo "== shared slots, 2 envs all staged"; OPFG_PREFER_ENVS=0 timeout 300 python tests/_time_hv_threads.py 128 2>&1 | tail -1
echo "== no sharing"; OPFG_SHARE_SLOTS=0 timeout 300 python tests/_time_hv_threads.py 128 2>&1 | tail -1
echo "== default T=96?"; timeout 300 python tests/_time_hv_threads.py 64 256 2>&1 | tail -2
timeout 300 python bench.py --config mixed --steps 50 --warmup 5 2>/dev/null | tail -1 | cut -c1-300
