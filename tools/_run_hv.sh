cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_convergence_boundary.py tests/test_lane_kernel.py tests/test_q_limits.py tests/test_dynamic_branches.py tests/test_islands.py tests/test_security_constrained.py tests/test_wrappers_mixed.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tests/_time_hv_threads.py 128 2>&1 | tail -1 | cut -c1-40
timeout 300 python tests/_time_quick.py 1-HV-urban--0-sw 8192 128 2>&1 | tail -2
for c in mixed ls_dyn; do timeout 300 python bench.py --config $c --steps 50 --warmup 5 2>/dev/null | tail -1 | cut -c1-200; done
timeout 600 compute-sanitizer --tool racecheck python tests/_sanitizer_meshed.py 2>&1 | tail -6
