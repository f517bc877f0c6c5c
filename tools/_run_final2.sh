cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02z_gpu_tests.log; cat gpurun_out/r02z_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
timeout 600 compute-sanitizer --tool memcheck python tests/_sanitizer_meshed.py 2>&1 | tail -5 > gpurun_out/r02z_memcheck_meshed.log; cat gpurun_out/r02z_memcheck_meshed.log
