# Developer script (gpurun): full GPU test suite, smoke, bench lines of the four configs and the reference arm -> gpurun_out/
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02z_gpu_tests.log; cat gpurun_out/r02z_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in vc32k ed64k mixed ls_dyn; do timeout 600 python bench.py --config $c 2>gpurun_out/bench_$c.err | tail -1 > gpurun_out/r02z_bench_$c.json; cut -c1-170 gpurun_out/r02z_bench_$c.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
