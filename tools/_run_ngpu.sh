# Developer script (gpurun --gpus N): ls_dyn bench line on N GPUs (BASELINE configs[4] sweep)
cd /root/repo
N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config ls_dyn --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench${N}_ls_dyn.err | grep '^{' | tail -1 > gpurun_out/r02z_bench_ls_dyn_${N}gpu.json
cut -c1-200 gpurun_out/r02z_bench_ls_dyn_${N}gpu.json
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_lane_kernel.py -m gpu -x -q 2>&1 | tail -2; fi
