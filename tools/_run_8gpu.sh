# Developer script (gpurun --gpus 8): 8-GPU bench lines of the meshed-grid configs -> gpurun_out/r02z_bench_*_8gpu.json
cd /root/repo
for c in ed64k ls_dyn; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config $c --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench8_$c.err | grep '^{' | tail -1 > gpurun_out/r02z_bench_${c}_8gpu.json
cut -c1-200 gpurun_out/r02z_bench_${c}_8gpu.json
done
