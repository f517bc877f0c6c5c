# Developer script (gpurun): ncu captures behind profiles/r02z_* -- "a": kernels of a VoltageControl step, "b": k_pf_multi of ed64k + per-kernel DRAM traffic of a step
cd /root/repo
if [ "$1" = "a" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pf_tree|k_score_warps|k_assemble_warps' -s 6 -c 3 -f -o gpurun_out/prof_step_r3a python tests/_prof_run.py > gpurun_out/prof_r3a.log 2>&1; tail -2 gpurun_out/prof_r3a.log
else
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pf_multi' -s 4 -c 1 -f -o gpurun_out/prof_ed_r3a python bench.py --config ed64k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_ed_r3a.log 2>&1; tail -2 gpurun_out/prof_ed_r3a.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r03_step_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03_traffic.log 2>&1; tail -1 gpurun_out/r03_traffic.log | cut -c1-100
fi
ls -la gpurun_out
