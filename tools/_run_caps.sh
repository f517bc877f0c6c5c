cd /root/repo
for sc in 2 4 1; do echo -n "score_cap=$sc: "; OPFG_SCORE_CAP=$sc timeout 200 python tests/_time_aux.py 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_features.py tests/test_islands.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 100 --warmup 10 2>/dev/null | tail -1 | cut -c1-420
