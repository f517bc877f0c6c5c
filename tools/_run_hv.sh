cd /root/repo
for m in "" 0 1; do echo "OPFG_DIAG_MODE=$m"; if [ -z "$m" ]; then unset OPFG_DIAG_MODE; else export OPFG_DIAG_MODE=$m; fi
timeout 300 python tests/_time_hv_threads.py 128 2>&1 | tail -1 | cut -c1-40
timeout 300 python bench.py --config ls_dyn --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-160; done
