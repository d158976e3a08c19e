# re-run of the build-dependent part of r2_evidence3.sh (the library of that run still carried a reverted per-tap experiment):
# smoke, launch list, final bench line, per-launch layer table, step profile
cd $GRAFT_REPO_ROOT
timeout 600 python __graft_entry__.py smoke 2>&1 | grep smoke
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or transpose or spark_step or anatomask or golden or graph" 2>&1 | tail -2
echo "=== ncu launch list (one eager step)"
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python tests/ncu_target_step.py > gpurun_out/ncu_g.log 2>&1
tail -1 gpurun_out/ncu_g.log; wc -l gpurun_out/r2g_launches.csv
echo "=== bench (final line)"
timeout 1200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; python -c "import sys,json; d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'])"
echo "=== bench dump"
AMB_BENCH_DUMP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r2g_dump.txt; grep -c LAUNCH gpurun_out/r2g_dump.txt
timeout 300 python tests/step_profile.py r2g_step 2>&1 | tail -2
echo "=== second bench (no extras)"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])"
