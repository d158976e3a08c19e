# round-2 final evidence (1 GPU): launch list of one step (time + DRAM bytes + tensor-pipe %), section captures of the HBM-bound
# kernels, final bench line (with sweep + GPU reference + CPU baseline), mask-ratio sweep
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "lean_zero or zero_shell or stem" 2>&1 | tail -2
echo "=== ncu launch list (one eager step)"
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python tests/ncu_target_step.py > gpurun_out/ncu_e.log 2>&1
tail -1 gpurun_out/ncu_e.log; wc -l gpurun_out/r2f_launches.csv
echo "=== ncu sections, HBM-bound kernels of one step"
timeout 1800 ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats \
    -k regex:'apply|reduce|patch_loss|hard_mask|ema_dev|adamw|sumsq|stem_|proj_|zero_shell|repack_batched|add_kernel|add_parity0|active_list' \
    -o gpurun_out/r2f_hbm python tests/ncu_target_step.py > gpurun_out/ncu_f.log 2>&1
tail -1 gpurun_out/ncu_f.log; ls -la gpurun_out/r2f_hbm.ncu-rep
echo "=== bench (final line)"
timeout 1200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 1500 gpurun_out/r2f_bench.json
echo "=== bench dump"
AMB_BENCH_DUMP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r2f_dump.txt; grep -c LAUNCH gpurun_out/r2f_dump.txt
