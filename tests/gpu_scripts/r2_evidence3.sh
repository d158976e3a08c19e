# round-2 final evidence (1 GPU, final build): GPU suite, launch list of one step (time + DRAM bytes + tensor-pipe %), ncu --set full
# of the two N-stacked conv kernels, section captures of the HBM-bound kernels, final bench line, per-launch layer table
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-300 | tail -8
echo "=== ncu launch list (one eager step)"
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python tests/ncu_target_step.py > gpurun_out/ncu_g.log 2>&1
tail -1 gpurun_out/ncu_g.log; wc -l gpurun_out/r2g_launches.csv
echo "=== ncu full: igemm4t (convT fwd 64->64 @64^3)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm4t -c 2 -o gpurun_out/r2g_igemm4t python tests/ncu_target_convT.py > gpurun_out/ncu_g2.log 2>&1; tail -1 gpurun_out/ncu_g2.log
echo "=== ncu full: igemm4 64->64 and 128->128"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm4_kernel -c 2 -o gpurun_out/r2g_igemm4_64x64 python tests/ncu_target.py > gpurun_out/ncu_g3.log 2>&1; tail -1 gpurun_out/ncu_g3.log
AMB_NT_CI=128 AMB_NT_CO=128 AMB_NT_S=64 timeout 600 ncu --set full --clock-control none -k regex:igemm4_kernel -c 2 -o gpurun_out/r2g_igemm4_128x128 python tests/ncu_target.py > gpurun_out/ncu_g4.log 2>&1; tail -1 gpurun_out/ncu_g4.log
echo "=== ncu sections, HBM-bound kernels of one step"
timeout 1800 ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats \
    -k regex:'apply|reduce|patch_loss|hard_mask|ema_dev|adamw|sumsq|stem_|proj_|zero_shell|repack_batched|add_kernel|add_parity0|active_list' \
    -o gpurun_out/r2g_hbm python tests/ncu_target_step.py > gpurun_out/ncu_g5.log 2>&1
tail -1 gpurun_out/ncu_g5.log; ls -la gpurun_out/*.ncu-rep | tail -5
echo "=== bench (final line)"
timeout 1200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 600 gpurun_out/r2g_bench.json
echo "=== bench dump"
AMB_BENCH_DUMP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r2g_dump.txt; grep -c LAUNCH gpurun_out/r2g_dump.txt
timeout 300 python tests/step_profile.py r2g_step 2>&1 | tail -2
