# round-2 evidence pass (1 GPU): parity calibration, bench with sweep + GPU reference, mask-ratio sweep, ncu launch list with
# DRAM / tensor-pipe counters for every kernel of the step, ncu --set full of the N-stacked conv kernel
cd $GRAFT_REPO_ROOT
echo "=== spark_report"; timeout 900 python -m tests.model_checks spark_report 2>&1 | grep RESULT
echo "=== bench"; AMB_BENCH_DUMP=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sweep --gpu-reference > gpurun_out/bench_r2b.json 2> gpurun_out/dump_r2b.txt; tail -c 600 gpurun_out/bench_r2b.json; tail -3 gpurun_out/dump_r2b.txt
echo "=== mask sweep"; timeout 600 python tests/conv_bench.py sweep > gpurun_out/r2_mask_sweep.jsonl 2>&1; tail -3 gpurun_out/r2_mask_sweep.jsonl
echo "=== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -s 560 -c 520 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log; wc -l gpurun_out/r2_launches.csv
echo "=== ncu full igemm4"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm4 -c 4 -o gpurun_out/r2_igemm4_64x64 python tests/ncu_target.py > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
AMB_NT_CO=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm4 -c 4 -o gpurun_out/r2_igemm4_64x32 python tests/ncu_target.py > gpurun_out/ncu_d.log 2>&1; tail -2 gpurun_out/ncu_d.log
ls -la gpurun_out/*.ncu-rep | tail -3
