# bench line + per-launch table (stderr) for one step
cd $GRAFT_REPO_ROOT
tag=${1:-x}
AMB_BENCH_DUMP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/dump_$tag.txt
tail -c 1500 gpurun_out/bench_$tag.json
