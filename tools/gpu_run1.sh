mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
bash tools/prof_one.sh BC7 bc7_encode bc7_r2h
