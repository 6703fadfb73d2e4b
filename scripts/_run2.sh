set -x
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/r1f_bench_n1.json; tail -3 gpurun_out/r1f_bench_n1.err
timeout 300 python scripts/profile_kernels.py --reps 10 --time --graph > gpurun_out/r1f_kernel_times.jsonl 2>&1; cat gpurun_out/r1f_kernel_times.jsonl
timeout 300 python scripts/profile_fwd.py --arch mbt2018-mean --hw 512x768 --hw 1365x2048 > gpurun_out/r1f_fwd_times.jsonl 2> gpurun_out/r1f_fwd_times.err; cat gpurun_out/r1f_fwd_times.jsonl; tail -3 gpurun_out/r1f_fwd_times.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r1f_launches_bench.csv python bench.py --steps 1 --warmup 1 --skip-cpu --skip-fwd > gpurun_out/r1f_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1f_launches_fwd_mbt_2k.csv python scripts/profile_fwd.py --arch mbt2018-mean --hw 1365x2048 --eager --reps 1 > gpurun_out/r1f_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
