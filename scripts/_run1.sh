set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s5_pytest.log
B200LIC_TC_DEBUG=3 timeout 120 python scripts/tc_timeline.py gdn > gpurun_out/s5_timeline_gdn.txt 2>&1; tail -25 gpurun_out/s5_timeline_gdn.txt
B200LIC_TC_DEBUG=3 timeout 120 python scripts/tc_timeline.py conv > gpurun_out/s5_timeline_conv.txt 2>&1; tail -25 gpurun_out/s5_timeline_conv.txt
timeout 300 python scripts/profile_kernels.py --reps 20 --time --graph > gpurun_out/s5_kernel_times_graph.jsonl 2>&1; cat gpurun_out/s5_kernel_times_graph.jsonl
