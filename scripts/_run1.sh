set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s5_pytest.log
timeout 300 python scripts/profile_kernels.py --reps 20 --time --graph --only lik > gpurun_out/s5_kt_a.jsonl 2>&1; cat gpurun_out/s5_kt_a.jsonl
timeout 300 python scripts/profile_kernels.py --reps 20 --time --graph --only 128,128 > gpurun_out/s5_kt_b.jsonl 2>&1; cat gpurun_out/s5_kt_b.jsonl
