mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r1f_pytest.log
for v in A B; do
  if [ $v = A ]; then unset B200LIC_LIB; else export B200LIC_LIB=$PWD/build/variants/$v/libb200lic.so; fi
  timeout 300 python bench.py --skip-cpu --steps 20 --warmup 3 2> gpurun_out/ab_$v.err > gpurun_out/ab_$v.json
  python -c "
import sys, json
d = json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
print('$v', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'roofline ms', round(d['roofline']['ms_per_launch'], 4), 'fwd', round(d['fwd_mpx_s'], 1), round(d['fwd_mpx_s_2k'], 1))
"
  tail -1 gpurun_out/ab_$v.err
done
unset B200LIC_LIB
timeout 200 python scripts/profile_kernels.py --reps 10 --time --graph --only gdn > gpurun_out/r1f_gdn.jsonl 2>&1; cat gpurun_out/r1f_gdn.jsonl
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc2_gather_gemm|tc_wgrad|factorized_lik_kernel|nhwc_split|lp_loss" -c 16 -f -o gpurun_out/r1f_full python scripts/profile_kernels.py --reps 1 > gpurun_out/r1f_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r1f_full.ncu-rep
