mkdir -p gpurun_out
for v in A B A; do
  if [ $v = A ]; then unset B200LIC_LIB; else export B200LIC_LIB=$PWD/build/variants/$v/libb200lic.so; fi
  timeout 300 python bench.py --skip-cpu --skip-fwd --steps 20 --warmup 3 2> gpurun_out/ab_$v.err > gpurun_out/ab_$v.json
  python -c "
import sys, json
d = json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
print('$v', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'roofline ms', round(d['roofline']['ms_per_launch'], 4))
"
done
