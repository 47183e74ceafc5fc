cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -q -x 2>&1 | tail -3
b() { name=$1; shift
  timeout 200 python bench.py --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2z_$name.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2z_$name.json')); print('N=1 $name', round(d['value']), round(d['ms_per_step'],4), {k:v for k,v in d['roofline']['kernels_ms_per_step'].items() if 'match' in k})"
}
b cfg2 --steps 300 --warmup 5
b cfg3 --workload cfg3_synapse_ct --scaling strong --steps 30 --warmup 3
b cfg5 --workload cfg5_stress_vitl --scaling strong --steps 12 --warmup 3
b cfg4 --workload cfg4_polyp --scaling strong --steps 30 --warmup 3
