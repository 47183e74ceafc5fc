cd $GRAFT_REPO_ROOT
b() { name=$1; shift
  timeout 200 python bench.py --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2z_$name.json 2> gpurun_out/r2z_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r2z_$name.json')); print('N=1 $name', round(d['value']), round(d['ms_per_step'],4), {k:v for k,v in d['roofline']['kernels_ms_per_step'].items() if 'match' in k or 'pack_q' in k})" || tail -3 gpurun_out/r2z_$name.err
}
b cfg3_ts --workload cfg3_synapse_ct --scaling strong --steps 30 --warmup 3
b cfg3_packed --workload cfg3_synapse_ct --scaling strong --steps 30 --warmup 3 --algo 2
b cfg3_ts_l2 --workload cfg3_synapse_ct --scaling strong --steps 30 --warmup 3 --lanes 2
b cfg3_packed_l2 --workload cfg3_synapse_ct --scaling strong --steps 30 --warmup 3 --algo 2 --lanes 2
b cfg5_ts --workload cfg5_stress_vitl --scaling strong --steps 12 --warmup 3 --lanes 2
b cfg5_packed --workload cfg5_stress_vitl --scaling strong --steps 12 --warmup 3 --lanes 2 --algo 2
b cfg4_ts --workload cfg4_polyp_1024 --scaling strong --steps 30 --warmup 3
b cfg4_packed --workload cfg4_polyp_1024 --scaling strong --steps 30 --warmup 3 --algo 2
