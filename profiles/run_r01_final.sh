set -x
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_final.log 2>&1; tail -2 gpurun_out/gpu_tests_final.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-400 gpurun_out/bench_final.json
python bench.py --steps 10 --warmup 3 --mode seq --no-cpu-baseline > gpurun_out/bench_seq_final.json 2> gpurun_out/bench_seq_final.err
for wl in 2s3z mmm2 27m; do python bench.py --steps 5 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_final.json 2> gpurun_out/bench_${wl}_final.err; python -c "
import json,sys; j=json.load(open('gpurun_out/bench_${wl}_final.json')); print('$wl', j['value'], j['ms_per_step'], j['e2e']['value'])"; done
python bench.py --steps 5 --warmup 3 --workload matrix --no-cpu-baseline > gpurun_out/bench_matrix_final.json 2> gpurun_out/bench_matrix_final.err; tail -c 300 gpurun_out/bench_matrix_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_final.log 2>&1
tail -2 gpurun_out/ncu_launches_final.log
