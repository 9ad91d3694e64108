python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2d.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu_r2d.log
CMD="python bench.py --steps 100 --warmup 5 --no-cpu --no-solve --no-strips --no-ensemble --no-tables"
$CMD > gpurun_out/bench_r2f_short.json 2> gpurun_out/bench_r2f_short.err; echo short bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_r2f_short.json')); print('e2e', d['e2e']['ms_per_call'], 'kernel', d['roofline']['kernel_ms'], d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo ncu launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:residual_ring -s 60 -c 3 -o gpurun_out/r2_residual_ring -f $CMD > gpurun_out/ncu_f.log 2>&1; echo ncu full rc=$?
ls -la gpurun_out/*.ncu-rep | tail -3
