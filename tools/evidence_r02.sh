set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; tail -c 600 gpurun_out/r02_bench_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'ss_dr_build|ss_dr_rounds|ss_dr_grow' -s 6 -c 3 -o gpurun_out/r02_full_c5 python tools/dr_time.py 16384 0 1 > gpurun_out/ncu_c5.log 2>&1; tail -2 gpurun_out/ncu_c5.log
ncu --set full --clock-control none -k regex:'ss_dr_build|ss_dr_rounds|ss_dr_grow|ss_diffuse' -s 8 -c 4 -o gpurun_out/r02_full_c4 python tools/dr_time.py 4096 1 1 > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
ncu --set full --clock-control none -k regex:'ss_dr_build|ss_dr_rounds|ss_dr_grow' -s 6 -c 3 -o gpurun_out/r02_full_c5u python tools/dr_time.py 16384 0 1 util=1 > gpurun_out/ncu_c5u.log 2>&1; tail -2 gpurun_out/ncu_c5u.log
