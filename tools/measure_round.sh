set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; echo rc=$? >> gpurun_out/gputest.log
python bench.py > gpurun_out/r2_v6_bench_default.json 2> gpurun_out/r2_v6_bench_default.err
for w in cfg1 cfg2 cfg3noise cfg4 cfg4lowflux cfg5 cfg5psf; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2_v6_bench_$w.json 2> gpurun_out/r2_v6_bench_$w.err; done
python bench.py --workload cfg2 --wfs pyramid --envs 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v6_bench_cfg2_pyramid.json 2> gpurun_out/r2_v6_bench_cfg2_pyramid.err
python bench.py --workload cfg2 --policy po4ao --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_v6_bench_cfg2_po4ao.json 2> gpurun_out/r2_v6_bench_cfg2_po4ao.err
K='regex:aoenv|gemm_tc'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2_v6_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l3.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2_v6_launches_cfg5.csv python bench.py --workload cfg5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l5.log 2>&1
AOENV_PROFILE_REGION=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$K" -c 40 -o gpurun_out/r2_v6_step -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/gputest.log
