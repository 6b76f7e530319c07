# End-of-round measurement pass on one B200 (run through gpurun): GPU tests, every bench workload, then the ncu passes.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; echo rc=$? >> gpurun_out/gputest.log
python bench.py > gpurun_out/r2_v6_bench_default.json 2> gpurun_out/r2_v6_bench_default.err
for w in cfg1 cfg2 cfg3noise cfg4 cfg4lowflux cfg5 cfg5psf; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2_v6_bench_$w.json 2> gpurun_out/r2_v6_bench_$w.err; done
python bench.py --workload cfg2 --wfs pyramid --envs 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v6_bench_cfg2_pyramid.json 2> gpurun_out/r2_v6_bench_cfg2_pyramid.err
python bench.py --workload cfg2 --policy po4ao --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_v6_bench_cfg2_po4ao.json 2> gpurun_out/r2_v6_bench_cfg2_po4ao.err
bash tools/measure_ncu.sh
tail -3 gpurun_out/gputest.log
