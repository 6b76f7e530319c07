# launch lists (cfg3, cfg5) and one --set full capture of the kernels of two timed steps; run only after the plain commands exited 0
K='regex:atm_|shwfs_|dm_|observe|command|envmax|gemm_|split_bf16|psf_'
AOENV_PROFILE_REGION=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_v6_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l3.log 2>&1
AOENV_PROFILE_REGION=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2_v6_launches_cfg5.csv python bench.py --workload cfg5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l5.log 2>&1
AOENV_PROFILE_REGION=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -c 40 -o gpurun_out/r2_v6_step -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
