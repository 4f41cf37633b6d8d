set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r01.json 2> gpurun_out/plain_r01.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_ekf20k.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cov_update_multi|k_gate|k_gain_single" -c 7 -f -o gpurun_out/ekf_scan_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_ekf.log 2>&1
python bench.py --workload pf --particles 262144 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_pf_r01.json 2> gpurun_out/plain_pf_r01.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01_pf256k.csv python bench.py --workload pf --particles 262144 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_pf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gather_rows|k_pf_sample_proposal|k_pf_feature_update|k_resample_search|k_scan_level" -c 10 -f -o gpurun_out/pf_r01 python bench.py --workload pf --particles 262144 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_pf.log 2>&1
ls -la gpurun_out/*.ncu-rep
