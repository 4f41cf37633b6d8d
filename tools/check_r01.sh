# round-1 robustness run (1 GPU): full default bench incl. CPU baseline + reference arm, sanitizer pass, PF gather ncu
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -k "batch or scan or single or gating or checkpoint or augment" > gpurun_out/sanitizer_ekf.log 2>&1; echo "sanitizer ekf rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_pf_gpu.py -m gpu -x -q > gpurun_out/sanitizer_pf.log 2>&1; echo "sanitizer pf rc=$?"
tail -3 gpurun_out/sanitizer_ekf.log gpurun_out/sanitizer_pf.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gather_rows|k_resample_search" -c 5 -f -o gpurun_out/pf_gather_r01 python bench.py --workload pf --particles 262144 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_pf2.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -2
