# what the driver runs at round end, on one GPU: tests, smoke, both bench arms
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/BENCH_reference.json 2> gpurun_out/BENCH_reference.err; tail -n 3 gpurun_out/BENCH_reference.err
( time python bench.py ) > gpurun_out/BENCH_ours.json 2> gpurun_out/BENCH_ours.err; tail -n 3 gpurun_out/BENCH_ours.err
