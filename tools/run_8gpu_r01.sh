# round-1 multi-GPU evidence run (one 8-GPU box): sharded parity at world 4, then the sharded benches
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multiproc.py -m gpu -x -q -k "4" 2>&1 | tail -3
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_ekf20k_8gpu.json 2> gpurun_out/bench_ekf20k_8gpu.err
$TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_ekf20k_4gpu.json 2> gpurun_out/bench_ekf20k_4gpu.err
$TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --landmarks 60000 --steps 10 --warmup 3 > gpurun_out/bench_ekf60k_8gpu.json 2> gpurun_out/bench_ekf60k_8gpu.err
$TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --landmarks 60000 --batch --steps 6 --warmup 3 > gpurun_out/bench_ekf60k_batch_8gpu.json 2> gpurun_out/bench_ekf60k_batch_8gpu.err
if [ "$1" = "pf" ]; then
$TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --workload pf --steps 10 --warmup 3 > gpurun_out/bench_pf_8gpu.json 2> gpurun_out/bench_pf_8gpu.err
$TR --nproc-per-node 4 --master-port 29526 bench.py --gpus 4 --workload pf --steps 10 --warmup 3 > gpurun_out/bench_pf_4gpu.json 2> gpurun_out/bench_pf_4gpu.err
fi
for f in gpurun_out/bench_*gpu.err; do tail -n 1 $f; done
