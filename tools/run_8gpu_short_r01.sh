set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_ekf20k_8gpu.json 2> gpurun_out/bench_ekf20k_8gpu.err
$TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_ekf20k_4gpu.json 2> gpurun_out/bench_ekf20k_4gpu.err
$TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --landmarks 60000 --steps 10 --warmup 3 > gpurun_out/bench_ekf60k_8gpu.json 2> gpurun_out/bench_ekf60k_8gpu.err
$TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --workload pf --steps 10 --warmup 3 > gpurun_out/bench_pf_8gpu.json 2> gpurun_out/bench_pf_8gpu.err
