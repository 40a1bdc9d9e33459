mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config C5 --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err; tail -c 1800 gpurun_out/bench_c5_n2.json; tail -3 gpurun_out/bench_c5_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 | tail -c 600
