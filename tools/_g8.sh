python -m pytest tests -m gpu -q -k "multi_gpu" 2>&1 | tail -2 > gpurun_out/r04h_multigpu_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r04h_bench_8gpu.json 2> gpurun_out/r04h_b8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --shard train --steps 10 --warmup 3 > gpurun_out/r04h_bench_train_sharded_8gpu_N4e6.json 2> gpurun_out/r04h_b8t.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r04h_bench_2gpu.json 2> gpurun_out/r04h_b2.err
cat gpurun_out/r04h_multigpu_pytest.txt; head -c 250 gpurun_out/r04h_bench_8gpu.json; echo; head -c 250 gpurun_out/r04h_bench_train_sharded_8gpu_N4e6.json; echo; head -c 250 gpurun_out/r04h_bench_2gpu.json
