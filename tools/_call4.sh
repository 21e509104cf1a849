set -x
mkdir -p gpurun_out/r02d
python -m pytest tests/test_native_trainer_gpu.py -x -q -m gpu --durations=5 > gpurun_out/r02d/pytest_native.log 2>&1
tail -30 gpurun_out/r02d/pytest_native.log
python -m pytest tests/test_multi_gpu_exchange.py -x -q -m gpu > gpurun_out/r02d/pytest_multi.log 2>&1
tail -30 gpurun_out/r02d/pytest_multi.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02d/bench_n1.json 2> gpurun_out/r02d/bench_n1.err
tail -5 gpurun_out/r02d/bench_n1.err
python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r02d/bench_n1_nograph.json 2> gpurun_out/r02d/bench_n1_nograph.err
python bench.py --steps 20 --warmup 3 --workload A > gpurun_out/r02d/bench_A.json 2> gpurun_out/r02d/bench_A.err
tail -5 gpurun_out/r02d/bench_A.err
python bench.py --steps 20 --warmup 3 --workload A --no-graph --no-cpu-baseline > gpurun_out/r02d/bench_A_nograph.json 2> gpurun_out/r02d/bench_A_nograph.err
python bench.py --mode train_step --workload C --steps 20 --warmup 3 > gpurun_out/r02d/bench_train_C.json 2> gpurun_out/r02d/bench_train_C.err
tail -5 gpurun_out/r02d/bench_train_C.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02d/bench_n2.json 2> gpurun_out/r02d/bench_n2.err
tail -5 gpurun_out/r02d/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --mode train_step --workload C > gpurun_out/r02d/bench_train_C_n2.json 2> gpurun_out/r02d/bench_train_C_n2.err
tail -5 gpurun_out/r02d/bench_train_C_n2.err
