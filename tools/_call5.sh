set -x
mkdir -p gpurun_out/r02e
python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r02e/pytest_all.log 2>&1
tail -25 gpurun_out/r02e/pytest_all.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e/bench_n1.json 2> gpurun_out/r02e/bench_n1.err
tail -3 gpurun_out/r02e/bench_n1.err
CUGS_B200_LIB=$PWD/cuda_gaussian_splatting_b200/libcugs_b200_tma.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e/bench_n1_tma.json 2> gpurun_out/r02e/bench_n1_tma.err
tail -3 gpurun_out/r02e/bench_n1_tma.err
python bench.py --steps 20 --warmup 3 --workload A --no-cpu-baseline > gpurun_out/r02e/bench_A.json 2> gpurun_out/r02e/bench_A.err
python bench.py --mode train_step --workload C --steps 20 --warmup 3 > gpurun_out/r02e/bench_train_C.json 2> gpurun_out/r02e/bench_train_C.err
python tools/sort_bench.py B > gpurun_out/r02e/sort_bench_B.jsonl 2> gpurun_out/r02e/sort_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02e/bench_n2.json 2> gpurun_out/r02e/bench_n2.err
tail -3 gpurun_out/r02e/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/exchange_probe.py > gpurun_out/r02e/exchange_probe_n2.json 2> gpurun_out/r02e/exchange_probe_n2.err
tail -3 gpurun_out/r02e/exchange_probe_n2.err
