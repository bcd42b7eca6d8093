set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_full.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
cp gpurun_out/kernel_detail_full.json gpurun_out/r02_kernel_detail_full.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --workload duf --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_duf.json 2> gpurun_out/r02_bench_duf.err
python tools/pair_ab.py --iters 15 --json gpurun_out/r02_pair_ab.json > gpurun_out/r02_pair_ab.log 2>&1
python tools/bw_bench.py --iters 10 --json gpurun_out/r02_bw_kernels.json > gpurun_out/r02_bw_kernels.log 2>&1
python bench.py --no-graph --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 800 -c 1200 --csv --log-file gpurun_out/r02_launches_time_dram.csv python bench.py --no-graph --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/r02_gpu_tests_full.log
