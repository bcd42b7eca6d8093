set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02b_gpu_tests_full.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_full.json 2> gpurun_out/r02b_bench_full.err
cp gpurun_out/kernel_detail_full.json gpurun_out/r02b_kernel_detail_full.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err
python bench.py --precision bf16x3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02b_bench_bf16x3.json 2> gpurun_out/r02b_bench_bf16x3.err
cp gpurun_out/kernel_detail_full.json gpurun_out/r02b_kernel_detail_bf16x3.json
python bench.py --precision bf16x3 --no-graph --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/plain_x3.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 4300 -c 1500 --csv --log-file gpurun_out/r02b_launches_bf16x3_time_dram.csv python bench.py --precision bf16x3 --no-graph --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench_x3.log 2>&1
cuobjdump -sass vsr_b200/lib/libvsr_sm100.so | grep -oE "UTCHMMA[.A-Z0-9]*|UTMALDG[.A-Z0-9]*|UTMASTG[.A-Z0-9]*|LDTM[.xA-Z0-9]*|UTCBAR[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|HMMA[.A-Z0-9]*" | sort | uniq -c > gpurun_out/r02b_sass_mnemonics.txt
tail -3 gpurun_out/r02b_gpu_tests_full.log
