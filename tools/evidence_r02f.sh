set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02f_gpu_tests_full.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_full.json 2> gpurun_out/r02f_bench_full.err
cp gpurun_out/kernel_detail_full.json gpurun_out/r02f_kernel_detail_full.json
python bench.py --workload duf --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_duf.json 2> gpurun_out/r02f_bench_duf.err
python __graft_entry__.py smoke > gpurun_out/r02f_smoke.log 2>&1
cuobjdump -sass vsr_b200/lib/libvsr_sm100.so | grep -oE "UTCHMMA[.A-Z0-9]*|UTMALDG[.A-Z0-9]*|UTMASTG[.A-Z0-9]*|LDTM[.xA-Z0-9]*|UTCBAR[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|HMMA[.A-Z0-9]*|FADD2|FMUL2" | sort | uniq -c > gpurun_out/r02f_sass_mnemonics.txt
tail -3 gpurun_out/r02f_gpu_tests_full.log; tail -1 gpurun_out/r02f_smoke.log
