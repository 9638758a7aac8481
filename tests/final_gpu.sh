mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "packed or pack_merge or varlen" 2>&1 | tail -25 > gpurun_out/final_tests2.log
tail -25 gpurun_out/final_tests2.log
timeout 220 python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/final_bench.json; tail -5 gpurun_out/final_bench.err
timeout 100 python bench.py --steps 3 --warmup 3 --pack --skip-e2e --no-cpu-baseline > gpurun_out/final_bench_pack.json 2> gpurun_out/final_bench_pack.err; echo "pack rc=$?"; tail -c 1500 gpurun_out/final_bench_pack.json; tail -5 gpurun_out/final_bench_pack.err
