# Round-end record on the GPU box: tests, both bench arms, launch list, one full ncu capture, stage stamps (developer tool)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_tests.txt
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hmpc_solve_kernel -s 9 -c 1 -f -o gpurun_out/prof_final \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
HMPC_CFG=2 HMPC_B=1024 python tests/tools/gpu_check.py > gpurun_out/final_gpu_check.log 2>&1
cat gpurun_out/final_tests.txt; tail -c 600 gpurun_out/final_bench.json; ls -la gpurun_out | tail -12
