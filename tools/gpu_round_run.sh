set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v10.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v10.log 2>&1
python bench.py > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_v10.json 2>> gpurun_out/bench_v10.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_v10_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:logmel_fused -c 1 -s 3 -o gpurun_out/r01_v10_logmel -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/pytest_gpu_v10.log; cat gpurun_out/smoke_v10.log | tail -2; cat gpurun_out/bench_v10.json
