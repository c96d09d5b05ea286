set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python tools/bench_small_kernels.py > gpurun_out/small_kernels.jsonl 2> gpurun_out/small_kernels.err
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/small_kernels.jsonl; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e'], d['sustained'], d['stats_pass'], d['config1'], d['config4'], d['cpu_baseline'], d['whisper_preset'])"
