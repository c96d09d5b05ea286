set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python tools/bench_whisper.py > gpurun_out/bench_whisper.json 2> gpurun_out/bench_whisper.err
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_whisper.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['whisper_preset'])"
