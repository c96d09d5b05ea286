python tools/bench_configs.py --config 1 2>/dev/null | tail -1 > gpurun_out/cfg1.json
python tools/bench_configs.py --config 3 2>/dev/null | tail -1 > gpurun_out/cfg3.json
python tools/bench_configs.py --config 4 2>/dev/null | tail -1 > gpurun_out/cfg4.json
python tools/bench_configs.py --config 4 --batch 256 2>/dev/null | tail -1 > gpurun_out/cfg4_256.json
python tools/bench_configs.py --config 5 2>/dev/null > gpurun_out/cfg5.jsonl
cat gpurun_out/cfg1.json gpurun_out/cfg3.json gpurun_out/cfg4.json gpurun_out/cfg4_256.json | cut -c1-420; wc -l gpurun_out/cfg5.jsonl
