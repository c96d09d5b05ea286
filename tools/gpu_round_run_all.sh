bash tools/gpu_round_run.sh 2>&1 | grep -v "^+" | tail -8
bash tools/gpu_whisper_run.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_whisper_launches.csv python tools/whisper_bench_step.py > /dev/null 2>&1
