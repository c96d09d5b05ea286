ncu --set full --clock-control none --import-source on -k regex:dftgemm_logmel -c 1 -s 2 -o gpurun_out/r01_w7_dftgemm -f python tools/whisper_bench_step.py > gpurun_out/ncu_w7.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dftgemm --csv python tools/whisper_bench_step.py 2>&1 | grep dftgemm | cut -d, -f5,15 | tail -2
