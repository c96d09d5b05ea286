# ncu evidence of the headline kernel (run on the GPU box): launch list of a short bench run, then one --set full capture
CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --stats-clips 0 --sustained-seconds 0 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:logmel_fused -s 3 -c 2 -o gpurun_out/r02_logmel $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_launches.log gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
