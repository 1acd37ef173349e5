O=gpurun_out; mkdir -p $O
for az in 2048 8192 32768; do timeout 120 python profiles/stage_times.py 4 $az > $O/stage_times_r2n_$az.log 2>&1; echo "stage_times $az rc=$?"; cat $O/stage_times_r2n_$az.log; done
timeout 120 python profiles/run_pipeline.py 3 > $O/plain_r2n.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_ --launch-skip 28 -c 14 -f -o $O/prof_r2n \
    python profiles/run_pipeline.py 3 > $O/ncu_full_r2n.log 2>&1
echo "full capture rc=$?"
