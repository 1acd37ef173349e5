O=gpurun_out; mkdir -p $O
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_knn_query --launch-skip 1 -c 1 -f -o $O/prof_knn_r2x python profiles/run_c4.py > $O/ncu_knn_r2x.log 2>&1; echo "rc=$?"
