O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k "dropin or fullsize" > $O/pytest_r2n2.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_r2n2.log
timeout 200 python profiles/node_times.py 2>&1 | grep -v Warning | tail -8
