O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k "normals or dropin or random" > $O/pytest_r2y.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_r2y.log
timeout 120 python profiles/default_cfg.py 2>&1 | grep -A3 "^default config"
VOXEL=0.1 RADIUS=0.5 timeout 120 python profiles/default_cfg.py 2>&1 | grep -A2 "^default config"
