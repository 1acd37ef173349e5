timeout 120 python profiles/default_cfg.py 2>&1 | grep -A4 "^default config"
VOXEL=0.1 RADIUS=0.5 timeout 120 python profiles/default_cfg.py 2>&1 | grep -A3 "^default config"
VOXEL=0.05 RADIUS=0.3 timeout 120 python profiles/default_cfg.py 2>&1 | grep -A3 "^default config"
