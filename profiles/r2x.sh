for mm in 2 4 6 10 32; do APC_KNN_MERGE_MIN=$mm timeout 120 python profiles/knn_ab.py 2>&1 | tail -1 | sed "s/^/merge_min=$mm /"; done
for h in 1.25 1.5 2.0; do APC_KNN_HINT=$h timeout 120 python profiles/knn_ab.py 2>&1 | tail -1 | sed "s/^/hint=$h /"; done
