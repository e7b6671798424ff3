python -m pytest tests/test_gpu_voxelize.py -m gpu -x -q 2>&1 | tail -1
for n in 1024 1280 1664 2048 4096; do tools/ab.sh python tools/fill_time.py $n dragon.obj 20 | grep -o "^---.*\|fill mean [0-9.]*" | tr '\n' ' '; echo " N=$n"; done
tools/timeline.sh 2048 | grep -v inside | tail -4
