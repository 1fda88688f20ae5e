timeout 900 python -m pytest tests -x -q -m gpu --timeout 200 2>&1 | tail -6
CONFIGS="C4 C3" ITERS=24 REPS=2 timeout 600 bash tools/gpu/ab_configs.sh base nocompact
