python -m pytest tests/test_gpu_renderers.py -x -q -m gpu 2>&1 | tail -3
CONFIGS="C2 C3 C4 C5" ITERS=32 REPS=3 bash tools/gpu/ab_configs.sh base nohint
