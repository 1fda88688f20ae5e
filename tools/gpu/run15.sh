#!/bin/bash
# statistical tests must not flake: run the BPT comparisons several times and print their statistics
for i in 1 2 3 4 5 6; do python -m pytest tests/test_gpu_renderers.py -q -x -k "bidirectional or script_front" 2>&1 | tail -2; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
