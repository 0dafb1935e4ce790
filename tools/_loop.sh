for i in 1 2 3 4 5 6; do echo "run $i: $(python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k 'pe_chimeric_repeats_n or test_repeat_heavy' 2>&1 | tail -1)"; done
