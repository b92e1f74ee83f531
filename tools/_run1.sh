timeout 200 python -m pytest tests -m gpu -x -q -k "staged or mp_only or large_and_small or (other_datasets and gin)" 2>&1 | tail -15
timeout 120 python tools/hep_probe.py
