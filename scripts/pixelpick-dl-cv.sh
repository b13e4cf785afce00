#!/bin/bash
# drop-in for the reference scripts/pixelpick-dl-cv.sh (python3 ../main_al.py --dataset_name cv --n_pixels_by_us 10 -qs margin_sampling)
cd "$(dirname "$0")/.." && python3 -m pixelpick_b200.main_al --dataset_name cv --n_pixels_by_us 10 -qs margin_sampling "$@"
