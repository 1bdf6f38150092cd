#!/bin/bash
# Freeze the algorithmic flop counts into profiles/roofline.json (run from anywhere).
set -euo pipefail
cd "$(dirname "$0")/.."
g++ -O1 -std=c++17 -I/usr/local/cuda/include -x c++ tools/count_flops.cpp -o /tmp/nwb_count_flops
/tmp/nwb_count_flops > profiles/roofline.json
cat profiles/roofline.json
