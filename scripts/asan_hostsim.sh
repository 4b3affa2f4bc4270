#!/bin/bash
# AddressSanitizer pass over the host engine (pb_engine.cpp) through the host simulator: rebuilds tests/hostsim/libpb_hostsim.so
# with -fsanitize=address, runs the host-simulator tests under it, then restores the normal build.  CPU only.
# (This is how the planner's use-after-free in the GroupNorm op was confirmed fixed at the end of round 1.)
set -e
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so)
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -fopenmp -std=c++17 -shared -fPIC -fvisibility=hidden \
    -I diffusion_pullback_b200/csrc -I include -o tests/hostsim/libpb_hostsim.so \
    diffusion_pullback_b200/csrc/pb_engine.cpp tests/hostsim/pbk_hostsim.cpp
rc=0
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 python -m pytest tests/test_engine_hostsim.py -x -q -p no:cacheprovider || rc=$?
python -c "from tests.hostsim.build import build; build(force=True)"
exit $rc
