#!/bin/sh
# usage: build_rig.sh <out> [sanitizer flags...]   — TEST INFRASTRUCTURE: engine host logic + fake CUDA, no GPU
set -e
OUT=$1; shift
HERE=$(dirname "$0")
CS=$HERE/../../demodel_b200/csrc
g++ -std=c++17 -O1 -g "$@" -I "$HERE/fake_cuda" -pthread -o "$OUT" \
    -x c++ "$CS/engine_core.cu" "$CS/engine_api.cu" "$CS/engine_cache.cu" "$CS/proxy_driver.cc" "$CS/manifest.cc" "$CS/gunzip.cc" "$HERE/fake_cuda.cc" "$HERE/engine_soak.cc"
