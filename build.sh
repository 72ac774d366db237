#!/bin/bash
# Builds the C-ABI library for sm_100a in-tree (same command as __graft_entry__.build()).
set -e
cd "$(dirname "$0")/hicpeaks_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" -o ../libhicpeaks_b200.so hp_api.cu hp_hostpack.cpp
gcc -O2 -shared -fPIC -I "$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')" hp_pyhelper.c -o ../_hpfast.so
