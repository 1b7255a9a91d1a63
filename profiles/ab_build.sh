#!/bin/bash
# A/B builds of the read kernel: profiles/ab_build.sh <name> <extra nvcc flags...> -> build_ab/<name>/libvgan_{gpu,host}.so
# (run a variant with VGAN_B200_LIB_DIR=build_ab/<name> python bench.py ...; build_ab/ is git-ignored but travels with gpurun)
set -e
cd "$(dirname "$0")/../vgan_b200/csrc"
name=$1; shift
out=../../build_ab/$name
mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC "$@" -c -o $out/hc_kernels.o hc_kernels.cu
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC "$@" -c -o $out/vgl_api.o vgl_api.cu   # (shares the sizing macros)
/usr/local/cuda/bin/nvcc $ARCH -shared -o $out/libvgan_gpu.so $out/vgl_api.o $out/hc_kernels.o euka_kernels.o recon_kernels.o gam_kernels.o -cudart static
g++ -O3 -std=c++17 -fPIC -fopenmp -shared -o $out/libvgan_host.so host/vgan_host.cpp host/synth.cpp host/euka_host.cpp -L$out -lvgan_gpu -lz -Wl,-rpath,'$ORIGIN'
rm -f $out/hc_kernels.o $out/vgl_api.o
echo built $out
