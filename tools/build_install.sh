#!/bin/bash
# Build libsvb200.so out of tree and install it atomically (a gpurun snapshot may be taken at any moment), then rebuild oracle/_ref.
set -e
R=/root/repo
cd $R/svfsiplus_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -I$R/include -I. -DSVB200_WITH_METIS \
  api.cu ../host/partition_metis.cpp /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a -o /tmp/libsvb200_new.so -ldl "$@"
cp /tmp/libsvb200_new.so $R/svfsiplus_b200/libsvb200.so.tmp
mv $R/svfsiplus_b200/libsvb200.so.tmp $R/svfsiplus_b200/libsvb200.so
make -s -C $R/oracle ref
ls -la $R/svfsiplus_b200/libsvb200.so $R/oracle/_ref/svmultiphysics_b200
