#!/bin/bash
# A/B builds of libvitk.so that differ only in the compile-time knobs of one source file (SRC=gemm2 by default,
# e.g. SRC=attention): tools/build_variants.sh name "-DX=1 -DY=2" ...
# → chest-x-ray-vit_b200/csrc/build/variants/libvitk_<name>.so   (select with VITK_LIB=<path>)
set -e
cd "$(dirname "$0")/../chest-x-ray-vit_b200/csrc"
make -s > /dev/null
mkdir -p build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
SRC=${SRC:-gemm2}
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  /usr/local/cuda/bin/nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $defs -c $SRC.cu -o build/variants/${SRC}_$name.o
  objs=$(ls build/*.o | grep -v "build/$SRC.o")
  /usr/local/cuda/bin/nvcc $ARCH -shared -cudart static -o build/variants/libvitk_$name.so $objs build/variants/${SRC}_$name.o
  echo "built build/variants/libvitk_$name.so ($defs)"
done
