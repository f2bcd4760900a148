# usage: build_variant.sh name "<nvcc extra flags>"   -> profiles/ab/libs/name.so (the product .so is left alone)
set -e
name=$1; shift
mkdir -p profiles/ab/libs
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart static --extended-lambda $@ \
  -o profiles/ab/libs/$name.so jpeg-encoder-and-decoder_b200/csrc/jpgb200.cu
echo built profiles/ab/libs/$name.so
