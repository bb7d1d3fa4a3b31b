#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG=.. -DFLAG2=.."  -> tools/variants/NAME/libskb.so (experiments; not tracked)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
V=$ROOT/tools/variants/$1
mkdir -p $V
cp $ROOT/skysampler_b200/csrc/*.cu $ROOT/skysampler_b200/csrc/*.h $ROOT/skysampler_b200/csrc/*.cuh $V/
sed -i "s#\"../../include/skb.h\"#\"$ROOT/include/skb.h\"#" $V/skb_internal.h
cd $V
for f in skb_api skb_score skb_sparse skb_sample skb_fit; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $2 -c -o $f.o $f.cu 2> $f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libskb.so skb_api.o skb_score.o skb_sparse.o skb_sample.o skb_fit.o
rm -f *.o
ls -la $V/libskb.so
