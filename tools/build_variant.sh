#!/bin/bash
# builds a variant of the library into variants/libp2g_<name>.so: tools/build_variant.sh name "-DFLAG=..."  (select it with P2G_LIBRARY)
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=/tmp/p2g_variant_$name
rm -rf $tmp && mkdir -p $tmp/plonky2_goldibear_b200 $tmp/include
cp -r $root/plonky2_goldibear_b200/csrc $tmp/plonky2_goldibear_b200/csrc
cp -r $root/include/. $tmp/include/
rm -rf $tmp/plonky2_goldibear_b200/csrc/build
make -s -C $tmp/plonky2_goldibear_b200/csrc EXTRA="$flags" > /dev/null
mkdir -p $root/variants
cp $tmp/plonky2_goldibear_b200/libp2g_cuda.so $root/variants/libp2g_$name.so
echo built variants/libp2g_$name.so
