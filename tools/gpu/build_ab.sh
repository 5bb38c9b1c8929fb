#!/bin/bash
# builds alternative librlerc_<tag>.so files in /tmp/ab and copies them into the package dir (git-ignored *.so, they travel to the GPU box)
# usage: tools/gpu/build_ab.sh tag1 "flags1" tag2 "flags2" ...
set -e
PKG=/root/repo/rle-based-voxel-raycasting_b200
while [ $# -gt 1 ]; do
  tag=$1; flags=$2; shift 2
  ( make -C $PKG/csrc BUILD=/tmp/ab/$tag LIB=$PKG/librlerc_$tag.so EXTRA_NVFLAGS="$flags" $PKG/librlerc_$tag.so > /tmp/ab_$tag.log 2>&1 && grep -h "Used" /tmp/ab/$tag/traverse_filter.ptxas.log | head -1 | sed "s/^/$tag: /" ) &
done
wait
