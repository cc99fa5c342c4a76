#!/bin/bash
# tools/ab_run.sh "<command>" name1 name2 ...: runs <command> once per ab/lib_<name>.so on the SAME box (boxes differ by several percent)
cd "$(dirname "$0")/.."; mkdir -p ab
cmd=$1; shift
cp vibevoice_b200/csrc/libvibevoice_b200.so ab/lib__orig.so
for rep in 1 2; do
  for n in "$@"; do
    cp ab/lib_$n.so vibevoice_b200/csrc/libvibevoice_b200.so
    echo "=== $n (pass $rep)"
    bash -c "$cmd"
  done
done
cp ab/lib__orig.so vibevoice_b200/csrc/libvibevoice_b200.so
