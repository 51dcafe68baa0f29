#!/bin/bash
for x in "" _x1 _x2 _x3; do
  echo "lib=$x"
  BG_LIB=$PWD/brepgen_b200/libbrepgen_b200$x.so timeout 200 python tools/attn_check.py 2>&1 | tail -2
done
