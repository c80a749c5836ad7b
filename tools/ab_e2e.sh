#!/bin/bash
# e2e A/B of tuning builds: tools/ab_e2e.sh lib1.so lib2.so ...  ("default" = the in-tree library); prints tools/prof_e2e.py lines
for lib in "$@"; do
  if [ "$lib" = default ]; then unset B2H_LIB; else export B2H_LIB=$PWD/$lib; fi
  echo "== $lib"; python tools/prof_e2e.py 4096 2>&1 | grep -v pageable | tail -4
done
