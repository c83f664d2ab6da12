#!/bin/bash
for t in 16 8 4; do
  echo "== B200_TILE1=$t B200_TILE2=$t"
  B200_TILE1=$t B200_TILE2=$t timeout 200 python tools/fwdprobe.py 64 "generic"
done
echo "== TILE1=4 TILE2=8"; B200_TILE1=4 B200_TILE2=8 timeout 200 python tools/fwdprobe.py 64 "generic"
echo "== TILE1=8 TILE2=4"; B200_TILE1=8 B200_TILE2=4 timeout 200 python tools/fwdprobe.py 64 "generic"
