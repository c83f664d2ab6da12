for F in 8; do echo "F=$F"; timeout 120 python tools/tailprof.py $F 2>&1 | tail -12; done
