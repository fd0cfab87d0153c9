for v in "2 254" "4 96" "6 64" "8 48" "8 32"; do set -- $v; echo "=== NS=$1 CH=$2"; B2_NS=$1 B2_CH=$2 timeout 200 python tools/opprof.py C4 | head -7; done
