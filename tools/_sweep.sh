for S in 32 16 8 4; do echo "== S=$S"; XG_INFLATE_S=$S timeout 300 python tools/prof_decode.py 2e7 2>&1 | grep "device want_seq=1" ; done
XG_DECODE_TIMING=1 timeout 300 python tools/prof_decode.py 2e7 2>&1 | grep -B30 "device want_seq=1" | tail -40
