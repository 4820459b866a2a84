for S in alpha15 twistycool apartment; do
MPLB_STREAM_MODE=2 python tools/stream_sweep.py $S 2>&1 | grep chunk | sed 's/^/zerocopy /'
MPLB_STREAM_MODE=1 MPLB_STREAM_CHUNK=131072 python tools/stream_sweep.py $S 2>&1 | grep chunk | sed 's/^/flags /'
MPLB_STREAM_MODE=0 python tools/stream_sweep.py $S 2>&1 | grep chunk | sed 's/^/pipelined /'
done
