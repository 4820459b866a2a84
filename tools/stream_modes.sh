# e2e of 2^20-pose mplb_check_states calls under the large-batch knobs (diagnostic)
for S in alpha15 twistycool apartment; do
echo "== $S default"; MPLB_TIMELINE=1 python tools/stream_sweep.py $S 2>&1 | tail -3
echo "== $S flag as 4-byte copy"; MPLB_STREAM_FLAGCOPY=1 MPLB_TIMELINE=1 python tools/stream_sweep.py $S 2>&1 | tail -3
echo "== $S throughput variant"; MPLB_STREAM_THROUGHPUT=1 MPLB_TIMELINE=1 python tools/stream_sweep.py $S 2>&1 | tail -3
done
for C in 65536 262144; do
echo "== alpha15 chunk=$C"; MPLB_STREAM_CHUNK=$C python tools/stream_sweep.py alpha15 2>&1 | tail -1
echo "== alpha15 chunk=$C flagcopy"; MPLB_STREAM_FLAGCOPY=1 MPLB_STREAM_CHUNK=$C python tools/stream_sweep.py alpha15 2>&1 | tail -1
done
echo "== alpha15 flat 131072"; MPLB_STREAM_FLAT=1 python tools/stream_sweep.py alpha15 2>&1 | tail -1
echo "== alpha15 pipelined"; MPLB_STREAM_MODE=0 python tools/stream_sweep.py alpha15 2>&1 | tail -1
echo "== alpha15 zerocopy"; MPLB_STREAM_MODE=2 python tools/stream_sweep.py alpha15 2>&1 | tail -1
