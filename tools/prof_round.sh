# ncu evidence of a round (run under gpurun, one GPU): launch list of the bench command, then one full capture each of
# the pose kernel (device-resident variant), the edge passes and the Fetch kernels.  R = file prefix (e.g. r2).
R=${1:-r2}
set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${R}_prof_plain.json 2> gpurun_out/${R}_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench_alpha15.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${R}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:checkStatesKernel<unsigned int, \(bool\)0, \(bool\)0, \(bool\)0>' -s 2 -c 1 -o gpurun_out/${R}_states_alpha15 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${R}_ncu_states.log 2>&1
python tools/edge_prof.py cubicles 65536 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:checkEdgesKernel -s 6 -c 3 -o gpurun_out/${R}_edges_cubicles -f python tools/edge_prof.py cubicles 65536 > gpurun_out/${R}_ncu_edges.log 2>&1
python tools/fetch_prof.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fetch -s 12 -c 6 -o gpurun_out/${R}_fetch_states -f python tools/fetch_prof.py > gpurun_out/${R}_ncu_fetch.log 2>&1
