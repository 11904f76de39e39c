# usage: bash tests/ubench/ncu_ops.sh <tag> <kernel-regex> <perf_vs_ref --only regex> [launch-count]
# Runs the selected ops once without ncu (must exit 0), then captures --set full for the matching kernels.
set -x
TAG=$1; KRE=$2; ONLY=$3; CNT=${4:-6}
CMD="python tests/perf_vs_ref.py --B 32 --no-ref --iters 2 --only $ONLY --out gpurun_out/${TAG}_perf.json"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KRE -c $CNT -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/${TAG}_plain.log
