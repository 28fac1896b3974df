#!/usr/bin/env bash
# compute-sanitizer tier for the hand-written sm_100a kernels (the reference has no sanitizer hooks; in-kernel
# peer-memory signalling and mbarrier / TMEM pipelines are the new race surface here).
#
#   ci_scripts/sanitize_kernels.sh                 # memcheck + racecheck + synccheck + initcheck on one GPU
#   TOOLS="memcheck" ci_scripts/sanitize_kernels.sh
#   NGPU=2 ci_scripts/sanitize_kernels.sh          # additionally: peer-memory kernels under memcheck with poisoned slabs
#
# Every tool runs `tools/kernel_check.py small` (each kernel once, tiny shapes).  B200_SYMM_DEBUG=1 makes the symmetric-heap
# back-ends poison their slabs with NaN before every exchange and assert that everything consumed was written.
set -uo pipefail
cd "$(dirname "$0")/.."
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
TOOLS=${TOOLS:-"memcheck racecheck synccheck initcheck"}
OUT=${OUT:-gpurun_out/sanitizer}
mkdir -p "$OUT"
rc=0
for tool in $TOOLS; do
    echo "== compute-sanitizer --tool $tool"
    timeout "${TMO:-600}" "$SAN" --tool "$tool" --error-exitcode 9 --print-limit 20 --log-file "$OUT/$tool.log" \
        python tools/kernel_check.py small > "$OUT/$tool.stdout" 2>&1
    r=$?
    tail -3 "$OUT/$tool.log" 2>/dev/null
    grep -E "SMALL_ALL_OK|SMALL_HAS_FAILURES|Error" "$OUT/$tool.stdout" | tail -3
    echo "== $tool exit code $r"
    [ $r -ne 0 ] && rc=$r
done
if [ "${NGPU:-1}" -gt 1 ]; then
    echo "== peer-memory kernels, poisoned slabs (B200_SYMM_DEBUG=1), $NGPU GPUs"
    B200_SYMM_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 \
        --master-port 29577 tools/fused_comm_check.py > "$OUT/symm_debug.stdout" 2>&1
    r=$?
    tail -3 "$OUT/symm_debug.stdout"
    [ $r -ne 0 ] && rc=$r
fi
exit $rc
