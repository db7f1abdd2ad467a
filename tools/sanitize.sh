#!/bin/bash
# compute-sanitizer pass over the hand-rolled mbarrier / tensor-memory / cluster-multicast pipelines.
# Run on the GPU box:   bash tools/sanitize.sh            (output: gpurun_out/sanitizer/*.txt, summary on stdout)
# memcheck + racecheck + synccheck + initcheck over smoke() and the small MLP / scoring / NGP / surrogate tests,
# the MLP kernel additionally with every cluster launch mode (EVPP_TC_CLUSTER = 1, 2, 4, 42).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/sanitizer
mkdir -p "$OUT"
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
PER=${PER:-420}            # seconds per run
SMOKE='import __graft_entry__ as g; g.smoke()'
TESTS_MLP='mlp_tensor_core or split_precision or score_only or mlp_fp32_known'
TESTS_REST='edge_cases or zero_density or surrogate_fit or ngp_march_counts or tsdf_render_rays or device_pixel_sampler or select'

run() {   # name tool extra-env command...
  local name=$1 tool=$2; shift 2
  local log="$OUT/${name}_${tool}.txt"
  echo "== $name / $tool" | tee "$log"
  timeout "$PER" "$CS" --tool "$tool" --target-processes all --error-exitcode 86 --print-limit 20 \
      $( [ "$tool" = memcheck ] && echo "--leak-check no" ) "$@" >> "$log" 2>&1
  local rc=$?
  echo "exit code $rc" >> "$log"
  local errs
  errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" "$log" | tr '\n' ';')
  echo "   rc=$rc  $errs"
}

for tool in memcheck racecheck synccheck initcheck; do
  run smoke "$tool" python -c "$SMOKE"
done
for c in 1 2 4 42; do
  for tool in memcheck synccheck; do
    EVPP_TC_CLUSTER=$c run "mlp_cluster$c" "$tool" python -m pytest tests -m gpu -x -q -k "$TESTS_MLP"
  done
done
EVPP_TC_CLUSTER=2 run mlp_cluster2 racecheck python -m pytest tests -m gpu -x -q -k "mlp_tensor_core or split_precision"
for tool in memcheck racecheck synccheck; do
  run rest "$tool" python -m pytest tests -m gpu -x -q -k "$TESTS_REST"
done
echo "== summary"
grep -H -E "ERROR SUMMARY|RACECHECK SUMMARY|exit code" "$OUT"/*.txt
