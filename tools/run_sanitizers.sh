#!/bin/bash
# run_sanitizers.sh OUTDIR — compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over smoke() and the randomised-scene /
# builder / refit / lifecycle GPU tests.  One log per tool under OUTDIR plus a summary line each; the logs of the shipped
# library are committed under profiles/.  racecheck covers the shared-memory stack of k_trace_wide, the block-local fit and the one-kernel builders;
# memcheck covers the speculative row-above-the-top stores and the atomic climb's global traffic.
out=${1:-gpurun_out/sanitizer}
mkdir -p "$out"
export PYTHONUNBUFFERED=1
san=/usr/local/cuda/bin/compute-sanitizer
tests="tests/test_gpu_fuzz.py tests/test_gpu_refit.py tests/test_gpu_lifecycle.py tests/test_watertight.py tests/test_gpu_parity.py"
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  extra=""
  limit=50
  [ "$tool" = racecheck ] && extra="--racecheck-report all" && limit=200000
  log="$out/${tool}_smoke.log"
  timeout 300 $san --tool $tool $extra --error-exitcode 99 --print-limit $limit python __graft_entry__.py smoke > "$log" 2>&1
  echo "$tool smoke rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY' "$log" | tr '\n' ' ')" | tee -a "$out/summary.txt"
  log="$out/${tool}_tests.log"
  # (the largest builder / TLAS stress cases and the 400 k-ray parity tests are left out: minutes each under a sanitizer)
  sel="not two_ranks and not soup3 and not large and not 32768 and not 32769 and not soup12800 and not traversal and not fullsize"
  # racecheck is ~100x slower: a subset — random scenes, refit, watertight, and the one-kernel builders (k_build_small / k_tlas_small:
  # counting sort, run merge, shared-memory fit) on sizes with one and with several blocks
  [ "$tool" = racecheck ] && sel="test_random_scene_parity and (0 or 4 or 8) or refit or watertight or soup257 or soup2049 or duplicates and not large or test_tlas_block_structure_bit_exact and (257 or 1025)"
  timeout 900 $san --tool $tool $extra --error-exitcode 99 --print-limit $limit python -m pytest $tests -m gpu -q -x -k "$sel" > "$log" 2>&1
  echo "$tool tests rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY| passed| failed' "$log" | tr '\n' ' ')" | tee -a "$out/summary.txt"
  if [ "$tool" = racecheck ]; then   # hazards per kernel and source line (the full logs are large: only this digest is kept)
    for f in "$out/racecheck_smoke.log" "$out/racecheck_tests.log"; do
      echo "== $f" >> "$out/racecheck_digest.txt"
      grep -E "^=========     (Write|Read) Thread" "$f" | sed -E 's/Thread \([0-9,]+\) at //; s/\(.*\)\+0x[0-9a-f]+ in / /' | sort | uniq -c | sort -rn | head -40 >> "$out/racecheck_digest.txt"
      grep -E "RACECHECK SUMMARY" "$f" >> "$out/racecheck_digest.txt"
      head -c 200000 "$f" > "$f.head"; mv "$f.head" "$f"
    done
  fi
done
