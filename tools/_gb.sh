echo "== racecheck: smoke + fused class side (K=2, Vc=301) + wide codes" > gpurun_out/sanitizer_race.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/sanitizer_race.log 2>&1
echo "racecheck smoke rc=$?" >> gpurun_out/sanitizer_race.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "2-301-256 or wide_codes" >> gpurun_out/sanitizer_race.log 2>&1
echo "racecheck tests rc=$?" >> gpurun_out/sanitizer_race.log
grep -E "RACECHECK SUMMARY|rc=|passed|failed|smoke|Race reported" gpurun_out/sanitizer_race.log | head -20
