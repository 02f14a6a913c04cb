O=gpurun_out/r01n; mkdir -p $O
SAN_CASES=0,6,7 timeout 420 compute-sanitizer --tool racecheck python tests/sanitizer_check.py > $O/sanitizer_racecheck.log 2>&1; grep -E "OK|MISMATCH|ERROR SUMMARY|RACECHECK SUMMARY|Race reported" $O/sanitizer_racecheck.log | head -20
timeout 300 compute-sanitizer --tool synccheck python tests/sanitizer_check.py > $O/sanitizer_synccheck.log 2>&1; grep -E "OK|MISMATCH|ERROR SUMMARY" $O/sanitizer_synccheck.log | tail -4
