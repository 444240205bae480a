for cfg in "3 16 1 1 3 64 8 1 0" "3 12 2 1 2 27 8 0 1 3 2" "3 24 1 1 3 52 24 0 0" "3 16 1 0 3 64 8 0 1" "3 32 1 1 3 64 64 0 0"; do
  echo "=== $cfg"; python tests/run_parity.py $cfg 2>&1 | grep -E "AEs|times|pcg iters|FAIL|PARITY|pattern_mismatch|pattern_only|ASSERT|rror|note" | tr '\n' ' ' ; echo
done
