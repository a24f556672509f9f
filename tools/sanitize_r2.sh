# tools/sanitize_r2.sh: compute-sanitizer over representative parity cases of every kernel (SE / PE x WGBS / RRBS / wide, index build,
# packed input, methratio incl. -r); logs -> gpurun_out/r2_sanitizer_*.log (copied to profiles/)
export PYTHONUNBUFFERED=1
SEL='test_se_matches_oracle_and_reference or test_pe_matches_oracle_and_reference or test_index_matches_oracle or test_empty_and_ragged or test_batching_is_invisible'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck parity rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_methratio_gpu.py -x -q -m gpu -k "api_counters or (in_process and (rrbs_se_A or pe_sam.r or se_cfg5))" >> gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck methratio rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(test_se_matches_oracle_and_reference and (se_cfg1 or se_n1 or se_cfg5 or rrbs_se_A)) or (test_pe_matches_oracle_and_reference and (pe_sam or rrbs_pe))" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r2_sanitizer_racecheck.log
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/r2_sanitizer_memcheck.log gpurun_out/r2_sanitizer_racecheck.log | tail -12
