#!/bin/bash
# compute-sanitizer over the small-configuration GPU tests: memcheck on
# the primitives + one front-end analysis of every type, racecheck on the kernels with
# hand-rolled mbarrier rings / warp-specialised producers (gram_proj, accum_u, the GEMM, the
# small-matrix kernels, the percentile selection).  Usage (on a GPU box):
#   bash scripts/sanitize.sh gpurun_out/sanitize
out=${1:-gpurun_out/sanitize}
mkdir -p "$out"
SEL_MEM='dgemm_matches_numpy or gram_proj_and_accum_u_match_numpy[50-1] or gram_proj_and_accum_u_match_numpy[128-10] or gram_proj_and_accum_u_match_numpy[777-41] or small_decomp_matches_lapack[10] or small_decomp_rank_deficient or percentile_ties or percentile_series or gaussian_tables or pvals_and_boot_ratio or crosscov_behavioral or crosscov_meancentered or index_tables_obey'
compute-sanitizer --tool memcheck --error-exitcode 1 --log-file "$out/memcheck_primitives.log" \
  python -m pytest tests/test_gpu_primitives.py -x -q -m gpu -k "$SEL_MEM" > "$out/memcheck_primitives.pytest" 2>&1
echo "memcheck primitives rc=$?" >> "$out/summary.txt"
compute-sanitizer --tool memcheck --error-exitcode 1 --log-file "$out/memcheck_frontend.log" \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matches_reference_golden and (linnerud or mc0_rot or t5 or 3d_mean or crossval_corr)" > "$out/memcheck_frontend.pytest" 2>&1
echo "memcheck front-end rc=$?" >> "$out/summary.txt"
SEL_RACE='dgemm_matches_numpy[128-128-16] or gram_proj_and_accum_u_match_numpy[128-10] or gram_proj_and_accum_u_match_numpy[777-41] or small_decomp_matches_lapack[10] or small_decomp_matches_lapack[25] or percentile_ties or crosscov_behavioral'
compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 --log-file "$out/racecheck_primitives.log" \
  python -m pytest tests/test_gpu_primitives.py -x -q -m gpu -k "$SEL_RACE" > "$out/racecheck_primitives.pytest" 2>&1
echo "racecheck primitives rc=$?" >> "$out/summary.txt"
compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 --log-file "$out/racecheck_frontend.log" \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matches_reference_golden and (linnerud or t5)" > "$out/racecheck_frontend.pytest" 2>&1
echo "racecheck front-end rc=$?" >> "$out/summary.txt"
for f in "$out"/*.log; do echo "== $f"; tail -3 "$f"; done >> "$out/summary.txt"
for f in "$out"/*.pytest; do echo "== $f"; tail -2 "$f"; done >> "$out/summary.txt"
