#!/bin/bash
# compute-sanitizer passes over small parity cases (every kernel family of the hot path, the fused consumers, the
# spin-diagonal SpMV): memcheck on the wider subset, racecheck + synccheck on the TMA/mbarrier pipelines.
# usage (GPU box): bash tools/sanitize.sh [outdir]      -> <outdir>/sanitize_{memcheck,racecheck,synccheck}.log
out=${1:-gpurun_out}
mkdir -p "$out"
CS=/usr/local/cuda/bin/compute-sanitizer
MEM_K='tiny or (test_recur_b and bulk_hoh) or (test_recur_b and impurity) or test_recur_b_ij or test_recur_scalar or test_chebyshev_recur or test_compute_moments_stochastic or test_spin_diagonal or test_edge_sizes'
RACE_K='(test_recur_b and tiny) or (test_recur_b and impurity_hoh) or (test_chebyshev_recur and not ij) or test_spin_diagonal_hoppings'
timeout 240 $CS --tool memcheck --error-exitcode 77 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$MEM_K" > "$out/sanitize_memcheck.log" 2>&1
echo "memcheck rc=$?" >> "$out/sanitize_memcheck.log"
timeout 150 $CS --tool memcheck --error-exitcode 77 --print-limit 20 \
  python -m pytest tests/test_gpu_post.py -x -q -m gpu -k "fused or intersite" > "$out/sanitize_memcheck_post.log" 2>&1
echo "memcheck rc=$?" >> "$out/sanitize_memcheck_post.log"
timeout 240 $CS --tool racecheck --error-exitcode 77 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$RACE_K" > "$out/sanitize_racecheck.log" 2>&1
echo "racecheck rc=$?" >> "$out/sanitize_racecheck.log"
timeout 150 $CS --tool synccheck --error-exitcode 77 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$RACE_K" > "$out/sanitize_synccheck.log" 2>&1
echo "synccheck rc=$?" >> "$out/sanitize_synccheck.log"
tail -n 4 "$out"/sanitize_*.log
