#!/bin/bash
# Corpus sweep with the reference's CSV contract (eval_csrmv.sh:1-17): one --quiet line per .mtx file.
if (( $# != 2 )); then
  echo "$0 <mtx dir> <cpu_spmv | gpu_spmv [--device=...]>"
  exit 0
fi
HERE="$(dirname "$(readlink -f "$0")")"
echo "file, num_rows, num_cols, num_nonzeros, row_length_mean, row_length_std_dev, row_length_variation, row_length_skewness, method_name, setup_ms, avg_spmv_ms, gflops, effective_GBs"
MTX_DIR=$1
shift
for i in $(find "$MTX_DIR" -name '*.mtx'); do
  "$HERE"/$@ --quiet --mtx="$i"
done
