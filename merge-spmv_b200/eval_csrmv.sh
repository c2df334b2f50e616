#!/usr/bin/env bash
# Corpus sweep: one --quiet CSV line per Matrix-Market file under a directory, preceded by the
# column header downstream plotting expects (same columns as the reference's eval_csrmv.sh:8).
#   usage: eval_csrmv.sh <directory with .mtx files> <cpu_spmv | gpu_spmv> [driver flags ...]
set -u
if [ "$#" -lt 2 ]; then
    printf 'usage: %s <mtx dir> <cpu_spmv | gpu_spmv [--device=...]>\n' "$0"
    exit 0
fi
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
corpus="$1"
driver="$2"
shift 2
columns=(file num_rows num_cols num_nonzeros row_length_mean row_length_std_dev row_length_variation
         row_length_skewness method_name setup_ms avg_spmv_ms gflops effective_GBs)
(IFS=,; printf '%s\n' "${columns[*]}" | sed 's/,/, /g')
find "$corpus" -type f -name '*.mtx' -print0 | sort -z | while IFS= read -r -d '' matrix; do
    "$here/$driver" "$@" --quiet --mtx="$matrix"
done
