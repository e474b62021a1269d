#!/bin/bash
# Counterpart of the reference's test/sparsity.sh: CSR then COO (and ELL) over sp_{d}_2048x2048 dirs.
# The dirs are generated (seeded) with scripts/gen_data.py if absent -- the reference's gen_sparse.py is unseeded.
# usage: test/sparsity.sh [DATA_ROOT]
HERE=$(cd "$(dirname "$0")/.." && pwd)
CLI=$HERE/cuda-optimization-for-spmm_b200/host/cuspmm
ROOT=${1:-/tmp/cuspmm_sparsity}
OUT=sparsity.json
echo "[" > $OUT
for sp in 0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9; do
  d=$ROOT/sp_${sp}_2048x2048
  [ -f "$d/dense.in" ] || python3 "$HERE/scripts/gen_data.py" "$d" --rows 2048 --cols 2048 --density $sp --N 1024 >&2
done
for fmt in csr coo ell; do
  for sp in 0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9; do
    "$CLI" --$fmt -d "$ROOT/sp_${sp}_2048x2048" --skip-cpu "${@:2}" >> $OUT
  done
done
sed -i '$ s/},$/}/' $OUT
echo "]" >> $OUT
python3 -c "import json; r=json.load(open('$OUT')); print(len(r), 'records')"
