#!/bin/bash
# Counterpart of the reference's test/ell.sh: run the CLI over the named data dirs and collect the records.
# Unlike the reference's (ell.json is a list of "{...}," fragments), the output is wrapped into valid JSON.
# usage: test/ell.sh [DATA_ROOT] [extra cuspmm flags...]      (DATA_ROOT defaults to tests/golden)
HERE=$(cd "$(dirname "$0")/.." && pwd)
CLI=$HERE/cuda-optimization-for-spmm_b200/host/cuspmm
ROOT=${1:-$HERE/tests/golden}; shift
OUT=ell.json
echo "[" > $OUT
for d in small_210 small_32x32 small_10x10 medium_1484 medium_2048 medium_2880 medium_4000 medium_4096 large_15120 large_20000 large_21074 large_25605; do
  [ -d "$ROOT/$d" ] || continue
  "$CLI" --ell -d "$ROOT/$d" "$@" >> $OUT || echo "cuspmm failed on $d" >&2
done
sed -i '$ s/},$/}/' $OUT
echo "]" >> $OUT
python3 -c "import json,sys; r=json.load(open('$OUT')); print(len(r), 'records,', sum(x['correct']=='1' for x in r), 'correct')"
