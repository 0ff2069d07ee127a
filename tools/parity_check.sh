#!/bin/bash
# usage: tools/parity_check.sh <conf> [driver options...]   -- runs the CPU oracle and the GPU engine on the
# same config/options and diffs every lp-io output file (test/dev helper; not part of the product path)
set -u
conf=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tag=$(basename "$conf" .conf)-$(echo "$*" | tr -c 'a-zA-Z0-9=\n' '_')
o=${PARITY_OUT:-/tmp/parity}/$tag
rm -rf "$o"; mkdir -p "$o"
"$root/oracle/dfdally_oracle" --quiet "$@" --lp-io-dir="$o/oracle" -- "$conf" 2> "$o/oracle.log" || { echo "ORACLE FAILED"; cat "$o/oracle.log"; exit 1; }
"$root/codes_b200/model-net-synthetic-dragonfly-all" "$@" --lp-io-dir="$o/gpu" -- "$conf" > "$o/gpu.log" 2>&1 || { echo "GPU FAILED"; cat "$o/gpu.log"; exit 1; }
rc=0
for f in "$o"/oracle/*; do
  b=$(basename "$f")
  if cmp -s "$f" "$o/gpu/$b"; then echo "  same  $b"; else echo "  DIFF  $b"; diff "$f" "$o/gpu/$b" | head -${PARITY_LINES:-6}; rc=1; fi
done
tail -1 "$o/oracle.log"
[ $rc = 0 ] && echo "PARITY OK $tag" || echo "PARITY FAIL $tag"
exit $rc
