#!/bin/bash
# usage: tools/ref_vs_oracle.sh <conf> [driver options]  -- reference-source oracle (oracle/_ref) vs restatement oracle
conf=$(readlink -f "$1"); shift
root=$(cd "$(dirname "$0")/.." && pwd)
d=$(mktemp -d)
( cd "$(dirname "$conf")" && "$root/oracle/_ref/model-net-synthetic-dragonfly-all" --sync=1 "$@" --lp-io-dir="$d/ref" -- "$conf" > "$d/ref.log" 2>&1 ) || { echo "REF FAILED"; tail -5 "$d/ref.log"; exit 1; }
"$root/oracle/dfdally_oracle" --quiet "$@" --lp-io-dir="$d/or" -- "$conf" 2> "$d/or.log" || { echo "ORACLE FAILED"; cat "$d/or.log"; exit 1; }
rc=0
for f in dragonfly-cn-stats dragonfly-link-stats synthetic-stats model-net-category-all model-net-category-test; do
  if [ -f "$d/ref/$f" ] || [ -f "$d/or/$f" ]; then cmp -s "$d/ref/$f" "$d/or/$f" || { echo "  DIFF $f"; diff "$d/ref/$f" "$d/or/$f" | head -4; rc=1; }; fi
done
e1=$(grep "Net Events Processed" "$d/ref.log" | awk '{print $NF}'); e2=$(grep "Net Events Processed" "$d/or/summary.txt" | awk '{print $NF}')
[ "$e1" = "$e2" ] || { echo "  events differ: ref $e1 oracle $e2"; rc=1; }
# the stdout statistics block must match too
grep -A40 "^Average number of hops" "$d/ref.log" | grep -v "^$" > "$d/a.txt"; grep -A40 "^Average number of hops" "$d/or/summary.txt" | grep -v "^$" > "$d/b.txt"
cmp -s "$d/a.txt" "$d/b.txt" || { echo "  stdout block differs"; diff "$d/a.txt" "$d/b.txt" | head -6; rc=1; }
grep shim: "$d/ref.log" | tail -1
[ $rc = 0 ] && echo "REF==ORACLE $(basename $conf) $* (events $e1)" || echo "MISMATCH $(basename $conf) $*"
rm -rf "$d"; exit $rc
