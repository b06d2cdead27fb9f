#!/bin/bash
# A/B of two source trees on the SAME box: tools/ab_tree.sh <other tree> -> us/step of the C2 / 32k-env rollouts for both
for t in "$1" .; do
  echo "== tree $t"
  (cd "$t" && python tools/ab_env.py "")
done
