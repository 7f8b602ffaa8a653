#!/bin/bash
# time one modulated_conv2d shape under several env settings: exp_conv_case.sh "CIN,COUT,RES,UP" KEY=VAL ...
CASE=$1; shift
for kv in "$@"; do
  echo -n "$kv  "; env $kv timeout 120 python scripts/conv_sweep.py --only $CASE 2>&1 | tail -1 | cut -c1-260
done
