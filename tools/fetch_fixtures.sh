#!/bin/bash
# Copies the reference's INPUT DATA fixtures (network tables, trajectories, control fragments;
# no source code) from /root/reference/test into data/ so that tests/bench run on the GPU box,
# where /root/reference does not exist.  Files are copied unmodified.
set -e
REF=${1:-/root/reference}
DST=$(dirname "$0")/../data
for d in Data_alpha Data_torch47 Data_SN160 Data_SN231 Data_SN231_logft Data_XRB304; do
  mkdir -p $DST/$d
  cp -r $REF/test/$d/. $DST/$d/
done
mkdir -p $DST/Test_Problems
cp -r $REF/test/Test_Problems/th_* $REF/test/Test_Problems/setup_* $DST/Test_Problems/
cp $REF/test/test_settings* $DST/
chmod -R u+w $DST
