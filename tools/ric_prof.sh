#!/bin/bash
# usage: tools/ric_prof.sh TAG -> gpurun_out/TAG_ricprof.txt (see tools/ric_prof.py; needs lib/libh1ilqr_prof.so)
TAG=${1:-rp}; O=gpurun_out; mkdir -p $O
export H1ILQR_LIB=$PWD/mpc-ilqr-mujoco_b200/lib/libh1ilqr_prof.so
for B in 1 8192; do for dense in 0 1; do H1_RIC_DENSE=$dense python tools/ric_prof.py $B 2>&1 | tail -3; done; done > $O/${TAG}_ricprof.txt 2>&1
cat $O/${TAG}_ricprof.txt
