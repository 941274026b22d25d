#!/bin/bash
# usage (GPU box): bash scripts/dev_sbem_local.sh -- StokesBEM -local solve times, per graph-priority / reduction setting
export LD_LIBRARY_PATH=$PWD/fmm_bem_relaxed_b200:$LD_LIBRARY_PATH
B=$PWD/fmm_bem_relaxed_b200/hostcxx/bin/stokes_bem
A="-recursions 6 -p 8 -k 4 -solver_tol 1e-5"
cd /tmp
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  for i in 1 2 3; do
    echo -n "node_priority $1 m2l_reduce $2 -local: "
    FMMB_GRAPH_NODE_PRIORITY=$1 FMMB_M2L_REDUCE=$2 $B $A -local 2>&1 | grep -o "solve : [0-9.e+-]*s"
  done
  echo -n "node_priority $1 m2l_reduce $2 plain: "
  FMMB_GRAPH_NODE_PRIORITY=$1 FMMB_M2L_REDUCE=$2 $B $A 2>&1 | grep -o "solve : [0-9.e+-]*s"
done
